// Fused pull-stream + BGK-collide + moment kernel on the box-dense SoA layout,
// with link-wise boundaries evaluated by the FLUID node ("fluid-side push").
//
// Replaces the reference's `update` + `boundary_stream` pair
// (ldc.cu:57-458, Poiseulle.cu:384-893, bifurcation.cu:429-1023,
// coronary.cu:352-944) with one launch per time step.
//
// Layout: f[q][z][y][x], x fastest, x pitch a multiple of 32 cells, one 1-D
// cell id c = x + px*(y + ny*zl).  The pull of direction q reads
// f[q][c - off_q], off_q = cx + px*cy + plane*cz: a constant shift, so a warp
// reads 32 consecutive reals per population (coalesced; the x-shifted ones
// straddle one extra sector).  No thread ever needs its coordinates on the
// bulk path.
//
// Boundaries (proven equivalent to the reference's stored-node
// scheme at every fluid node): only fluid nodes compute.  For a direction q
// whose source s = x - c_q is not fluid, the value the reference's wall / inlet /
// outlet node s would hold in slot (q,s) after `boundary_stream` depends only on
// x's own post-collision state and moments:
//     wall             : g_opp(q)(x)                               (bif:655-798)
//     inlet/outlet, q in its set:
//         feq_q(rho_bc,u_bc(s)) + (g_q(x) - feq_q(rho_x,u_x))(1-1/tau)   (bif:877-1021)
//     otherwise ("static"): the initial equilibrium of s, never rewritten
// so x itself writes it into slot (q,s) of the destination buffer at the end of
// its own update ("push into the solid node's slot").  The next step's pull is
// then uniform for every node: f_q(x) = src[q][x - c_q].
//
// One warp = one 32-cell segment.  seg[] classifies segments: BULK (all 32 are
// fluid with fluid sources only -> no node-word load, no divergence), MIXED, or
// EMPTY (nothing to do).
#pragma once
#include <cstdlib>
#include "lattice.cuh"
#include "lbm_internal.h"

namespace lbm {

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    return __ldg(p);
}

// prescribed boundary speed of BC entry `e` at boundary node (gx, gy, gz) (global coords)
template <typename T>
__device__ __forceinline__ T bc_speed(const StepParams<T> &p, const BcEntry &e, int gx, int gz) {
    T u;
    if (e.source == LBM_SRC_CONST) {
        u = (T)e.value;
    } else if (e.source == LBM_SRC_PARABOLA) {
        // pos.cu:597 -- evaluated at the boundary node's own (i,k); squares of
        // half-integers are exact, so a*a equals the reference's powf(a,2)
        T cx = T(p.box.nx - 1) / T(2.0), cz = T(p.box.nz - 1) / T(2.0), r = T(p.box.nx - 1) / T(2.0);
        T dx = T(gx) - cx, dz = T(gz) - cz;
        u = (T)e.value * (T(1.0) - (dx * dx + dz * dz) / (r * r));
    } else if (e.source == LBM_SRC_PLANE_INLET) {
        u = p.plane_in[gx + (long long)gz * p.box.nx];
    } else {
        u = p.plane_out[gx + (long long)gz * p.box.nx];
    }
    if (e.pulsatile) u = u * p.pulse_scale;
    return u;
}

// equilibrium of a prescribed axis-aligned velocity, the reference's literal
// polynomials (pos.cu:766, ldc.cu:402,441,454)
template <typename T>
__device__ __forceinline__ T feq_bc_axis(T rw, int cs, T u) {
    if (cs > 0) return rw * (T(1.0) + T(3.0) * u + T(3.0) * u * u);
    if (cs < 0) return rw * (T(1.0) - T(3.0) * u + T(3.0) * u * u);
    return rw * (T(1.0) - T(1.5) * u * u);
}

// Slow path of a boundary link (inlet / outlet / lid / static source): value this
// fluid node (cell c, moments rho/u, post-collision g_q and g_opp) must leave in
// slot (q, s = c - off_q); NaN-free sentinel `false` for a static link.  Only
// nodes next to an inlet/outlet/lid get here -- wall-only nodes bounce back
// inline -- so it is kept out of line and the bulk path stays small.
template <typename T>
__device__ __noinline__ bool boundary_link(const StepParams<T> &p, long long c, int q, T rho, T ux, T uy, T uz, T gq,
                                           T gopp, T *out) {
    const Box &b = p.box;
    const long long s = c - ((long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q));
    const int lab = p.label8[s];
    if (lab == 1) {  // half-way bounce-back
        *out = gopp;
        return true;
    }
    if (lab < 2 || lab >= LBM_MAX_BC) return false;
    const BcEntry &e = p.bc[lab];
    if (e.kind == LBM_BC_NONE || caxis(q, e.naxis) != e.nsign) return false;
    const T wden = q < 7 ? T(18.0) : T(36.0);
    const T feq = feq_lit<T>(q, rho / T(3.0), rho / T(18.0), rho / T(36.0), ux, uy, uz);
    T tmp;
    if (e.kind == LBM_BC_P) {
        const T one = T(1.0);
        tmp = feq_lit<T>(q, one / T(3.0), one / T(18.0), one / T(36.0), ux, uy, uz);
    } else {
        const int sx = (int)(s % b.px), sz = (int)(s / b.plane) + b.z0;
        const T u = bc_speed<T>(p, e, sx, sz);
        const T rw = e.kind == LBM_BC_V ? rho / wden : T(1.0) / wden;
        tmp = feq_bc_axis<T>(rw, caxis(q, e.vaxis), u);
    }
    *out = tmp + (gq - feq) * p.om1;
    return true;
}

// ---------------------------------------------------------------------------
// two-buffer pull step
// ---------------------------------------------------------------------------
// One warp per 32-cell segment, one CTA per B consecutive cells.  Two forms:
//   SPEC = false : class byte -> (node word) -> 19 population loads.  Nothing is
//                  read for EMPTY segments or non-fluid cells; the class byte
//                  (L2-resident) sits on the critical path.
//   SPEC = true  : the 19 population loads, the class byte and the node word are
//                  all issued up front, for every thread; non-fluid threads drop
//                  the values.  No dependent load precedes the DRAM reads.  Chosen
//                  by the host when >= 90 % of the launched cells are fluid (dense
//                  cavities), where the wasted reads are a few percent.
// The kernel is latency-bound per warp, so resident warps matter more than
// anything else: CTA shape and register cap are picked per precision
// (fp64: 128 threads x 5 CTAs/SM, 96 registers; fp32: 128 x 8, 64 registers).
// Measurements behind every choice here: profiles/r01_notes.md.
__host__ __device__ constexpr int cfg_block(int cfg) {
    constexpr int b[4] = {256, 256, 128, 128};
    return b[cfg];
}
__host__ __device__ constexpr int cfg_minb(int cfg) {
    constexpr int m[4] = {2, 3, 5, 8};
    return m[cfg];
}
template <typename T>
constexpr int default_cfg() {
    return sizeof(T) == 8 ? 2 : 3;
}

__device__ __forceinline__ double ld_spec(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_spec(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// collide, store, moments, boundary links of one fluid cell whose post-streaming
// populations are already in f[]
template <typename T, bool STRICT, bool MOMENTS, bool RESID>
__device__ __forceinline__ void finish_cell(const StepParams<T> &p, long long c, uint32_t node, T (&f)[Q], double &velsum) {
    const Box &b = p.box;
    T rho, ux, uy, uz;
    collide_bgk<T, STRICT>(f, p.tau, p.inv_tau, rho, ux, uy, uz);
    T *dst = p.dst + c;
#pragma unroll
    for (int q = 0; q < Q; q++) dst[(long long)q * p.qstride] = f[q];
    if (MOMENTS) {
        p.rho[c] = rho, p.ux[c] = ux, p.uy[c] = uy, p.uz[c] = uz;
    }
    if (RESID) velsum += (double)(T)sqrt((double)(ux * ux + uy * uy + uz * uz));  // |u| as ldc.cu:464 forms it
    if (node & NODE_LINKS) {
        if (node & NODE_WALLS_ONLY) {
            // half-way bounce-back, inline: slot (q, x - c_q) <- g_opp(q)(x)   (bif:781-798)
#pragma unroll
            for (int q = 1; q < Q; q++) {
                if (node & (1u << q)) {
                    const long long off = (long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q);
                    dst[(long long)q * p.qstride - off] = f[oppq(q)];
                }
            }
        } else {
#pragma unroll
            for (int q = 1; q < Q; q++) {
                if (node & (1u << q)) {
                    T h;
                    if (boundary_link<T>(p, c, q, rho, ux, uy, uz, f[q], f[oppq(q)], &h)) {
                        const long long off = (long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q);
                        dst[(long long)q * p.qstride - off] = h;
                    }
                }
            }
        }
    }
}

template <typename T, bool STRICT, bool MOMENTS, bool RESID, bool SPEC, int CFG>
__global__ void __launch_bounds__(cfg_block(CFG), cfg_minb(CFG)) k_step_dense_ab(const __grid_constant__ StepParams<T> p) {
    const Box &b = p.box;
    const long long c = p.c_begin + (long long)blockIdx.x * cfg_block(CFG) + threadIdx.x;
    if (c >= p.c_end) return;  // ranges are whole planes (multiples of 32 cells): warp-uniform
    const T *src = p.src + c;
    T f[Q];
    uint32_t node;
    const uint32_t kind = p.seg[c >> 5];
    if (SPEC) {
        // every thread pulls; the buffers carry a tail guard so all addresses are mapped
        node = p.node[c];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const long long off = (long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q);
            f[q] = ld_spec(src + (long long)q * p.qstride - off);
        }
        if (kind == SEG_EMPTY) return;
    } else {
        if (kind == SEG_EMPTY) return;
        node = kind == SEG_MIXED ? p.node[c] : 0u;
    }
    double velsum = 0.0;
    if (!(node & NODE_SKIP)) {
        if (!SPEC) {
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const long long off = (long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q);
                f[q] = ld_stream(src + (long long)q * p.qstride - off);
            }
        }
        finish_cell<T, STRICT, MOMENTS, RESID>(p, c, node, f, velsum);
    }
    if (RESID) {
        // warp shuffle tree, then one atomic per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) velsum += __shfl_xor_sync(0xffffffffu, velsum, o);
        if ((threadIdx.x & 31) == 0 && velsum != 0.0) atomicAdd(p.resid, velsum);
    }
}

template <typename T, bool STRICT, bool MOMENTS, bool RESID, int CFG>
cudaError_t launch_cfg(const StepParams<T> &p, cudaStream_t s) {
    constexpr int B = cfg_block(CFG);
    const unsigned nb = (unsigned)((p.c_end - p.c_begin + B - 1) / B);
    if (p.speculative) k_step_dense_ab<T, STRICT, MOMENTS, RESID, true, CFG><<<nb, B, 0, s>>>(p);
    else k_step_dense_ab<T, STRICT, MOMENTS, RESID, false, CFG><<<nb, B, 0, s>>>(p);
    return cudaGetLastError();
}

template <typename T, bool STRICT>
cudaError_t launch_step_dense_impl(const StepParams<T> &p, bool moments, bool resid, int storage, cudaStream_t s) {
    if (p.c_end <= p.c_begin) return cudaSuccess;
    (void)storage;
    constexpr int C = default_cfg<T>();
    if (moments && resid) return launch_cfg<T, STRICT, true, true, C>(p, s);
    if (moments) return launch_cfg<T, STRICT, true, false, C>(p, s);
    if (resid) return launch_cfg<T, STRICT, false, true, C>(p, s);
    return launch_cfg<T, STRICT, false, false, C>(p, s);
}

}  // namespace lbm
