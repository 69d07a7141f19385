// Case pre-processing on the GPU: node labels (geo_pre), outer-wall-neighbour
// marking, stream compaction (index_transform), node words for the fused step,
// the initial equilibrium state (initialize), and the small gather / reduce /
// halo kernels around the hot loop.
//
// Reference (serial host code): ldc.cu:468-580, Poiseulle.cu:52-382,
// bifurcation.cu:36-427, coronary.cu:31-350.  All integer results are
// bit-exact with it (tests compare against the CPU oracle).
//
// This file is compiled with -fmad=false: the initial state is written with the
// reference's literal expression order so that STRICT runs start bit-identical.
#include "lattice.cuh"
#include "lbm_internal.h"
#include "init_rule.cuh"

namespace lbm {

namespace {

__host__ __device__ inline long long cell_of(const Box &b, int x, int y, int z) {
    return (long long)x + (long long)b.px * ((long long)y + (long long)b.ny * (long long)(z - b.z0));
}

struct Coord {
    int x, y, z;  // global
};
__device__ inline Coord coord_of(const Box &b, long long c) {
    Coord r;
    r.x = (int)(c % b.px);
    long long t = c / b.px;
    r.y = (int)(t % b.ny);
    r.z = (int)(t / b.ny) + b.z0;
    return r;
}

// binary voxel flag; 0 outside the global box, outside the held z range, or in the x padding
__device__ inline int flag_at(const uint8_t *flag, const Box &b, int x, int y, int z) {
    if (x < 0 || x >= b.nx || y < 0 || y >= b.ny || z < b.z0 || z >= b.z1 || z < 0 || z >= b.nz) return 0;
    return flag[cell_of(b, x, y, z)];
}
__device__ inline int imin(int a, int b) { return a < b ? a : b; }
__device__ inline int min6(const uint8_t *f, const Box &b, int x, int y, int z) {
    int mx = imin(flag_at(f, b, x + 1, y, z), flag_at(f, b, x - 1, y, z));
    int my = imin(flag_at(f, b, x, y - 1, z), flag_at(f, b, x, y + 1, z));
    int mz = imin(flag_at(f, b, x, y, z - 1), flag_at(f, b, x, y, z + 1));
    return imin(imin(mx, my), mz);
}
__device__ inline bool interior(const Box &b, int x, int y, int z) {
    return x >= 1 && x <= b.nx - 2 && y >= 1 && y <= b.ny - 2 && z >= 1 && z <= b.nz - 2;
}

// Poiseuille binary field: disc in (x,z), float arithmetic (Poiseulle.cu:80-91)
__global__ void k_make_flag_pos(uint8_t *flag, Box b) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.cells()) return;
    Coord p = coord_of(b, c);
    uint8_t v = 0;
    if (p.x < b.nx && p.y >= 1 && p.y <= b.ny - 2) {
        float radius = (b.nx - 1) / 2.0f, cx = (b.nx - 1) / 2.0f, cz = (b.nz - 1) / 2.0f;
        float dx = p.x - cx, dz = p.z - cz;
        float dist = sqrtf(dx * dx + dz * dz);
        v = dist <= radius ? 1 : 0;
    }
    flag[c] = v;
}

// label before the inlet/outlet plane rules of the GEO_Y_INOUT case (bifurcation.cu:63-91)
__device__ inline int bif_base(const uint8_t *f, const Box &b, int x, int y, int z) {
    int g = flag_at(f, b, x, y, z);
    bool xz_in = x >= 1 && x <= b.nx - 2 && z >= 1 && z <= b.nz - 2;
    if (xz_in && (y == 0 || y == b.ny - 1)) g = 0;
    if (xz_in && y >= 2 && y <= b.ny - 3) g += 3 * min6(f, b, x, y, z);
    return g;
}

__global__ void k_labels(const uint8_t *flag, int32_t *label, Box b, GeoRules r) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.cells()) return;
    Coord p = coord_of(b, c);
    int g = 0;
    if (p.x < b.nx) {
        const int x = p.x, y = p.y, z = p.z;
        if (r.case_rule == LBM_CASE_LDC) {
            // ldc.cu:468-502: layers 0 / 1 / 3, lid 2 on y = NY-2
            auto in = [&](int m) {
                return x >= m && x <= b.nx - 1 - m && y >= m && y <= b.ny - 1 - m && z >= m && z <= b.nz - 1 - m;
            };
            g = in(2) ? 3 : (in(1) ? 1 : 0);
            if (y == b.ny - 2 && x >= 1 && x <= b.nx - 2 && z >= 1 && z <= b.nz - 2) g = 2;
        } else if (r.case_rule == LBM_CASE_POISEUILLE) {
            // Poiseulle.cu:94-134
            g = flag_at(flag, b, x, y, z);
            bool xz_in = x >= 1 && x <= b.nx - 2 && z >= 1 && z <= b.nz - 2;
            if (xz_in && y >= 2 && y <= b.ny - 3) g += 3 * min6(flag, b, x, y, z);
            if (xz_in && (y == 1 || y == b.ny - 2)) {
                int m4 = imin(imin(flag_at(flag, b, x + 1, y, z), flag_at(flag, b, x - 1, y, z)),
                              imin(flag_at(flag, b, x, y, z - 1), flag_at(flag, b, x, y, z + 1)));
                if (y == 1) g += m4;
                if (y == b.ny - 2) g += 2 * m4;  // both apply if NY == 3; not a real case
            }
        } else if (r.case_rule == LBM_CASE_GEO_Y_INOUT) {
            // bifurcation.cu:63-119
            g = bif_base(flag, b, x, y, z);
            bool xz_in = x >= 1 && x <= b.nx - 2 && z >= 1 && z <= b.nz - 2;
            if (xz_in && y == 1) {
                int g2 = bif_base(flag, b, x, 2, z);
                g = g2 == 1 ? 1 : (g2 == 4 ? 2 : 0);
            }
            if (xz_in && y == b.ny - 2) {
                // the reference copies from y = NY-3 *after* the inlet rule ran
                int g2 = bif_base(flag, b, x, b.ny - 3, z);
                if (b.ny - 3 == 1) {
                    int g3 = bif_base(flag, b, x, 2, z);
                    g2 = g3 == 1 ? 1 : (g3 == 4 ? 2 : 0);
                }
                g = g2 == 1 ? 1 : (g2 == 4 ? 3 : 0);
            }
        } else {
            // coronary.cu:60-141
            g = flag_at(flag, b, x, y, z);
            if (interior(b, x, y, z)) g += 3 * min6(flag, b, x, y, z);
            for (int k = 0; k < r.n_open; k++) {
                const lbm_opening_rule &o = r.open[k];
                int pc = o.axis == 0 ? x : (o.axis == 1 ? y : z);
                int a = o.axis == 0 ? y : x;
                int bb = o.axis == 2 ? y : z;
                if (pc != o.coord || a < o.lo_a || a > o.hi_a || bb < o.lo_b || bb > o.hi_b) continue;
                int m;
                if (o.axis == 0)
                    m = imin(imin(flag_at(flag, b, x, y - 1, z), flag_at(flag, b, x, y + 1, z)),
                             imin(flag_at(flag, b, x, y, z - 1), flag_at(flag, b, x, y, z + 1)));
                else if (o.axis == 1)
                    m = imin(imin(flag_at(flag, b, x - 1, y, z), flag_at(flag, b, x + 1, y, z)),
                             imin(flag_at(flag, b, x, y, z - 1), flag_at(flag, b, x, y, z + 1)));
                else
                    m = imin(imin(flag_at(flag, b, x, y - 1, z), flag_at(flag, b, x, y + 1, z)),
                             imin(flag_at(flag, b, x - 1, y, z), flag_at(flag, b, x + 1, y, z)));
                g += o.reps * m;
            }
        }
    }
    label[c] = g;
}

// Outer-wall-neighbour marking, gather form (Poiseulle.cu:138-254,
// bifurcation.cu:123-239): a label-0 node becomes -1 when one of its 18
// neighbours is a marking source lying inside [1,N-2]^3.  In place: only 0 -> -1
// transitions happen and neither value is a source, so order does not matter.
__global__ void k_mark(int32_t *label, Box b, GeoRules r) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.cells()) return;
    Coord p = coord_of(b, c);
    if (p.x >= b.nx || label[c] != 0) return;
    bool hit = false;
#pragma unroll
    for (int q = 1; q < Q; q++) {
        int x = p.x + cxq(q), y = p.y + cyq(q), z = p.z + czq(q);
        if (!interior(b, x, y, z) || z < b.z0 || z >= b.z1) continue;
        int g = label[cell_of(b, x, y, z)];
        if (g > 0 && g < 32 && ((r.mark_sources >> g) & 1u)) hit = true;
    }
    if (hit) label[c] = -1;
}

// ---------------------------------------------------------------- compaction
// index_transform (Poiseulle.cu:257-271) as a three-kernel exclusive scan over
// the array in memory order (= z,y,x; x-padding cells carry label 0).
// "stored" = has an entry in the reference's compact arrays: label != 0, or -- for the LDC
// rule, which stores the whole box (ldc.cu:54) -- any cell inside the box (not x padding).
struct StoredRule {
    int px, nx, all;
};
__device__ inline bool is_stored(const int32_t *label, long long c, StoredRule r) {
    if (r.all) return (int)(c % r.px) < r.nx;
    return label[c] != 0;
}
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 16;  // cells per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ inline int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[SCAN_BLOCK / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_BLOCK / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    int base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[SCAN_BLOCK / 32 - 1];
    __syncthreads();
    return base + incl - v;
}

__global__ void k_tile_counts(const int32_t *label, long long cells, StoredRule sr, int32_t *tile_count) {
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        long long c = base + i;
        if (c < cells && is_stored(label, c, sr)) cnt++;
    }
    int total;
    block_exclusive_scan(cnt, &total);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = total;
}
// single block: exclusive scan of the tile counts (carry across chunks); 64-bit totals
__global__ void k_scan_tiles(const int32_t *tile_count, long long *tile_offset, int ntiles, long long *total_out) {
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int start = 0; start < ntiles; start += SCAN_BLOCK) {
        int i = start + threadIdx.x;
        int v = i < ntiles ? tile_count[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_offset[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
__global__ void k_scatter_index(const int32_t *label, int32_t *index, long long cells, StoredRule sr,
                                long long base_index, const long long *tile_offset, long long plane,
                                long long *plane_first) {
    long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int flags[SCAN_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        long long c = base + i;
        flags[i] = (c < cells && is_stored(label, c, sr)) ? 1 : 0;
        cnt += flags[i];
    }
    int total;
    int ex = block_exclusive_scan(cnt, &total);
    long long run = base_index + tile_offset[blockIdx.x] + ex;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        long long c = base + i;
        if (c < cells) {
            if (plane_first && c % plane == 0) plane_first[c / plane] = run;  // compact id the plane starts at
            index[c] = flags[i] ? (int32_t)run : -1;
            run += flags[i];
        }
    }
}

__global__ void k_count_stored(const int32_t *label, long long cells, StoredRule sr, long long *out) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int v = (c < cells && is_stored(label, c, sr)) ? 1 : 0;
    unsigned m = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long *)out, (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------- node words
// One thread per cell of the state box.  Fluid nodes of the OWNED planes get
// the 18 "source is not fluid" bits; everything else is NODE_SKIP.
struct BcTable {
    BcEntry e[LBM_MAX_BC];
};
__global__ void k_node_words(const int32_t *label, uint32_t *node, uint32_t *wall, uint8_t *seg, int8_t *label8, Box b,
                             int own_z0, int own_z1, int fluid_label, BcTable bc, long long *nfluid) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = c < b.cells();
    uint32_t w = NODE_SKIP, wm = 0;
    if (valid) {
        Coord p = coord_of(b, c);
        int g = label[c];
        label8[c] = (int8_t)g;
        if (p.x < b.nx && g == fluid_label && p.z >= own_z0 && p.z < own_z1) {
            w = 0;
            bool walls_only = true;
#pragma unroll
            for (int q = 1; q < Q; q++) {
                int x = p.x - cxq(q), y = p.y - cyq(q), z = p.z - czq(q);
                bool inb = x >= 0 && x < b.nx && y >= 0 && y < b.ny && z >= b.z0 && z < b.z1;
                int gs = inb ? label[cell_of(b, x, y, z)] : 0;
                if (gs != fluid_label) {
                    w |= (1u << q);
                    if (gs == 1) wm |= (1u << q);
                    else walls_only = false;
                    if (gs >= 2 && gs < LBM_MAX_BC && bc.e[gs].kind != LBM_BC_NONE && caxis(q, bc.e[gs].naxis) == bc.e[gs].nsign)
                        w |= NODE_HAS_BC;
                }
            }
            if (w && walls_only) w |= NODE_WALLS_ONLY;
        }
        node[c] = w;
        wall[c] = wm;
    }
    unsigned fluid_m = __ballot_sync(0xffffffffu, valid && !(w & NODE_SKIP));
    unsigned link_m = __ballot_sync(0xffffffffu, valid && (w & NODE_LINKS));
    if ((threadIdx.x & 31) == 0 && valid) {
        uint8_t k = fluid_m == 0 ? SEG_EMPTY : ((fluid_m == 0xffffffffu && link_m == 0) ? SEG_BULK : SEG_MIXED);
        seg[c >> 5] = k;
        if (fluid_m) atomicAdd((unsigned long long *)nfluid, (unsigned long long)__popc(fluid_m));
    }
}

// ---------------------------------------------------------------- initialize
// rho = 1, u = 0 on every cell; boundary planes / labels get their initial
// velocity; f = feq(rho,u) into both buffers (ldc.cu:504-580, pos:273-382,
// bif:329-427, cor:277-350).  Cells the reference does not store (label 0,
// padding) are given the rest state, the value "static" links read.
// In-place (AA) storage starts with an even, purely local step, so slot (q,c) is
// seeded with the population that step would have pulled: feq_q of cell c - c_q.
template <typename T>
__global__ void k_init(const __grid_constant__ InitParams<T> p) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const Box &b = p.box;
    if (c >= b.cells()) return;
    Coord co = coord_of(b, c);
    T feq[Q];
    if (!p.aa) {
        T ux, uy, uz;
        const int g = co.x < b.nx ? p.label[c] : 0;
        init_velocity<T>(p.case_rule, p.u_max, p.bc, p.plane_in, p.plane_out, b, g, co.x < b.nx ? co.x : -1, co.y, co.z,
                         ux, uy, uz);
        if (p.case_rule == LBM_CASE_LDC) {
            feq_all_ldc_init<T>(T(1.0), ux, uy, uz, feq);
        } else {
            const T r3 = T(1.0) / T(3.0), r18 = T(1.0) / T(18.0), r36 = T(1.0) / T(36.0);
#pragma unroll
            for (int q = 0; q < Q; q++) feq[q] = feq_lit<T>(q, r3, r18, r36, ux, uy, uz);
        }
    } else {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int x = co.x - cxq(q), y = co.y - cyq(q), z = co.z - czq(q);
            const bool in = co.x < b.nx && x >= 0 && x < b.nx && y >= 0 && y < b.ny && z >= b.z0 && z < b.z1;
            const int g = in ? p.label[cell_of(b, x, y, z)] : 0;
            T ux, uy, uz;
            init_velocity<T>(p.case_rule, p.u_max, p.bc, p.plane_in, p.plane_out, b, g, in ? x : -1, y, z, ux, uy, uz);
            feq[q] = init_feq_q<T>(p.case_rule, q, T(1.0), ux, uy, uz);
        }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
        p.fa[(long long)q * p.qstride + c] = feq[q];
        if (p.fb != p.fa) p.fb[(long long)q * p.qstride + c] = feq[q];
    }
    p.rho[c] = T(0), p.ux[c] = T(0), p.uy[c] = T(0), p.uz[c] = T(0);
}

// ---------------------------------------------------------------- gathers
// `sid` (optional): the moment arrays are not box-dense but numbered by sid[cell] (in-place sparse storage)
template <typename T>
__global__ void k_gather_fields(const T *rho, const T *ux, const T *uy, const T *uz, const int32_t *label,
                                const int32_t *index, const int32_t *sid, Box b, long long c0, long long c1, int fluid_label,
                                long long first, T *orho, T *oux, T *ouy, T *ouz) {
    long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    int i = index[c];
    if (i < 0) return;
    long long o = (long long)i - first;
    bool fl = label[c] == fluid_label;
    const long long e = (fl && sid) ? (long long)sid[c] : c;
    orho[o] = fl ? rho[e] : T(0);
    oux[o] = fl ? ux[e] : T(0);
    ouy[o] = fl ? uy[e] : T(0);
    ouz[o] = fl ? uz[e] : T(0);
}
template <typename T>
__global__ void k_gather_pops(const T *f, long long qstride, const int32_t *index, Box b, long long c0, long long c1,
                              long long first, long long count, int layout, T *out) {
    long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    int i = index[c];
    if (i < 0) return;
    long long o = (long long)i - first;
    const long long cells = b.cells();
    for (int q = 0; q < Q; q++) {
        // "as if in d_scr after the swap": slot q of stored node y = what y + c_q pulls next
        long long src;
        if (layout == 0) src = (long long)q * qstride + c;
        else if (layout == 1) src = (long long)q * qstride + c + ((long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q));
        else src = (long long)oppq(q) * qstride + c;
        const long long cell = src - (long long)(layout == 2 ? oppq(q) : q) * qstride;
        out[(long long)q * count + o] = (cell >= 0 && cell < cells) ? f[src] : T(0);
    }
}

// kind 0: sum sqrt(u^2) over every stored entry (non-fluid entries are 0)    ldc.cu:460-466,662
// kind 1: sum u^2 over fluid nodes of the trimmed box z[1,NZ-2] y[2,NY-3] x[1,NX-2]   bif:1158-1175
template <typename T>
__global__ void k_reduce_fields(const T *ux, const T *uy, const T *uz, const int32_t *label, Box b, long long c0,
                                long long c1, int kind, int fluid_label, int case_rule, double *out) {
    long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (c < c1) {
        int g = label[c];
        Coord p = coord_of(b, c);
        if (p.x < b.nx) {
            T s = ux[c] * ux[c] + uy[c] * uy[c] + uz[c] * uz[c];
            if (kind == 0) {
                if (g == fluid_label) v = (double)(T)sqrt((double)s);
            } else {
                bool ok = case_rule == LBM_CASE_GEO_OPENINGS ? (g == 4) : (g >= 4);
                if (case_rule == LBM_CASE_LDC) ok = g == fluid_label;
                bool trimmed = p.x >= 1 && p.x <= b.nx - 2 && p.y >= 2 && p.y <= b.ny - 3 && p.z >= 1 && p.z <= b.nz - 2;
                if (ok && trimmed) v = (double)s;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += ws[i];
        if (t != 0.0) atomicAdd(out, t);
    }
}

// ---------------------------------------------------------------- halo planes
// side 0 (low-z face): the plane's populations with c_z = -1 leave; side 1: c_z = +1.
__device__ inline int halo_q(int side, int k) {
    const int up[5] = {5, 11, 13, 15, 16};    // c_z = +1
    const int down[5] = {6, 12, 14, 17, 18};  // c_z = -1
    return side ? up[k] : down[k];
}
template <typename T>
__global__ void k_halo_pack(const T *f, long long qstride, Box b, int zl, int side, T *buf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.plane) return;
    long long c = (long long)zl * b.plane + i;
#pragma unroll
    for (int k = 0; k < 5; k++) buf[(long long)k * b.plane + i] = f[(long long)halo_q(side, k) * qstride + c];
}
// `side` is the face of the RECEIVING slab: side 0 receives the c_z = +1 set into its low halo plane
template <typename T>
__global__ void k_halo_unpack(T *f, long long qstride, const int8_t *label8, int fluid_label, Box b, int zl, int side,
                              const T *buf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.plane) return;
    long long c = (long long)zl * b.plane + i;
    if (label8[c] != fluid_label) return;  // slots of solid halo nodes belong to the local fluid neighbour
#pragma unroll
    for (int k = 0; k < 5; k++) f[(long long)halo_q(1 - side, k) * qstride + c] = buf[(long long)k * b.plane + i];
}

// ---------------------------------------------------------------- sparse storage
// LBM_STORE_SPARSE_AB keeps the populations in the reference's own compact order
// (one entry per stored node, geo != 0).  Consecutive stored x positions of a row have
// consecutive compact ids and every source of a fluid node is a stored node (the -1
// marking guarantees it), so for a run of fluid nodes of one row the sources of each
// direction are one contiguous run of compact ids: 19 base ids per run, not 19 ids per
// node.
//
// A warp of the step kernel owns one aligned chunk of 32 compact ids (all its loads and
// stores of direction 0 and all its stores are 256-byte aligned rows of the SoA arrays).
// A chunk usually holds the tail of one row's run and the head of the next one, so a
// segment record describes up to TWO "pieces" (= run intersected with the chunk, one
// plane); further pieces of a chunk, and pieces of another z-plane, go to further records
// of the same chunk.  Record = 2 x 24 int32, per piece:
//   [0..18] source compact id of direction q for LANE 0 of the chunk (source of lane l is
//           base + l; [0] is the chunk's first id), [19] first lane | length << 8 (0 = no
//           piece), [20],[21] Cartesian cell id of lane 0 (lo, hi), [22] 1 if any node of
//           the piece has a boundary link, [23] unused.
// row_src(r): source offset of the 8 neighbouring rows, as the direction with c_x = 0 in that row
__device__ __forceinline__ int row_src_dir(int r) {
    const int a[8] = {3, 4, 5, 6, 15, 16, 17, 18};
    return a[r];
}
constexpr int ROW_INVALID = INT32_MIN;
// id of the node at THIS lane's x in neighbouring row r, derived from any fluid source the lane has there
// (x, x-1, x+1 for the rows that hold three directions); ROW_INVALID when the lane has no fluid source in that row
__device__ __forceinline__ int row_anchor(const int32_t *sid, const int32_t *label, const Box &b, int fluid_label,
                                          long long cartc, int r) {
    const int k = row_src_dir(r);
    const long long c0 = cartc - ((long long)b.px * cyq(k) + b.plane * czq(k));
    const int x = (int)(cartc % b.px);
    const int ndx = r < 4 ? 3 : 1;
    for (int t = 0; t < ndx; t++) {
        const int dx = t == 0 ? 0 : (t == 1 ? -1 : 1);
        if (x + dx < 0 || x + dx >= b.nx) continue;
        if (label[c0 + dx] == fluid_label) return sid[c0 + dx] - dx;
    }
    return ROW_INVALID;
}
struct ChunkScan {
    bool fluid, start;
    long long cartc;
    int my_rec, my_slot, nrec;
};
// rows != nullptr (in-place storage, own numbering): a piece additionally ends where, in one of the 8
// neighbouring rows, the fluid sources stop being one consecutive id range (anchors[r] receives this
// lane's id of "the node at my x" in row r, or ROW_INVALID)
struct RowInfo {
    const int32_t *sid, *label;
    int fluid_label;
};
__device__ __forceinline__ ChunkScan scan_chunk(const uint32_t *nodec, const long long *cart, Box b, long long i,
                                                long long id0, long long id1, const RowInfo *rows = nullptr,
                                                int *anchors = nullptr) {
    const int lane = threadIdx.x & 31;
    ChunkScan r;
    const bool valid = i >= id0 && i < id1;
    r.fluid = valid && !(nodec[i] & NODE_SKIP);
    r.cartc = valid ? cart[i] : -2;
    const long long prev_c = __shfl_up_sync(0xffffffffu, r.cartc, 1);
    const bool prev_f = __shfl_up_sync(0xffffffffu, r.fluid ? 1 : 0, 1) != 0;
    const bool joined = lane > 0 && prev_f && prev_c == r.cartc - 1 && (r.cartc % b.px) != 0;
    r.start = r.fluid && !joined;
    if (rows) {
        const unsigned starts = __ballot_sync(0xffffffffu, r.start);
        const unsigned le = lane == 31 ? 0xffffffffu : ((2u << lane) - 1u), lt = (1u << lane) - 1u;
        const unsigned mine = starts & le;
        const int rs = mine ? 31 - __clz(mine) : 0;  // first lane of my run
        bool brk = false;
        for (int rr = 0; rr < 8; rr++) {
            const int v = r.fluid ? row_anchor(rows->sid, rows->label, b, rows->fluid_label, r.cartc, rr) : ROW_INVALID;
            const int val = v == ROW_INVALID ? ROW_INVALID : v - lane;  // constant along a consecutive id range
            anchors[rr] = val;
            const unsigned V = __ballot_sync(0xffffffffu, val != ROW_INVALID);
            const unsigned prev = V & lt & ~((1u << rs) - 1u);
            const int lp = prev ? 31 - __clz(prev) : 0;
            const int other = __shfl_sync(0xffffffffu, val, lp);
            if (r.fluid && val != ROW_INVALID && prev && other != val) brk = true;
        }
        r.start = r.start || brk;
    }
    const long long plane = r.fluid ? r.cartc / b.plane : -1;
    unsigned remaining = __ballot_sync(0xffffffffu, r.start);
    r.my_rec = 0, r.my_slot = 0, r.nrec = 0;
    while (remaining) {  // one round per z-plane present in the chunk (ids are plane-ordered)
        const int leader = __ffs(remaining) - 1;
        const long long pl = __shfl_sync(0xffffffffu, plane, leader);
        const unsigned grp = __ballot_sync(0xffffffffu, r.start && plane == pl);
        if (r.start && plane == pl) {
            const int j = __popc(grp & ((1u << lane) - 1u));
            r.my_rec = r.nrec + (j >> 1), r.my_slot = j & 1;
        }
        r.nrec += (__popc(grp) + 1) >> 1;
        remaining &= ~grp;
    }
    return r;
}
__global__ void k_seg_count(const uint32_t *nodec, const long long *cart, Box b, long long base_id, long long id0,
                            long long id1, int32_t *counts) {
    const long long i = base_id + (long long)blockIdx.x * blockDim.x + threadIdx.x;  // base_id multiple of 32
    const ChunkScan r = scan_chunk(nodec, cart, b, i, id0, id1);
    if ((threadIdx.x & 31) == 0 && i < id1) counts[(i - base_id) >> 5] = r.nrec;
}
__global__ void k_seg_fill(const uint32_t *nodec, const long long *cart, const int32_t *index, Box b, int own_zl0,
                           long long base_id, long long id0, long long id1, long long id_first,
                           const long long *chunk_offset, int32_t *rec, long long *plane_seg) {
    const long long i = base_id + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const ChunkScan r = scan_chunk(nodec, cart, b, i, id0, id1);
    const unsigned cont = __ballot_sync(0xffffffffu, r.fluid && !r.start);
    const bool has_link = r.fluid ? (nodec[i] & NODE_LINKS) != 0 : false;
    const unsigned links = __ballot_sync(0xffffffffu, has_link);
    if (!r.start) return;  // one thread per piece: its first lane
    int len = 1;
    if (lane < 31) {
        const unsigned after = (~cont) >> (lane + 1);  // first lane after me that does not continue my run
        len = 1 + (after ? __ffs(after) - 1 : 31 - lane);
        if (len > 32 - lane) len = 32 - lane;
    }
    const unsigned piece = (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << lane;
    const long long seg = chunk_offset[(i - base_id) >> 5] + r.my_rec;
    int32_t *o = rec + seg * SEG_REC + r.my_slot * SEG_HALF;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const long long sc = r.cartc - ((long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q));
        // every source of a fluid node is stored when the mask came from geo_pre; an
        // unstored one (the reference would read d_scr[q*NLATTICE-1]) is clamped for memory safety
        long long id = (long long)index[sc] - id_first;
        if (id < 0) id = 0;
        o[q] = (int32_t)(id - lane);
    }
    o[19] = lane | (len << 8);
    const long long c_lane0 = r.cartc - lane;
    o[20] = (int32_t)(c_lane0 & 0xffffffffLL);
    o[21] = (int32_t)(c_lane0 >> 32);
    o[22] = (links & piece) ? 1 : 0;
    o[23] = 0;
    if (r.my_slot == 0) atomicMin(plane_seg + (r.cartc / b.plane - own_zl0), seg);
}
// Records of the in-place storage: same layout as k_seg_fill, but per piece only the EIGHT row ids are
// filled in (slot of the c_x = 0 direction of each row: base + lane = id of the node at the lane's x in
// that row), taken from the first lane of the piece that has a fluid source in the row.
__global__ void k_seg_count_rows(const uint32_t *nodec, const long long *cart, RowInfo rows, Box b, long long base_id,
                                 long long id0, long long id1, int32_t *counts) {
    const long long i = base_id + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int anchors[8];
    const ChunkScan r = scan_chunk(nodec, cart, b, i, id0, id1, &rows, anchors);
    if ((threadIdx.x & 31) == 0 && i < id1) counts[(i - base_id) >> 5] = r.nrec;
}
__global__ void k_seg_fill_rows(const uint32_t *nodec, const long long *cart, RowInfo rows, Box b, int own_zl0,
                                long long base_id, long long id0, long long id1, const long long *chunk_offset,
                                int32_t *rec, long long *plane_seg) {
    const long long i = base_id + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int anchors[8];
    const ChunkScan r = scan_chunk(nodec, cart, b, i, id0, id1, &rows, anchors);
    const unsigned cont = __ballot_sync(0xffffffffu, r.fluid && !r.start);
    const bool has_link = r.fluid ? (nodec[i] & NODE_LINKS) != 0 : false;
    const unsigned links = __ballot_sync(0xffffffffu, has_link);
    int len = 1;
    if (lane < 31) {
        const unsigned after = (~cont) >> (lane + 1);  // first lane after me that does not continue my piece
        len = 1 + (after ? __ffs(after) - 1 : 31 - lane);
        if (len > 32 - lane) len = 32 - lane;
    }
    const unsigned piece = (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << lane;
    int base[8];
    for (int rr = 0; rr < 8; rr++) {  // every lane takes part in the shuffles; only piece starts use the result
        const unsigned V = __ballot_sync(0xffffffffu, anchors[rr] != ROW_INVALID) & piece;
        const int lp = V ? __ffs(V) - 1 : lane;
        const int v = __shfl_sync(0xffffffffu, anchors[rr], lp);
        base[rr] = (V && r.start) ? v : 0;
    }
    if (!r.start) return;  // one thread per piece: its first lane
    const long long seg = chunk_offset[(i - base_id) >> 5] + r.my_rec;
    int32_t *o = rec + seg * SEG_REC + r.my_slot * SEG_HALF;
    for (int q = 0; q < SEG_HALF; q++) o[q] = 0;
    o[0] = (int32_t)(i - lane);
    for (int rr = 0; rr < 8; rr++) o[row_src_dir(rr)] = base[rr];
    o[19] = lane | (len << 8);
    const long long c_lane0 = r.cartc - lane;
    o[20] = (int32_t)(c_lane0 & 0xffffffffLL);
    o[21] = (int32_t)(c_lane0 >> 32);
    o[22] = (links & piece) ? 1 : 0;
    if (r.my_slot == 0) atomicMin(plane_seg + (r.cartc / b.plane - own_zl0), seg);
}
// Inlet / outlet links of the in-place sparse storage, precomputed once per initialize(): the same decisions
// boundary_node (step_dense.cuh) takes every step -- source label, is the direction in that label's set, which
// speed does the source node prescribe -- stored as a short list per node.
template <typename T>
__global__ void k_bc_links(const uint32_t *nodec, const uint32_t *wallc, const long long *cart, const int8_t *label8, Box b,
                           BcTable bc, const T *plane_in, const T *plane_out, long long i0, long long i1, int *total,
                           int32_t *bcslot, BcLink<T> *links) {
    const long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const uint32_t node = nodec[i];
    if ((node & NODE_SKIP) || !(node & NODE_HAS_BC)) return;
    const uint32_t rest = node & NODE_LINKS & ~wallc[i];
    const long long c = cart[i];
    const Coord co = coord_of(b, c);
    uint32_t bcm = 0u;
    int lab[Q];
    for (int q = 1; q < Q; q++) {
        lab[q] = 0;
        if (!(rest & (1u << q))) continue;
        const int l = label8[c - ((long long)cxq(q) + (long long)b.px * cyq(q) + b.plane * czq(q))];
        lab[q] = l;
        if (l >= 2 && l < LBM_MAX_BC && bc.e[l].kind != LBM_BC_NONE && caxis(q, bc.e[l].naxis) == bc.e[l].nsign) bcm |= 1u << q;
    }
    const int n = 1 + __popc(bcm);
    if (!links) {
        atomicAdd(total, n);
        return;
    }
    const int slot = atomicAdd(total, n);
    bcslot[i] = slot;
    BcLink<T> *o = links + slot;
    o[0].meta = bcm, o[0].pad = 0u, o[0].ubc = T(0);
    int k = 1;
    for (int q = 1; q < Q; q++) {
        if (!(bcm & (1u << q))) continue;
        const BcEntry &e = bc.e[lab[q]];
        o[k].meta = (uint32_t)e.kind | ((uint32_t)(caxis(q, e.vaxis) + 1) << 2) | (e.pulsatile ? 16u : 0u);
        o[k].pad = 0u;
        // the speed is sampled at the boundary node s = x - c_q itself (pos.cu:597)
        o[k].ubc = e.kind == LBM_BC_P ? T(0) : bc_speed_unscaled<T>(e, b, plane_in, plane_out, co.x - cxq(q), co.z - czq(q));
        k++;
    }
}
// One link word per lane of every record (NODE_SKIP for lanes outside the record's pieces): the neighbour
// (odd) step of the in-place storage loads it NEXT TO the record instead of fetching the node word by the
// compact id the record holds, which took a dependent round trip to memory out of every warp's life.
__global__ void k_rec_links(const int32_t *rec, const uint32_t *nodec, long long nseg, uint32_t *links) {
    const long long seg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (seg >= nseg) return;
    const int32_t *r = rec + seg * SEG_REC;
    const int mA = r[19], mB = r[SEG_HALF + 19];
    const bool inA = lane >= (mA & 255) && lane < (mA & 255) + (mA >> 8);
    const bool inB = (mB >> 8) != 0 && lane >= (mB & 255) && lane < (mB & 255) + (mB >> 8);
    links[seg * 32 + lane] = (inA || inB) ? nodec[r[0] + lane] : NODE_SKIP;
}
// exclusive scan of int32 counts into int64 offsets (single block, carry across chunks)
__global__ void k_scan_i32(const int32_t *counts, long long *offsets, long long n, long long *total_out) {
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long start = 0; start < n; start += SCAN_BLOCK) {
        long long i = start + threadIdx.x;
        int v = i < n ? counts[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (i < n) offsets[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}
// compact id -> Cartesian cell, node word and int8 label per stored node of the state box
__global__ void k_compact_maps(const int32_t *index, const uint32_t *node, const uint32_t *wall, const int32_t *label,
                               long long cells, long long id_first, long long *cart, uint32_t *nodec, uint32_t *wallc,
                               int8_t *labelc) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    int i = index[c];
    if (i < 0) return;
    long long o = (long long)i - id_first;
    cart[o] = c;
    nodec[o] = node[c];
    wallc[o] = wall[c];
    labelc[o] = (int8_t)label[c];
}
template <typename T>
__global__ void k_init_sparse(const __grid_constant__ InitParams<T> p, const long long *cart, long long nstored) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstored) return;
    const Box &b = p.box;
    const long long c = cart[i];
    Coord co = coord_of(b, c);
    T ux, uy, uz, feq[Q];
    if (p.aa && p.label[c] == p.fluid_label) {
        // in-place storage: slot (q, x) of a fluid node holds what its first (even) step reads, the
        // PRE-STREAMED population feq_q of the source x - c_q (every source of a fluid node is stored)
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int x = co.x - cxq(q), y = co.y - cyq(q), z = co.z - czq(q);
            const bool in = x >= 0 && x < b.nx && y >= 0 && y < b.ny && z >= b.z0 && z < b.z1;
            const int g = in ? p.label[cell_of(b, x, y, z)] : 0;
            init_velocity<T>(p.case_rule, p.u_max, p.bc, p.plane_in, p.plane_out, b, g, in ? x : -1, y, z, ux, uy, uz);
            feq[q] = init_feq_q<T>(p.case_rule, q, T(1.0), ux, uy, uz);
        }
    } else {
        init_velocity<T>(p.case_rule, p.u_max, p.bc, p.plane_in, p.plane_out, b, p.label[c], co.x, co.y, co.z, ux, uy, uz);
        if (p.case_rule == LBM_CASE_LDC) {
            feq_all_ldc_init<T>(T(1.0), ux, uy, uz, feq);
        } else {
            const T r3 = T(1.0) / T(3.0), r18 = T(1.0) / T(18.0), r36 = T(1.0) / T(36.0);
#pragma unroll
            for (int q = 0; q < Q; q++) feq[q] = feq_lit<T>(q, r3, r18, r36, ux, uy, uz);
        }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) {
        p.fa[(long long)q * p.qstride + i] = feq[q];
        if (p.fb != p.fa) p.fb[(long long)q * p.qstride + i] = feq[q];
    }
    p.rho[i] = T(0), p.ux[i] = T(0), p.uy[i] = T(0), p.uz[i] = T(0);
}
// in-place sparse storage seen as the reference's d_scr: entry (q, s) = what the fluid node x = s + c_q pulls next.
// Before an even step that is a[q][x]; before an odd step a[opp q][s] when s is fluid, else the link's own slot a[q][x].
// One thread per cell of the owned planes; `sid` is the storage's own numbering, `index` the reference's.
template <typename T>
__global__ void k_gather_pops_sparse_aa(const T *a, long long qstride, const int32_t *label, const int32_t *index,
                                        const int32_t *sid, Box b, int fluid_label, long long first, long long c0,
                                        long long c1, long long n, int odd, T *out) {
    const long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int ci = index[c];
    if (ci < 0) return;
    const long long t = (long long)ci - first;
    const Coord co = coord_of(b, c);
    const bool s_fluid = label[c] == fluid_label;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int x = co.x + cxq(q), y = co.y + cyq(q), z = co.z + czq(q);
        T v = T(0);
        if (x >= 0 && x < b.nx && y >= 0 && y < b.ny && z >= b.z0 && z < b.z1) {
            const long long cx = cell_of(b, x, y, z);
            if (label[cx] == fluid_label) {
                const long long xi = sid[cx];
                v = !odd ? a[(long long)q * qstride + xi] : (s_fluid ? a[(long long)oppq(q) * qstride + sid[c]] : a[(long long)q * qstride + xi]);
            }
        }
        out[(long long)q * n + t] = v;
    }
}
// ---- the in-place sparse storage's own numbering -----------------------------------------------------------
// Only fluid nodes carry state there (boundary links live in the fluid node's own slot), so only they get
// an id -- plus the single solid cells that sit between two fluid cells of a row ("F S F"): with those
// filled in, the ids of the up-to-three sources a node has in a neighbouring row (x-1, x, x+1) are
// always consecutive, i.e. ONE id per neighbouring row describes them.  Ids run over z, y, x like the
// reference's index_transform, so a z-plane is one contiguous id range.
__global__ void k_span_flags(const int32_t *label, Box b, int fluid_label, int32_t *keep) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.cells()) return;
    const int x = (int)(c % b.px);
    int k = 0;
    if (x < b.nx) {
        if (label[c] == fluid_label) k = 1;
        else if (x > 0 && x < b.nx - 1 && label[c - 1] == fluid_label && label[c + 1] == fluid_label) k = 1;
    }
    keep[c] = k;
}
// per aligned chunk of 32 ids: which lanes are fluid, and whether any of them has a link that is not a wall
__global__ void k_chunk_meta(const uint32_t *nodec, long long ns, uint2 *meta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t node = i < ns ? nodec[i] : NODE_SKIP;
    const unsigned fl = __ballot_sync(0xffffffffu, !(node & NODE_SKIP));
    const unsigned sp = __ballot_sync(0xffffffffu, !(node & NODE_SKIP) && (node & NODE_LINKS) && !(node & NODE_WALLS_ONLY));
    if ((threadIdx.x & 31) == 0 && i < ns) meta[i >> 5] = make_uint2(fl, sp ? 1u : 0u);
}
// reductions over compact arrays (see k_reduce_fields)
template <typename T>
__global__ void k_reduce_fields_sparse(const T *ux, const T *uy, const T *uz, const int8_t *labelc, const long long *cart,
                                       Box b, long long i0, long long i1, int kind, int fluid_label, int case_rule,
                                       double *out) {
    long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < i1) {
        int g = labelc[i];
        T s = ux[i] * ux[i] + uy[i] * uy[i] + uz[i] * uz[i];
        if (kind == 0) {
            if (g == fluid_label) v = (double)(T)sqrt((double)s);
        } else {
            Coord p = coord_of(b, cart[i]);
            bool ok = case_rule == LBM_CASE_GEO_OPENINGS ? (g == 4) : (g >= 4);
            if (case_rule == LBM_CASE_LDC) ok = g == fluid_label;
            bool trimmed = p.x >= 1 && p.x <= b.nx - 2 && p.y >= 2 && p.y <= b.ny - 3 && p.z >= 1 && p.z <= b.nz - 2;
            if (ok && trimmed) v = (double)s;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += ws[k];
        if (t != 0.0) atomicAdd(out, t);
    }
}
// halo planes in compact storage: a plane's stored nodes are one contiguous id range
template <typename T>
__global__ void k_halo_pack_sparse(const T *f, long long qstride, long long i0, long long n, int side, T *buf,
                                   long long bufstride) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 5; k++) buf[(long long)k * bufstride + i] = f[(long long)halo_q(side, k) * qstride + i0 + i];
}
template <typename T>
__global__ void k_halo_unpack_sparse(T *f, long long qstride, const int8_t *labelc, int fluid_label, long long i0,
                                     long long n, int side, const T *buf, long long bufstride) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (labelc[i0 + i] != fluid_label) return;
#pragma unroll
    for (int k = 0; k < 5; k++) f[(long long)halo_q(1 - side, k) * qstride + i0 + i] = buf[(long long)k * bufstride + i];
}

inline unsigned nblocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

// ---------------------------------------------------------------- launchers
cudaError_t launch_make_flag_pos(uint8_t *flag, Box ext, cudaStream_t s) {
    k_make_flag_pos<<<nblocks(ext.cells(), 256), 256, 0, s>>>(flag, ext);
    return cudaGetLastError();
}
cudaError_t launch_labels(const uint8_t *flag, int32_t *label, Box ext, GeoRules r, cudaStream_t s) {
    k_labels<<<nblocks(ext.cells(), 256), 256, 0, s>>>(flag, label, ext, r);
    return cudaGetLastError();
}
cudaError_t launch_mark(int32_t *label, Box ext, GeoRules r, cudaStream_t s) {
    if (r.mark_sources == 0) return cudaSuccess;
    k_mark<<<nblocks(ext.cells(), 256), 256, 0, s>>>(label, ext, r);
    return cudaGetLastError();
}
size_t compact_scratch_ints(long long cells) {
    long long ntiles = (cells + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(ntiles + 2 * ntiles + 4);  // int32 counts + int64 offsets
}
cudaError_t launch_compact(const int32_t *label, int32_t *index, long long cells, int px, int nx, int all,
                           long long base, int32_t *scratch, size_t scratch_ints, long long *total_out_dev,
                           long long plane, long long *plane_first_dev, cudaStream_t s) {
    StoredRule sr{px, nx, all};
    long long ntiles = (cells + SCAN_TILE - 1) / SCAN_TILE;
    if ((size_t)(3 * ntiles + 4) > scratch_ints) return cudaErrorInvalidValue;
    int32_t *counts = scratch;
    // 8-byte aligned offsets behind the counts
    long long *offsets = reinterpret_cast<long long *>(scratch + ((ntiles + 1) & ~1LL));
    k_tile_counts<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(label, cells, sr, counts);
    k_scan_tiles<<<1, SCAN_BLOCK, 0, s>>>(counts, offsets, (int)ntiles, total_out_dev);
    k_scatter_index<<<(unsigned)ntiles, SCAN_BLOCK, 0, s>>>(label, index, cells, sr, base, offsets, plane, plane_first_dev);
    return cudaGetLastError();
}
// sum over cells of (label + 2) * odd multiplier of the cell id: commutative, so the atomics may land in any order
__global__ void k_label_hash(const int32_t *label, long long cells, unsigned long long *out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (c < cells) v = (unsigned long long)(label[c] + 2) * (((unsigned long long)c * 0x9E3779B97F4A7C15ull) | 1ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
cudaError_t launch_label_hash(const int32_t *label, long long cells, unsigned long long *out_dev, cudaStream_t s) {
    k_label_hash<<<nblocks(cells, 256), 256, 0, s>>>(label, cells, out_dev);
    return cudaGetLastError();
}
cudaError_t launch_count_stored(const int32_t *label, long long cells, int px, int nx, int all, long long *out_dev,
                                cudaStream_t s) {
    StoredRule sr{px, nx, all};
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(long long), s);
    if (e != cudaSuccess) return e;
    k_count_stored<<<nblocks(cells, 256), 256, 0, s>>>(label, cells, sr, out_dev);
    return cudaGetLastError();
}
cudaError_t launch_node_words(const int32_t *label, uint32_t *node, uint32_t *wall, uint8_t *seg, int8_t *label8, Box box,
                              int own_z0, int own_z1, int fluid_label, const BcEntry *bc, long long *nfluid_dev,
                              cudaStream_t s) {
    BcTable t;
    for (int i = 0; i < LBM_MAX_BC; i++) t.e[i] = bc[i];
    cudaError_t e = cudaMemsetAsync(nfluid_dev, 0, sizeof(long long), s);
    if (e != cudaSuccess) return e;
    k_node_words<<<nblocks(box.cells(), 256), 256, 0, s>>>(label, node, wall, seg, label8, box, own_z0, own_z1, fluid_label,
                                                          t, nfluid_dev);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_init(const InitParams<T> &p, cudaStream_t s) {
    k_init<T><<<nblocks(p.box.cells(), 128), 128, 0, s>>>(p);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_gather_fields(const T *rho, const T *ux, const T *uy, const T *uz, const int32_t *label,
                                 const int32_t *index, const int32_t *sid, Box box, int own_z0, int own_z1, int fluid_label,
                                 long long first, T *orho, T *oux, T *ouy, T *ouz, cudaStream_t s) {
    long long c0 = (long long)(own_z0 - box.z0) * box.plane, c1 = (long long)(own_z1 - box.z0) * box.plane;
    if (c1 <= c0) return cudaSuccess;
    k_gather_fields<T><<<nblocks(c1 - c0, 256), 256, 0, s>>>(rho, ux, uy, uz, label, index, sid, box, c0, c1, fluid_label,
                                                            first, orho, oux, ouy, ouz);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_gather_pops(const T *f, long long qstride, const int32_t *index, Box box, int own_z0, int own_z1,
                               long long first, long long count, int layout, T *out, cudaStream_t s) {
    long long c0 = (long long)(own_z0 - box.z0) * box.plane, c1 = (long long)(own_z1 - box.z0) * box.plane;
    if (c1 <= c0) return cudaSuccess;
    k_gather_pops<T><<<nblocks(c1 - c0, 256), 256, 0, s>>>(f, qstride, index, box, c0, c1, first, count, layout, out);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_reduce_fields(const T *ux, const T *uy, const T *uz, const int32_t *label, Box box, int own_z0,
                                 int own_z1, int kind, int fluid_label, int case_rule, double *out_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), s);
    if (e != cudaSuccess) return e;
    long long c0 = (long long)(own_z0 - box.z0) * box.plane, c1 = (long long)(own_z1 - box.z0) * box.plane;
    if (c1 <= c0) return cudaSuccess;
    k_reduce_fields<T><<<nblocks(c1 - c0, 256), 256, 0, s>>>(ux, uy, uz, label, box, c0, c1, kind, fluid_label, case_rule,
                                                            out_dev);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_halo_pack(const T *f, long long qstride, Box box, int zl, int side, T *buf, cudaStream_t s) {
    k_halo_pack<T><<<nblocks(box.plane, 256), 256, 0, s>>>(f, qstride, box, zl, side, buf);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_halo_unpack(T *f, long long qstride, const int8_t *label8, int fluid_label, Box box, int zl, int side,
                               const T *buf, cudaStream_t s) {
    k_halo_unpack<T><<<nblocks(box.plane, 256), 256, 0, s>>>(f, qstride, label8, fluid_label, box, zl, side, buf);
    return cudaGetLastError();
}

// ---- mailboxes of the dense in-place storage (StepParams::mail): copy the slots of the 5 entering directions
// between the population buffer and a side's mailbox
template <typename T>
__global__ void k_mail_copy(T *a, long long qs, T *mail, long long ms, long long G, long long face_c0, long long halo_c0,
                            long long plane, int side, int dir) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
#pragma unroll
    for (int q = 1; q < Q; q++) {
        if (czq(q) != (side == 0 ? 1 : -1)) continue;
        T *mA = mail + (long long)kslot(q) * ms + G + i, *mB = mail + (long long)(5 + kslot(q)) * ms + G + i;
        T *pA = a + (long long)oppq(q) * qs + halo_c0 + i, *pB = a + (long long)q * qs + face_c0 + i;
        if (dir == 0) *mA = *pA, *mB = *pB;
        else *pA = *mA, *pB = *mB;
    }
}
template <typename T>
cudaError_t launch_mail_copy(T *a, long long qstride, T *mail, long long ms, long long G, long long face_c0, long long halo_c0,
                             long long plane, int side, int dir, cudaStream_t s) {
    k_mail_copy<T><<<nblocks(plane, 256), 256, 0, s>>>(a, qstride, mail, ms, G, face_c0, halo_c0, plane, side, dir);
    return cudaGetLastError();
}
template cudaError_t launch_mail_copy<float>(float *, long long, float *, long long, long long, long long, long long, long long, int, int, cudaStream_t);
template cudaError_t launch_mail_copy<double>(double *, long long, double *, long long, long long, long long, long long, long long, int, int, cudaStream_t);

// Staged mailboxes (transport without peer mapping: NCCL send/recv or any other copy).  The step kernel stores the
// leaving populations into a LOCAL staging buffer laid out like the neighbour's mailbox, prefilled with all-one
// bits; after the transfer the receiver takes over exactly the elements the sender wrote -- the others belong to
// the receiver's own boundary slots.  (All-one bits are a NaN no arithmetic produces.)
template <typename T>
__global__ void k_mail_merge(T *mail, const T *in, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = in[i];
    bool written;
    if (sizeof(T) == 8) written = __double_as_longlong((double)v) != -1LL;
    else written = __float_as_int((float)v) != -1;
    if (written) mail[i] = v;
}
template <typename T>
cudaError_t launch_mail_merge(T *mail, const T *in, long long n, cudaStream_t s) {
    k_mail_merge<T><<<nblocks(n, 256), 256, 0, s>>>(mail, in, n);
    return cudaGetLastError();
}
template cudaError_t launch_mail_merge<float>(float *, const float *, long long, cudaStream_t);
template cudaError_t launch_mail_merge<double>(double *, const double *, long long, cudaStream_t);

// ---- neighbour handshake between z-slabs that live in different processes (one process per GPU)
// A slab may start the face launches of step t+1 only after both neighbours finished the face launches of
// step t (they read the halo plane this slab is about to overwrite, and wrote the one it is about to read).
// Each slab owns a small sync block in its own device memory: [0] steps finished by the LOW neighbour,
// [1] by the HIGH neighbour -- written remotely by that neighbour (peer memory over NVLink), polled locally
// -- [2] is set when a wait gave up.  Two single-thread kernels per step; the waiting one sleeps between
// polls and gives up after `timeout_ns` instead of hanging the device when a neighbour died.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_slab_wait(unsigned long long *sync, unsigned long long need_lo, unsigned long long need_hi,
                            unsigned long long timeout_ns) {
    const unsigned long long t0 = global_ns();
    for (int side = 0; side < 2; side++) {
        const unsigned long long need = side ? need_hi : need_lo;
        if (need == 0) continue;
        while (ld_acquire_sys(sync + side) < need) {
            if (global_ns() - t0 > timeout_ns) {
                sync[2] = 1ull;
                return;
            }
            __nanosleep(100);
        }
    }
}
__global__ void k_slab_signal(unsigned long long *peer_lo, unsigned long long *peer_hi, unsigned long long value) {
    __threadfence_system();  // the face launches before this kernel are complete; order this thread's store after them
    if (peer_lo) st_release_sys(peer_lo, value);
    if (peer_hi) st_release_sys(peer_hi, value);
}
cudaError_t launch_slab_wait(unsigned long long *sync, unsigned long long need_lo, unsigned long long need_hi,
                             unsigned long long timeout_ns, cudaStream_t s) {
    k_slab_wait<<<1, 1, 0, s>>>(sync, need_lo, need_hi, timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_slab_signal(unsigned long long *peer_lo, unsigned long long *peer_hi, unsigned long long value,
                               cudaStream_t s) {
    k_slab_signal<<<1, 1, 0, s>>>(peer_lo, peer_hi, value);
    return cudaGetLastError();
}

// ---- sparse storage
cudaError_t launch_build_segments(const uint32_t *nodec, const long long *cart, const int32_t *index, Box box,
                                  int own_zl0, long long id0, long long id1, long long id_first, int32_t *counts,
                                  long long *offsets, long long *nseg_dev, int32_t *rec, long long *plane_seg,
                                  cudaStream_t s) {
    const long long base_id = id0 & ~31LL;
    const long long nchunks = (id1 - base_id + 31) >> 5;
    if (nchunks <= 0) return cudaSuccess;
    if (!rec) {  // pass 1: records per chunk and scan
        k_seg_count<<<nblocks(nchunks * 32, 256), 256, 0, s>>>(nodec, cart, box, base_id, id0, id1, counts);
        k_scan_i32<<<1, SCAN_BLOCK, 0, s>>>(counts, offsets, nchunks, nseg_dev);
    } else {     // pass 2: fill the records
        k_seg_fill<<<nblocks(nchunks * 32, 256), 256, 0, s>>>(nodec, cart, index, box, own_zl0, base_id, id0, id1, id_first,
                                                              offsets, rec, plane_seg);
    }
    return cudaGetLastError();
}
// in-place storage: the same two passes with the row-break rule and row ids (k_seg_*_rows)
cudaError_t launch_build_segments_rows(const uint32_t *nodec, const long long *cart, const int32_t *sid, const int32_t *label,
                                       int fluid_label, Box box, int own_zl0, long long id0, long long id1, int32_t *counts,
                                       long long *offsets, long long *nseg_dev, int32_t *rec, long long *plane_seg,
                                       cudaStream_t s) {
    const long long base_id = id0 & ~31LL;
    const long long nchunks = (id1 - base_id + 31) >> 5;
    if (nchunks <= 0) return cudaSuccess;
    RowInfo rows{sid, label, fluid_label};
    if (!rec) {
        k_seg_count_rows<<<nblocks(nchunks * 32, 256), 256, 0, s>>>(nodec, cart, rows, box, base_id, id0, id1, counts);
        k_scan_i32<<<1, SCAN_BLOCK, 0, s>>>(counts, offsets, nchunks, nseg_dev);
    } else {
        k_seg_fill_rows<<<nblocks(nchunks * 32, 256), 256, 0, s>>>(nodec, cart, rows, box, own_zl0, base_id, id0, id1, offsets,
                                                                   rec, plane_seg);
    }
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_bc_links(const uint32_t *nodec, const uint32_t *wallc, const long long *cart, const int8_t *label8, Box box,
                            const BcEntry *bc, const T *plane_in, const T *plane_out, long long i0, long long i1,
                            int *total_dev, int32_t *bcslot, BcLink<T> *links, cudaStream_t s) {
    if (i1 <= i0) return cudaSuccess;
    BcTable t;
    for (int i = 0; i < LBM_MAX_BC; i++) t.e[i] = bc[i];
    k_bc_links<T><<<nblocks(i1 - i0, 128), 128, 0, s>>>(nodec, wallc, cart, label8, box, t, plane_in, plane_out, i0, i1, total_dev,
                                                        bcslot, links);
    return cudaGetLastError();
}
cudaError_t launch_rec_links(const int32_t *rec, const uint32_t *nodec, long long nseg, uint32_t *links, cudaStream_t s) {
    if (nseg <= 0) return cudaSuccess;
    k_rec_links<<<nblocks(nseg * 32, 256), 256, 0, s>>>(rec, nodec, nseg, links);
    return cudaGetLastError();
}
cudaError_t launch_span_flags(const int32_t *label, Box box, int fluid_label, int32_t *keep, cudaStream_t s) {
    k_span_flags<<<nblocks(box.cells(), 256), 256, 0, s>>>(label, box, fluid_label, keep);
    return cudaGetLastError();
}
cudaError_t launch_chunk_meta(const uint32_t *nodec, long long ns, uint2 *meta, cudaStream_t s) {
    if (ns <= 0) return cudaSuccess;
    k_chunk_meta<<<nblocks(((ns + 31) / 32) * 32, 256), 256, 0, s>>>(nodec, ns, meta);
    return cudaGetLastError();
}
cudaError_t launch_compact_maps(const int32_t *index, const uint32_t *node, const uint32_t *wall, const int32_t *label,
                                long long cells, long long id_first, long long *cart, uint32_t *nodec, uint32_t *wallc,
                                int8_t *labelc, cudaStream_t s) {
    k_compact_maps<<<nblocks(cells, 256), 256, 0, s>>>(index, node, wall, label, cells, id_first, cart, nodec, wallc, labelc);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_init_sparse(const InitParams<T> &p, const long long *cart, long long nstored, cudaStream_t s) {
    if (nstored <= 0) return cudaSuccess;
    k_init_sparse<T><<<nblocks(nstored, 128), 128, 0, s>>>(p, cart, nstored);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_gather_pops_sparse_aa(const T *a, long long qstride, const int32_t *label, const int32_t *index,
                                         const int32_t *sid, Box box, int own_z0, int own_z1, int fluid_label, long long first,
                                         long long n, int odd, T *out, cudaStream_t s) {
    const long long c0 = (long long)(own_z0 - box.z0) * box.plane, c1 = (long long)(own_z1 - box.z0) * box.plane;
    if (n <= 0 || c1 <= c0) return cudaSuccess;
    k_gather_pops_sparse_aa<T><<<nblocks(c1 - c0, 128), 128, 0, s>>>(a, qstride, label, index, sid, box, fluid_label, first, c0,
                                                                     c1, n, odd, out);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_reduce_fields_sparse(const T *ux, const T *uy, const T *uz, const int8_t *labelc, const long long *cart,
                                        Box box, long long i0, long long i1, int kind, int fluid_label, int case_rule,
                                        double *out_dev, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), s);
    if (e != cudaSuccess || i1 <= i0) return e;
    k_reduce_fields_sparse<T><<<nblocks(i1 - i0, 256), 256, 0, s>>>(ux, uy, uz, labelc, cart, box, i0, i1, kind,
                                                                   fluid_label, case_rule, out_dev);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_halo_pack_sparse(const T *f, long long qstride, long long i0, long long n, int side, T *buf,
                                    long long bufstride, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_halo_pack_sparse<T><<<nblocks(n, 256), 256, 0, s>>>(f, qstride, i0, n, side, buf, bufstride);
    return cudaGetLastError();
}
template <typename T>
cudaError_t launch_halo_unpack_sparse(T *f, long long qstride, const int8_t *labelc, int fluid_label, long long i0,
                                      long long n, int side, const T *buf, long long bufstride, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_halo_unpack_sparse<T><<<nblocks(n, 256), 256, 0, s>>>(f, qstride, labelc, fluid_label, i0, n, side, buf, bufstride);
    return cudaGetLastError();
}

#define LBM_INST(T)                                                                                                     \
    template cudaError_t launch_init<T>(const InitParams<T> &, cudaStream_t);                                           \
    template cudaError_t launch_gather_fields<T>(const T *, const T *, const T *, const T *, const int32_t *,           \
                                                 const int32_t *, const int32_t *, Box, int, int, int, long long, T *,  \
                                                 T *, T *, T *, cudaStream_t);                                          \
    template cudaError_t launch_gather_pops<T>(const T *, long long, const int32_t *, Box, int, int, long long,         \
                                               long long, int, T *, cudaStream_t);                                           \
    template cudaError_t launch_reduce_fields<T>(const T *, const T *, const T *, const int32_t *, Box, int, int, int,  \
                                                 int, int, double *, cudaStream_t);                                     \
    template cudaError_t launch_halo_pack<T>(const T *, long long, Box, int, int, T *, cudaStream_t);                   \
    template cudaError_t launch_init_sparse<T>(const InitParams<T> &, const long long *, long long, cudaStream_t);      \
    template cudaError_t launch_bc_links<T>(const uint32_t *, const uint32_t *, const long long *, const int8_t *, Box, \
                                            const BcEntry *, const T *, const T *, long long, long long, int *,         \
                                            int32_t *, BcLink<T> *, cudaStream_t);                                      \
    template cudaError_t launch_gather_pops_sparse_aa<T>(const T *, long long, const int32_t *, const int32_t *,        \
                                                         const int32_t *, Box, int, int, int, long long, long long,     \
                                                         int, T *, cudaStream_t);                                       \
    template cudaError_t launch_reduce_fields_sparse<T>(const T *, const T *, const T *, const int8_t *,                \
                                                        const long long *, Box, long long, long long, int, int, int,   \
                                                        double *, cudaStream_t);                                        \
    template cudaError_t launch_halo_pack_sparse<T>(const T *, long long, long long, long long, int, T *, long long,   \
                                                    cudaStream_t);                                                      \
    template cudaError_t launch_halo_unpack_sparse<T>(T *, long long, const int8_t *, int, long long, long long, int,   \
                                                      const T *, long long, cudaStream_t);                              \
    template cudaError_t launch_halo_unpack<T>(T *, long long, const int8_t *, int, Box, int, int, const T *,           \
                                               cudaStream_t);
LBM_INST(float)
LBM_INST(double)
#undef LBM_INST

}  // namespace lbm
