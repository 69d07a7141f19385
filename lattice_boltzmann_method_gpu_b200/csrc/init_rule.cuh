// The reference's initial state, as a pure function of a cell: rho = 1, u = 0 except on the
// planes / labels each program seeds with its boundary velocity, f = feq(rho,u)
// (ldc.cu:504-580, Poiseulle.cu:273-382, bifurcation.cu:329-427, coronary.cu:277-350).
// Used by the initialisation kernel and, in the in-place (AA) storage, by the odd step to
// re-create the never-changing population of a "static" link.
#pragma once
#include "lattice.cuh"
#include "lbm_internal.h"

namespace lbm {

template <typename T>
__device__ __forceinline__ T parabola_at(const Box &b, T umax, int x, int z) {
    // pos.cu:301 -- squares of half-integers are exact, so a*a equals the reference's powf(a,2)
    T cx = T(b.nx - 1) / T(2.0), cz = T(b.nz - 1) / T(2.0), r = T(b.nx - 1) / T(2.0);
    T dx = T(x) - cx, dz = T(z) - cz;
    return umax * (T(1.0) - (dx * dx + dz * dz) / (r * r));
}

// prescribed boundary speed of BC entry `e` at boundary node (gx, ., gz) (global coordinates), before the
// pulsatile scale: a constant (ldc.cu:378, cor:717), the analytic parabola evaluated at the boundary node's own
// (i,k) (pos.cu:597; squares of half-integers are exact, so a*a equals the reference's powf(a,2)), or the
// bc.txt planes (bif:650,951)
template <typename T>
__device__ __forceinline__ T bc_speed_unscaled(const BcEntry &e, const Box &b, const T *plane_in, const T *plane_out, int gx,
                                               int gz) {
    if (e.source == LBM_SRC_CONST) return (T)e.value;
    if (e.source == LBM_SRC_PARABOLA) {
        T cx = T(b.nx - 1) / T(2.0), cz = T(b.nz - 1) / T(2.0), r = T(b.nx - 1) / T(2.0);
        T dx = T(gx) - cx, dz = T(gz) - cz;
        return (T)e.value * (T(1.0) - (dx * dx + dz * dz) / (r * r));
    }
    if (e.source == LBM_SRC_PLANE_INLET) return plane_in[gx + (long long)gz * b.nx];
    return plane_out[gx + (long long)gz * b.nx];
}

// initial velocity of the cell with label g at GLOBAL coordinates (x,y,z); zero for cells the
// reference does not store (label 0) and outside the box
template <typename T>
__device__ __forceinline__ void init_velocity(int case_rule, T u_max, const BcEntry *bc, const T *plane_in,
                                              const T *plane_out, const Box &b, int g, int x, int y, int z, T &ux, T &uy,
                                              T &uz) {
    ux = T(0.0), uy = T(0.0), uz = T(0.0);
    if (x < 0 || x >= b.nx || y < 0 || y >= b.ny || z < 0 || z >= b.nz) return;
    if (case_rule == LBM_CASE_LDC) {
        if (y == b.ny - 1 || y == b.ny - 2) uz = u_max;  // ldc.cu:523-531 (every node of both planes)
    } else if (case_rule == LBM_CASE_POISEUILLE) {
        if (g != 0 && (y <= 1 || y >= b.ny - 2)) uy = parabola_at<T>(b, u_max, x, z);  // pos:295-341
    } else if (case_rule == LBM_CASE_GEO_Y_INOUT) {
        if (g != 0 && y == 1) uy = plane_in[x + (long long)z * b.nx];  // bif:349-373
        if (g != 0 && y == b.ny - 2) uy = plane_out[x + (long long)z * b.nx];
    } else {
        if (g >= 2 && g < LBM_MAX_BC && g != 4 && bc[g].kind != LBM_BC_NONE) {  // cor:302-306
            T v = (T)bc[g].init_value;
            if (bc[g].vaxis == 0) ux = v;
            else if (bc[g].vaxis == 1) uy = v;
            else uz = v;
        }
    }
}

// one direction of the initial equilibrium, in the form the program uses (ldc: factored)
template <typename T>
__device__ __forceinline__ T init_feq_q(int case_rule, int q, T rho, T ux, T uy, T uz) {
    if (case_rule == LBM_CASE_LDC) {
        T feq[Q];
        feq_all_ldc_init<T>(rho, ux, uy, uz, feq);
        T v = feq[0];
#pragma unroll
        for (int k = 1; k < Q; k++) v = k == q ? feq[k] : v;
        return v;
    }
    return feq_lit<T>(q, rho / T(3.0), rho / T(18.0), rho / T(36.0), ux, uy, uz);
}

}  // namespace lbm
