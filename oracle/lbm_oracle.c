/*
 * lbm_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C, CPU restatement of the hot path of
 * Xinhuan-Imperial/Lattice-Boltzmann-Method-GPU (D3Q19 BGK), written from the
 * semantics of the four reference programs (stored-node formulation: every
 * stored node -- fluid, wall, inlet/outlet, outer-wall-neighbour -- carries 19
 * populations in two buffers; `update` then `boundary_stream` then swap).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (lattice_boltzmann_method_gpu_b200/, liblbm_b200.so) never does.
 *
 * Citations are file:line into the read-only reference tree:
 *   ldc = Lid_driven_cavity/ldc.cu      pos = Poiseulle_flow/Poiseulle.cu
 *   bif = bifurcation/bifurcation.cu    cor = coronary_cfd/coronary.cu
 *
 * Parity pin status: pinned by (a) the thesis' lattice count 65820 for the
 * shipped bifurcation geo.txt (tests/test_oracle_golden.py), (b) the analytic
 * Poiseuille profile, and (c) on a GPU box, outputs of the reference programs
 * themselves compiled unmodified into oracle/_ref (tests/test_reference_outputs.py).
 * The pulsatile-inlet extension (pulse_amp != 0) has no reference code:
 * parity unpinned for that one feature.
 *
 * Differences from the reference that do not affect any fluid node:
 *   - arrays are addressed in plain Cartesian order [z][y][x]; the reference's
 *     8x8xBZ tiled index (ldc:71, bif:54) is an internal detail of its kernels;
 *   - boundary_stream gathers from a snapshot of dst instead of racing
 *     in place (race note);
 *   - ldc: walls bounce on src before fluid nodes pull (ldc:75-202 executed
 *     for all walls first), the defined LDC semantics;
 *   - reads that the reference would make out of bounds / through index -1 are
 *     replaced by 0; rho/u of never-updated stored nodes are 0 (the reference
 *     leaves them uninitialised, ldc:635-638).
 *
 * Compiled twice (REAL=float / REAL=double) by oracle/Makefile with
 * -ffp-contract=off so evaluation order is exactly the one written here.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUF
#define SUF _f32
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

#define R(x) ((REAL)(x))

enum { CASE_LDC = 0, CASE_POS = 1, CASE_BIF = 2, CASE_COR = 3 };

/* Lattice, derived from the pull offsets ldc:207-313 and moment sums ldc:320-322. */
static const int CX[19] = {0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, 0, 0, 0, 0};
static const int CY[19] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
static const int CZ[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, 1, -1, -1};
/* swap list ldc:184-201 */
static const int OPP[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

/* ------------------------------------------------------------------------ */
/* Integer pre-processing (precision independent; exported once, no suffix). */
/* ------------------------------------------------------------------------ */
#ifdef ORC_EXPORT_INT

static inline int imin(int a, int b) { return a < b ? a : b; }
#define CID(x, y, z) ((size_t)(x) + (size_t)nx * ((size_t)(y) + (size_t)ny * (size_t)(z)))

/* ldc:468-502 : labels by layer. 0 ghost, 1 wall, 2 lid, 3 fluid. */
void orc_geo_pre_ldc(int nx, int ny, int nz, int32_t *geo) {
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) geo[CID(x, y, z)] = 0;
    for (int z = 1; z < nz - 1; z++)
        for (int y = 1; y < ny - 1; y++)
            for (int x = 1; x < nx - 1; x++) geo[CID(x, y, z)] = 1;
    for (int z = 2; z < nz - 2; z++)
        for (int y = 2; y < ny - 2; y++)
            for (int x = 2; x < nx - 2; x++) geo[CID(x, y, z)] = 3;
    for (int z = 1; z < nz - 1; z++)
        for (int x = 1; x < nx - 1; x++) geo[CID(x, ny - 2, z)] = 2;
}

/* min over the six axis neighbours of the binary flag (bif:82-86). */
static int min6(const int32_t *flag, int nx, int ny, int nz, int x, int y, int z) {
    (void)nz;
    int mx = imin(flag[CID(x + 1, y, z)], flag[CID(x - 1, y, z)]);
    int my = imin(flag[CID(x, y - 1, z)], flag[CID(x, y + 1, z)]);
    int mz = imin(flag[CID(x, y, z - 1)], flag[CID(x, y, z + 1)]);
    return imin(imin(mx, my), mz);
}

/* pos:138-254, bif:123-239, cor:144-260 : outer-wall-neighbour marking.
 * Every source node inside [1,N-2]^3 turns each label-0 node among its 18
 * neighbours into -1.  srcmask bit L set <=> label L is a source. */
static void mark_outer(int nx, int ny, int nz, int32_t *geo, unsigned srcmask) {
    for (int z = 1; z < nz - 1; z++)
        for (int y = 1; y < ny - 1; y++)
            for (int x = 1; x < nx - 1; x++) {
                int g = geo[CID(x, y, z)];
                if (g < 0 || g > 31 || !((srcmask >> g) & 1u)) continue;
                for (int q = 1; q < 19; q++) {
                    size_t n = CID(x + CX[q], y + CY[q], z + CZ[q]);
                    if (geo[n] == 0) geo[n] = -1;
                }
            }
}

/* pos:52-254 : circular pipe along y.  Binary disc in float (pos:80-91),
 * label = flag + 3*min6 on y in [2,NY-3] (pos:94-108), inlet y=1: flag+min4
 * (pos:110-120), outlet y=NY-2: flag+2*min4 (pos:122-134), marking sources
 * {1,2,3} (pos:142). */
void orc_geo_pre_pos(int nx, int ny, int nz, int32_t *geo) {
    size_t n = (size_t)nx * ny * nz;
    int32_t *flag = (int32_t *)calloc(n, sizeof(int32_t));
    float radius = (nx - 1) / 2.0f, cx = (nx - 1) / 2.0f, cz = (nz - 1) / 2.0f;
    memset(geo, 0, n * sizeof(int32_t));
    for (int x = 0; x < nx; x++)
        for (int y = 1; y < ny - 1; y++)
            for (int z = 0; z < nz; z++) {
                float dist = sqrtf(powf(x - cx, 2) + powf(z - cz, 2));
                if (dist <= radius) {
                    flag[CID(x, y, z)] = 1;
                    geo[CID(x, y, z)] = 1;
                }
            }
    for (int t = 0; t < 3; t++)
        for (int x = 1; x < nx - 1; x++)
            for (int y = 2; y < ny - 2; y++)
                for (int z = 1; z < nz - 1; z++) geo[CID(x, y, z)] += min6(flag, nx, ny, nz, x, y, z);
    for (int pass = 0; pass < 2; pass++) {
        int y = pass == 0 ? 1 : ny - 2, reps = pass == 0 ? 1 : 2;
        for (int t = 0; t < reps; t++)
            for (int x = 1; x < nx - 1; x++)
                for (int z = 1; z < nz - 1; z++) {
                    int mx = imin(flag[CID(x + 1, y, z)], flag[CID(x - 1, y, z)]);
                    int mz = imin(flag[CID(x, y, z - 1)], flag[CID(x, y, z + 1)]);
                    geo[CID(x, y, z)] += imin(mx, mz);
                }
    }
    free(flag);
    mark_outer(nx, ny, nz, geo, (1u << 1) | (1u << 2) | (1u << 3));
}

/* bif:36-239 : geometry from a binary voxel field (geo.txt order z,y,x = our
 * Cartesian order).  Planes y=0,NY-1 zeroed on the interior x,z range
 * (bif:63-73); label = flag + 3*min6 on y in [2,NY-3] (bif:77-91); inlet plane
 * y=1 copied from y=2 (1->1, 4->2, else 0; bif:94-104); outlet plane y=NY-2
 * copied from y=NY-3 (1->1, 4->3, else 0; bif:107-119); marking sources {1}. */
void orc_geo_pre_bif(int nx, int ny, int nz, const int32_t *flag, int32_t *geo) {
    size_t n = (size_t)nx * ny * nz;
    memcpy(geo, flag, n * sizeof(int32_t));
    for (int x = 1; x < nx - 1; x++)
        for (int z = 1; z < nz - 1; z++) {
            geo[CID(x, 0, z)] = 0;
            geo[CID(x, ny - 1, z)] = 0;
        }
    for (int t = 0; t < 3; t++)
        for (int x = 1; x < nx - 1; x++)
            for (int y = 2; y < ny - 2; y++)
                for (int z = 1; z < nz - 1; z++) geo[CID(x, y, z)] += min6(flag, nx, ny, nz, x, y, z);
    for (int x = 1; x < nx - 1; x++)
        for (int z = 1; z < nz - 1; z++) {
            int g2 = geo[CID(x, 2, z)];
            geo[CID(x, 1, z)] = g2 == 1 ? 1 : (g2 == 4 ? 2 : 0);
        }
    for (int x = 1; x < nx - 1; x++)
        for (int z = 1; z < nz - 1; z++) {
            int g2 = geo[CID(x, ny - 3, z)];
            geo[CID(x, ny - 2, z)] = g2 == 1 ? 1 : (g2 == 4 ? 3 : 0);
        }
    mark_outer(nx, ny, nz, geo, 1u << 1);
}

/* cor:31-260 : label = flag + 3*min6 on the whole interior box (cor:60-74),
 * then a list of "opening" planes, each adding reps * in-plane-min4 inside a
 * window (cor:77-141).  rules[r] = {axis, coord, lo_a, hi_a, lo_b, hi_b, reps}
 * where (a,b) are the two in-plane axes in xyz order and the ranges are
 * inclusive.  The reference's own list is
 *   {0,3,1,NY-2,1,NZ-2,1} {0,272,1,NY-2,1,NZ-2,2} {2,185,217,236,113,137,4}
 *   {2,191,160,205,159,199,5} {2,204,1,NX-2,1,NY-2,6}.
 * The flag field is Cartesian [z][y][x]; the reference's geo.txt for this case
 * is stored y-fastest (cor:45-56) -- transposing is the reader's job. */
void orc_geo_pre_cor(int nx, int ny, int nz, const int32_t *flag, int nrules, const int32_t *rules,
                     int32_t *geo) {
    size_t n = (size_t)nx * ny * nz;
    memcpy(geo, flag, n * sizeof(int32_t));
    for (int t = 0; t < 3; t++)
        for (int x = 1; x < nx - 1; x++)
            for (int y = 1; y < ny - 1; y++)
                for (int z = 1; z < nz - 1; z++) geo[CID(x, y, z)] += min6(flag, nx, ny, nz, x, y, z);
    for (int r = 0; r < nrules; r++) {
        const int32_t *ru = rules + 7 * r;
        int axis = ru[0], c = ru[1];
        for (int t = 0; t < ru[6]; t++)
            for (int a = ru[2]; a <= ru[3]; a++)
                for (int b = ru[4]; b <= ru[5]; b++) {
                    int x, y, z, m;
                    if (axis == 0) {
                        x = c, y = a, z = b;
                        m = imin(imin(flag[CID(x, y - 1, z)], flag[CID(x, y + 1, z)]),
                                 imin(flag[CID(x, y, z - 1)], flag[CID(x, y, z + 1)]));
                    } else if (axis == 1) {
                        x = a, y = c, z = b;
                        m = imin(imin(flag[CID(x - 1, y, z)], flag[CID(x + 1, y, z)]),
                                 imin(flag[CID(x, y, z - 1)], flag[CID(x, y, z + 1)]));
                    } else {
                        x = a, y = b, z = c;
                        m = imin(imin(flag[CID(x, y - 1, z)], flag[CID(x, y + 1, z)]),
                                 imin(flag[CID(x - 1, y, z)], flag[CID(x + 1, y, z)]));
                    }
                    geo[CID(x, y, z)] += m;
                }
    }
    mark_outer(nx, ny, nz, geo, 1u << 1);
}

/* pos:257-271, bif:241-252, cor:262-273 : running count over z,y,x of geo!=0. */
int orc_index_transform(int nx, int ny, int nz, const int32_t *geo, int32_t *index) {
    int nlat = 0;
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                size_t c = CID(x, y, z);
                if (geo[c] != 0) index[c] = nlat++;
                else index[c] = -1;
            }
    return nlat;
}

/* ldc stores every node of the box (ldc:54): index = Cartesian id. */
int orc_index_dense(int nx, int ny, int nz, int32_t *index) {
    size_t n = (size_t)nx * ny * nz;
    for (size_t c = 0; c < n; c++) index[c] = (int32_t)c;
    return (int)n;
}

/* geo.txt reader, bif:50-61 ("%d " tokens, x fastest then y then z) or the
 * cor order (cor:45-56, y fastest then x then z) when yfast != 0.  Output is
 * always Cartesian [z][y][x].  Returns number of tokens read. */
long orc_read_geo_file(const char *path, int nx, int ny, int nz, int yfast, int32_t *flag) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    long cnt = 0;
    int tmp;
    for (int z = 0; z < nz; z++) {
        if (!yfast) {
            for (int y = 0; y < ny; y++)
                for (int x = 0; x < nx; x++) {
                    if (fscanf(f, "%d ", &tmp) != 1) tmp = 0; else cnt++;
                    flag[CID(x, y, z)] = tmp;
                }
        } else {
            for (int x = 0; x < nx; x++)
                for (int y = 0; y < ny; y++) {
                    if (fscanf(f, "%d ", &tmp) != 1) tmp = 0; else cnt++;
                    flag[CID(x, y, z)] = tmp;
                }
        }
    }
    fclose(f);
    return cnt;
}

/* bif:255-327 : bc.txt -> inlet u_y on (x,z) where geo(x,1,z)==2, then outlet
 * u_y where geo(x,NY-2,z)==3; rest of the file ignored.  Always float: the
 * reference parses with "%f". */
long orc_read_vel_file(const char *path, int nx, int ny, int nz, const int32_t *geo, float *inlety,
                       float *outlety) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    long cnt = 0;
    float tmp;
    for (int z = 0; z < nz; z++)
        for (int x = 0; x < nx; x++) {
            if (fscanf(f, "%f ", &tmp) != 1) tmp = 0.f; else cnt++;
            inlety[x + z * nx] = geo[CID(x, 1, z)] == 2 ? tmp : 0.f;
        }
    for (int z = 0; z < nz; z++)
        for (int x = 0; x < nx; x++) {
            if (fscanf(f, "%f ", &tmp) != 1) tmp = 0.f; else cnt++;
            outlety[x + z * nx] = geo[CID(x, ny - 2, z)] == 3 ? tmp : 0.f;
        }
    fclose(f);
    return cnt;
}
#undef CID
#endif /* ORC_EXPORT_INT */

/* ------------------------------------------------------------------------ */
/* Floating-point state                                                      */
/* ------------------------------------------------------------------------ */
typedef struct {
    int case_id, nx, ny, nz, nlat, fluid_label;
    int32_t *geo, *index; /* Cartesian box, owned copies */
    int32_t *cart;        /* [nlat] compact -> Cartesian id */
    REAL *src, *dst;      /* 19*nlat each, q-major (bif:1219-1220) */
    REAL *rho, *ux, *uy, *uz;
    REAL tau, u_max;
    REAL u_bc; /* speed literal of the boundary kernel (pos:590); initialize uses u_max (pos:44) */
    REAL *inlety, *outlety; /* nx*nz planes (bif:1200-1203) */
    REAL cor_uin, cor_uout, cor_usub;
    double pulse_amp, pulse_period; /* extension: u_in(t) = u_in * (1 + A sin(2 pi t / T)) */
    long step_count;
    int nfl;         /* fluid nodes */
    int32_t *flist;  /* their compact ids, ascending */
    int32_t *pull;   /* [19][nfl] compact id of the pull source x - c_q (or -1): the index
                        lookups of bif:445-571 hoisted out of the time loop */
    int nb;          /* boundary (label 1,2,3,5,6,7 / ldc 1,2) nodes */
    int32_t *blist;  /* their compact ids */
    REAL *bscratch;  /* 19*nb snapshot outputs */
    /* ldc only: model of the order in which ldc.cu's `update` launch really executes (see
     * orc_set_ldc_order): 0 = the defined semantics "all walls bounce, then fluid pulls" */
    int ldc_order;
    uint32_t *stale; /* [nfl] bit q: link q of this fluid node reads the wall slot BEFORE this launch's bounce */
    int32_t *brow;   /* [nlat] row of a boundary node in bscratch, -1 otherwise */
} FN(orc_state);

#define CIDS(x, y, z) ((size_t)(x) + (size_t)s->nx * ((size_t)(y) + (size_t)s->ny * (size_t)(z)))

/* ---- literal equilibrium forms (ldc:330-348 = pos:543-561 = bif:597-634) ---- */
static inline REAL feq_rest(REAL rw, REAL ux, REAL uy, REAL uz) {
    return rw * (R(1.0) - R(1.5) * ux * ux - R(1.5) * uy * uy - R(1.5) * uz * uz);
}
/* rho/18 * (1 +- 3 ua + 3 ua ua - 1.5 ub ub - 1.5 uc uc) */
static inline REAL feq_axis(REAL rw, int sgn, REAL ua, REAL ub, REAL uc) {
    REAL t = sgn > 0 ? R(1.0) + R(3.0) * ua : R(1.0) - R(3.0) * ua;
    return rw * (t + R(3.0) * ua * ua - R(1.5) * ub * ub - R(1.5) * uc * uc);
}
/* rho/36 * (1 +- 3 lin + 3 ua ua + 3 ub ub +- 9 ua ub - 1.5 uc uc)
 * form 0: +3(ua+ub), cross +   form 1: +3(ua-ub), cross -
 * form 2: +3(ub-ua), cross -   form 3: -3(ua+ub), cross +                      */
static inline REAL feq_diag(REAL rw, int form, REAL ua, REAL ub, REAL uc) {
    REAL t;
    switch (form) {
    case 0: t = R(1.0) + R(3.0) * (ua + ub); break;
    case 1: t = R(1.0) + R(3.0) * (ua - ub); break;
    case 2: t = R(1.0) + R(3.0) * (ub - ua); break;
    default: t = R(1.0) - R(3.0) * (ua + ub); break;
    }
    t = t + R(3.0) * ua * ua + R(3.0) * ub * ub;
    if (form == 0 || form == 3) t = t + R(9.0) * ua * ub;
    else t = t - R(9.0) * ua * ub;
    return rw * (t - R(1.5) * uc * uc);
}
/* Direction 14 carries a double literal in every copy of the reference
 * (ldc:344, pos:557, bif:625, cor:540): `3.0*uz*uz`, which promotes the rest of
 * that sum to double.  Restated for REAL=float; a no-op for REAL=double. */
static inline REAL feq_14(REAL rw, REAL ux, REAL uy, REAL uz) {
    REAL head = R(1.0) - R(3.0) * (ux + uz) + R(3.0) * ux * ux;
    double t = (double)head + 3.0 * (double)uz * (double)uz;
    t = t + (double)(R(9.0) * ux * uz);
    t = t - (double)(R(1.5) * uy * uy);
    return (REAL)((double)rw * t);
}
/* one direction of the literal form; r3,r18,r36 = rho/3, rho/18, rho/36 */
static inline REAL feq_q(int q, REAL r3, REAL r18, REAL r36, REAL ux, REAL uy, REAL uz) {
    switch (q) {
    case 0: return feq_rest(r3, ux, uy, uz);
    case 1: return feq_axis(r18, +1, ux, uy, uz);
    case 2: return feq_axis(r18, -1, ux, uy, uz);
    case 3: return feq_axis(r18, +1, uy, ux, uz);
    case 4: return feq_axis(r18, -1, uy, ux, uz);
    case 5: return feq_axis(r18, +1, uz, ux, uy);
    case 6: return feq_axis(r18, -1, uz, ux, uy);
    case 7: return feq_diag(r36, 0, ux, uy, uz);
    case 8: return feq_diag(r36, 1, ux, uy, uz);
    case 9: return feq_diag(r36, 2, ux, uy, uz);
    case 10: return feq_diag(r36, 3, ux, uy, uz);
    case 11: return feq_diag(r36, 0, ux, uz, uy);
    case 12: return feq_diag(r36, 1, ux, uz, uy);
    case 13: return feq_diag(r36, 2, ux, uz, uy);
    case 14: return feq_14(r36, ux, uy, uz);
    case 15: return feq_diag(r36, 0, uy, uz, ux);
    case 16: return feq_diag(r36, 2, uy, uz, ux);
    case 17: return feq_diag(r36, 1, uy, uz, ux);
    default: return feq_diag(r36, 3, uy, uz, ux);
    }
}
static void feq_all(REAL rho, REAL ux, REAL uy, REAL uz, REAL *feq) {
    REAL r3 = rho / R(3.0), r18 = rho / R(18.0), r36 = rho / R(36.0);
    for (int q = 0; q < 19; q++) feq[q] = feq_q(q, r3, r18, r36, ux, uy, uz);
}
/* factored form used only by ldc's initialize (ldc:542-571) */
static void feq_all_ldc_init(REAL rho, REAL ux, REAL uy, REAL uz, REAL *feq) {
    const REAL w0 = R(1.0) / R(3.0), w1 = R(1.0) / R(18.0), w2 = R(1.0) / R(36.0);
    REAL ux2 = ux * ux, uy2 = uy * uy, uz2 = uz * uz;
    REAL u2 = ux2 + uy2 + uz2, xy2 = ux2 + uy2, xz2 = ux2 + uz2, yz2 = uy2 + uz2;
    REAL xy = R(2.0) * ux * uy, xz = R(2.0) * ux * uz, yz = R(2.0) * uy * uz;
    feq[0] = rho * w0 * (R(1.0) - R(1.5) * u2);
    feq[1] = rho * w1 * (R(1.0) + R(3.0) * ux + R(4.5) * ux2 - R(1.5) * u2);
    feq[2] = rho * w1 * (R(1.0) - R(3.0) * ux + R(4.5) * ux2 - R(1.5) * u2);
    feq[3] = rho * w1 * (R(1.0) + R(3.0) * uy + R(4.5) * uy2 - R(1.5) * u2);
    feq[4] = rho * w1 * (R(1.0) - R(3.0) * uy + R(4.5) * uy2 - R(1.5) * u2);
    feq[5] = rho * w1 * (R(1.0) + R(3.0) * uz + R(4.5) * uz2 - R(1.5) * u2);
    feq[6] = rho * w1 * (R(1.0) - R(3.0) * uz + R(4.5) * uz2 - R(1.5) * u2);
    feq[7] = rho * w2 * (R(1.0) + R(3.0) * (ux + uy) + R(4.5) * (xy2 + xy) - R(1.5) * u2);
    feq[8] = rho * w2 * (R(1.0) + R(3.0) * (ux - uy) + R(4.5) * (xy2 - xy) - R(1.5) * u2);
    feq[9] = rho * w2 * (R(1.0) + R(3.0) * (uy - ux) + R(4.5) * (xy2 - xy) - R(1.5) * u2);
    feq[10] = rho * w2 * (R(1.0) - R(3.0) * (ux + uy) + R(4.5) * (xy2 + xy) - R(1.5) * u2);
    feq[11] = rho * w2 * (R(1.0) + R(3.0) * (ux + uz) + R(4.5) * (xz2 + xz) - R(1.5) * u2);
    feq[12] = rho * w2 * (R(1.0) + R(3.0) * (ux - uz) + R(4.5) * (xz2 - xz) - R(1.5) * u2);
    feq[13] = rho * w2 * (R(1.0) + R(3.0) * (uz - ux) + R(4.5) * (xz2 - xz) - R(1.5) * u2);
    feq[14] = rho * w2 * (R(1.0) - R(3.0) * (ux + uz) + R(4.5) * (xz2 + xz) - R(1.5) * u2);
    feq[15] = rho * w2 * (R(1.0) + R(3.0) * (uy + uz) + R(4.5) * (yz2 + yz) - R(1.5) * u2);
    feq[16] = rho * w2 * (R(1.0) + R(3.0) * (uz - uy) + R(4.5) * (yz2 - yz) - R(1.5) * u2);
    feq[17] = rho * w2 * (R(1.0) + R(3.0) * (uy - uz) + R(4.5) * (yz2 - yz) - R(1.5) * u2);
    feq[18] = rho * w2 * (R(1.0) - R(3.0) * (uy + uz) + R(4.5) * (yz2 + yz) - R(1.5) * u2);
}

/* ------------------------------------------------------------------------ */
FN(orc_state) *FN(orc_create)(int case_id, int nx, int ny, int nz, const int32_t *geo,
                              const int32_t *index, int nlat, double tau, double u_max) {
    FN(orc_state) *s = (FN(orc_state) *)calloc(1, sizeof(*s));
    size_t nbox = (size_t)nx * ny * nz;
    s->case_id = case_id;
    s->nx = nx, s->ny = ny, s->nz = nz, s->nlat = nlat;
    s->fluid_label = case_id == CASE_LDC ? 3 : 4;
    s->tau = (REAL)tau;
    s->u_max = (REAL)u_max;
    s->u_bc = (REAL)u_max;
    s->geo = (int32_t *)malloc(nbox * sizeof(int32_t));
    s->index = (int32_t *)malloc(nbox * sizeof(int32_t));
    memcpy(s->geo, geo, nbox * sizeof(int32_t));
    memcpy(s->index, index, nbox * sizeof(int32_t));
    s->cart = (int32_t *)malloc((size_t)nlat * sizeof(int32_t));
    s->nb = 0;
    for (size_t c = 0; c < nbox; c++)
        if (index[c] >= 0) {
            s->cart[index[c]] = (int32_t)c;
            int g = geo[c];
            if (g > 0 && g != s->fluid_label) s->nb++;
        }
    s->blist = (int32_t *)malloc((size_t)(s->nb ? s->nb : 1) * sizeof(int32_t));
    int k = 0;
    for (int i = 0; i < nlat; i++) {
        int g = geo[s->cart[i]];
        if (g > 0 && g != s->fluid_label) s->blist[k++] = i;
    }
    s->bscratch = (REAL *)calloc((size_t)19 * (s->nb ? s->nb : 1), sizeof(REAL));
    s->src = (REAL *)calloc((size_t)19 * nlat, sizeof(REAL));
    s->dst = (REAL *)calloc((size_t)19 * nlat, sizeof(REAL));
    s->rho = (REAL *)calloc(nlat, sizeof(REAL));
    s->ux = (REAL *)calloc(nlat, sizeof(REAL));
    s->uy = (REAL *)calloc(nlat, sizeof(REAL));
    s->uz = (REAL *)calloc(nlat, sizeof(REAL));
    s->inlety = (REAL *)calloc((size_t)nx * nz, sizeof(REAL));
    s->outlety = (REAL *)calloc((size_t)nx * nz, sizeof(REAL));
    s->nfl = 0;
    for (int i = 0; i < nlat; i++) s->nfl += geo[s->cart[i]] == s->fluid_label;
    s->flist = (int32_t *)malloc((size_t)(s->nfl ? s->nfl : 1) * sizeof(int32_t));
    s->pull = (int32_t *)malloc((size_t)19 * (s->nfl ? s->nfl : 1) * sizeof(int32_t));
    k = 0;
    for (int i = 0; i < nlat; i++)
        if (geo[s->cart[i]] == s->fluid_label) s->flist[k++] = i;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < s->nfl; j++) {
        size_t c = (size_t)s->cart[s->flist[j]];
        int x = (int)(c % nx), y = (int)((c / nx) % ny), z = (int)(c / ((size_t)nx * ny));
        for (int q = 0; q < 19; q++) {
            int xx = x - CX[q], yy = y - CY[q], zz = z - CZ[q];
            int ok = xx >= 0 && xx < nx && yy >= 0 && yy < ny && zz >= 0 && zz < nz;
            s->pull[(size_t)q * s->nfl + j] = ok ? index[(size_t)xx + (size_t)nx * ((size_t)yy + (size_t)ny * zz)] : -1;
        }
    }
    return s;
}

void FN(orc_destroy)(FN(orc_state) *s) {
    if (!s) return;
    free(s->geo), free(s->index), free(s->cart), free(s->blist), free(s->bscratch), free(s->flist), free(s->pull);
    free(s->stale), free(s->brow);
    free(s->src), free(s->dst), free(s->rho), free(s->ux), free(s->uy), free(s->uz);
    free(s->inlety), free(s->outlety);
    free(s);
}

/* bif: BC planes as read by read_vel (float in the file, bif:296-325) */
void FN(orc_set_bc_planes)(FN(orc_state) *s, const float *inlety, const float *outlety) {
    for (size_t i = 0; i < (size_t)s->nx * s->nz; i++) {
        s->inlety[i] = (REAL)inlety[i];
        s->outlety[i] = (REAL)outlety[i];
    }
}
/* cor: the three literal speeds (cor:302-306, 717, 796, 871), already / C_U */
void FN(orc_set_cor_speeds)(FN(orc_state) *s, double uin, double uout, double usub) {
    s->cor_uin = (REAL)uin, s->cor_uout = (REAL)uout, s->cor_usub = (REAL)usub;
}
void FN(orc_set_u_bc)(FN(orc_state) *s, double u_bc) { s->u_bc = (REAL)u_bc; }

/* ldc only.  ldc.cu's `update` bounces the wall nodes IN PLACE on d_scr (ldc:75-202) in the same
 * launch in which fluid nodes pull from them (ldc:204-313): whether a fluid node sees this launch's
 * bounce ("fresh") or the slot's previous content -- the bounce of two iterations ago, the buffers
 * alternate (ldc:664-666) -- ("stale") depends on the order the launch executes in.  Most of that
 * order is fixed by the code: a thread owns a z-column of BLOCK_Z = 8 nodes and walks it from the top
 * down (koff = 7..0, ldc:66), all 512 blocks of 64 threads are resident at once, and inside one koff
 * iteration the wall branch (ldc:75) precedes the fluid branch (ldc:204) in program order.  With
 * t(node) = 7 - z % 8 (the koff iteration that processes it):
 *   t(wall) < t(fluid)                      fresh   (e.g. the z-high wall, and every link with c_z = -1
 *                                                    unless the fluid node is the top of its column)
 *   t(wall) > t(fluid)                      stale   (e.g. the z-low wall seen from z = 2)
 *   equal, same warp (x/8, y/8, z/8, (y%8)/4 equal)   fresh: divergent branches of one warp run in program order
 *   equal, different warp                   a true race: mode 1 calls it fresh, mode 2 stale
 * mode 0 is the defined semantics (everything fresh) the product implements. */
void FN(orc_set_ldc_order)(FN(orc_state) *s, int mode) {
    s->ldc_order = mode;
    free(s->stale), free(s->brow);
    s->stale = NULL, s->brow = NULL;
    if (!mode || s->case_id != CASE_LDC) return;
    s->stale = (uint32_t *)calloc((size_t)(s->nfl ? s->nfl : 1), sizeof(uint32_t));
    s->brow = (int32_t *)malloc((size_t)s->nlat * sizeof(int32_t));
    for (int i = 0; i < s->nlat; i++) s->brow[i] = -1;
    for (int b = 0; b < s->nb; b++) s->brow[s->blist[b]] = b;
    for (int j = 0; j < s->nfl; j++) {
        size_t c = (size_t)s->cart[s->flist[j]];
        int x = (int)(c % s->nx), y = (int)((c / s->nx) % s->ny), z = (int)(c / ((size_t)s->nx * s->ny));
        for (int q = 1; q < 19; q++) {
            int n = s->pull[(size_t)q * s->nfl + j];
            if (n < 0 || s->geo[s->cart[n]] != 1) continue;
            int xs = x - CX[q], ys = y - CY[q], zs = z - CZ[q];
            int tf = 7 - z % 8, tw = 7 - zs % 8, st;
            if (tw != tf) st = tw > tf;
            else if (xs / 8 == x / 8 && ys / 8 == y / 8 && (ys % 8) / 4 == (y % 8) / 4) st = 0;
            else if (mode == 3) {
                /* a reader whose own warp holds no wall node at this z skips the wall branch and issues its
                 * loads at once, ahead of the other warp's bounce; a mixed warp first waits for its own walls */
                int has_wall = 0;
                for (int yy = y / 4 * 4; yy < y / 4 * 4 + 4 && yy < s->ny; yy++)
                    for (int xx = x / 8 * 8; xx < x / 8 * 8 + 8 && xx < s->nx; xx++) has_wall |= s->geo[CIDS(xx, yy, z)] == 1;
                st = !has_wall;
            } else st = mode == 2;
            if (st) s->stale[j] |= 1u << q;
        }
    }
}
void FN(orc_set_pulse)(FN(orc_state) *s, double amp, double period) {
    s->pulse_amp = amp, s->pulse_period = period;
}

/* pos:301 / pos:597 : analytic parabola at the node's own (i,k), in float like
 * the reference (powf), widened for REAL=double. */
static inline REAL parabola(const FN(orc_state) *s, REAL umax, int i, int k) {
#if defined(ORC_IS_DOUBLE)
    double cx = (s->nx - 1) / 2.0, cz = (s->nz - 1) / 2.0, r = (s->nx - 1) / 2.0;
    return (REAL)(umax * (1.0 - (pow(i - cx, 2.0) + pow(k - cz, 2.0)) / pow(r, 2.0)));
#else
    float cx = (s->nx - 1) / 2.0f, cz = (s->nz - 1) / 2.0f, r = (s->nx - 1) / 2.0f;
    return umax * (1.0f - (powf(i - cx, 2.f) + powf(k - cz, 2.f)) / powf(r, 2.f));
#endif
}

/* ldc:504-580, pos:273-382, bif:329-427, cor:277-350 */
void FN(orc_initialize)(FN(orc_state) *s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, nlat = s->nlat;
    for (int i = 0; i < nlat; i++) s->rho[i] = R(1.0), s->ux[i] = s->uy[i] = s->uz[i] = R(0.0);
    if (s->case_id == CASE_LDC) {
        for (int z = 0; z < nz; z++)
            for (int x = 0; x < nx; x++) {
                s->uz[s->index[CIDS(x, ny - 1, z)]] = s->u_max;
                s->uz[s->index[CIDS(x, ny - 2, z)]] = s->u_max;
            }
    } else if (s->case_id == CASE_POS) {
        const int planes[4] = {0, 1, ny - 1, ny - 2};
        for (int p = 0; p < 4; p++)
            for (int x = 0; x < nx; x++)
                for (int z = 0; z < nz; z++) {
                    int idx = s->index[CIDS(x, planes[p], z)];
                    if (idx >= 0) s->uy[idx] = parabola(s, s->u_max, x, z);
                }
    } else if (s->case_id == CASE_BIF) {
        for (int x = 0; x < nx; x++)
            for (int z = 0; z < nz; z++) {
                int idx = s->index[CIDS(x, 1, z)];
                if (idx >= 0) s->ux[idx] = R(0.0), s->uy[idx] = s->inlety[x + z * nx];
                idx = s->index[CIDS(x, ny - 2, z)];
                if (idx >= 0) s->ux[idx] = R(0.0), s->uy[idx] = s->outlety[x + z * nx];
            }
    } else {
        for (int i = 0; i < nlat; i++) {
            int g = s->geo[s->cart[i]];
            if (g == 2) s->ux[i] = s->cor_uin;
            if (g == 3) s->ux[i] = s->cor_uout;
            if (g == 5 || g == 6 || g == 7) s->uz[i] = s->cor_usub;
        }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nlat; i++) {
        REAL feq[19];
        if (s->case_id == CASE_LDC) feq_all_ldc_init(s->rho[i], s->ux[i], s->uy[i], s->uz[i], feq);
        else feq_all(s->rho[i], s->ux[i], s->uy[i], s->uz[i], feq);
        for (int q = 0; q < 19; q++) {
            s->dst[(size_t)q * nlat + i] = feq[q];
            s->src[(size_t)q * nlat + i] = feq[q];
        }
    }
    /* the reference never writes rho/u of non-fluid nodes on the device; its
     * host copies are overwritten by the first D2H -> 0. */
    for (int i = 0; i < nlat; i++) s->rho[i] = s->ux[i] = s->uy[i] = s->uz[i] = R(0.0);
    s->step_count = 0;
}

/* compact id of the neighbour of Cartesian node (x,y,z) displaced by (dx,dy,dz);
 * -1 when outside the box or not stored.  ywrap: pos/bif wrap y mod NY in the
 * wall branch only (bif:670,677). */
static inline int nb_idx(const FN(orc_state) *s, int x, int y, int z, int dx, int dy, int dz, int ywrap) {
    int i = x + dx, j = y + dy, k = z + dz;
    if (ywrap) j = (j + s->ny) % s->ny;
    if (i < 0 || i >= s->nx || j < 0 || j >= s->ny || k < 0 || k >= s->nz) return -1;
    return s->index[CIDS(i, j, k)];
}

/* wall branch: ldc:75-202 (on src) and bif:654-799 (on dst): gather
 * fnq[p] = buf_p(w - c_p), then buf_q(w) = fnq[opp q].  Output to out[1..18]. */
static void wall_gather(const FN(orc_state) *s, const REAL *buf, int x, int y, int z, int ywrap, REAL *out) {
    REAL fnq[19];
    for (int p = 1; p < 19; p++) {
        int n = nb_idx(s, x, y, z, -CX[p], -CY[p], -CZ[p], ywrap);
        fnq[p] = n >= 0 ? buf[(size_t)p * s->nlat + n] : R(0.0);
    }
    for (int q = 1; q < 19; q++) out[q] = fnq[OPP[q]];
}

/* fluid branch of `update`: ldc:204-369, pos:405-582, bif:445-635 */
static void update_fluid(FN(orc_state) *s) {
    const int nlat = s->nlat, nfl = s->nfl;
    const REAL tau = s->tau;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nfl; j++) {
        const int i = s->flist[j];
        REAL f[19], feq[19];
        for (int q = 0; q < 19; q++) {
            int n = s->pull[(size_t)q * nfl + j];
            f[q] = n >= 0 ? s->src[(size_t)q * nlat + n] : R(0.0);
            /* ldc order model: this launch's wall bounce sits in bscratch until the launch ends */
            if (s->ldc_order && n >= 0 && s->brow[n] >= 0 && !(s->stale[j] & (1u << q)))
                f[q] = s->bscratch[(size_t)19 * s->brow[n] + q];
        }
        REAL rho = R(0.0);
        for (int q = 0; q < 19; q++) rho = rho + f[q];
        REAL ux = (f[1] - f[2] + f[7] + f[8] - f[9] - f[10] + f[11] + f[12] - f[13] - f[14]) / rho;
        REAL uy = (f[3] - f[4] + f[7] - f[8] + f[9] - f[10] + f[15] - f[16] + f[17] - f[18]) / rho;
        REAL uz = (f[5] - f[6] + f[11] - f[12] + f[13] - f[14] + f[15] + f[16] - f[17] - f[18]) / rho;
        s->rho[i] = rho, s->ux[i] = ux, s->uy[i] = uy, s->uz[i] = uz;
        feq_all(rho, ux, uy, uz, feq);
        for (int q = 0; q < 19; q++) s->dst[(size_t)q * nlat + i] = f[q] - (f[q] - feq[q]) / tau;
    }
}

/* prescribed-velocity equilibrium along one axis, literal polynomial forms:
 * c_q.axis = +1 : rw*(1 + 3u + 3u u)   (pos:766 style)
 * c_q.axis = -1 : rw*(1 - 3u + 3u u)
 * c_q.axis =  0 : rw*(1 - 1.5u u)      (ldc:402) */
static inline REAL feq_bc_axis(REAL rw, int cs, REAL u) {
    if (cs > 0) return rw * (R(1.0) + R(3.0) * u + R(3.0) * u * u);
    if (cs < 0) return rw * (R(1.0) - R(3.0) * u + R(3.0) * u * u);
    return rw * (R(1.0) - R(1.5) * u * u);
}

/* One non-equilibrium-extrapolation population (ldc:393-455, pos:748-891,
 * bif:877-1021, cor:716-942): nb = b + c_q,
 *   dst_q(b) = feq_bc + (dst_q(nb) - feq_q(rho_nb,u_nb)) * (1 - 1/tau).
 * kind 0 (V): rho_bc = rho_nb, u = u_presc on vaxis.  kind 1 (P): rho_bc = 1,
 * u = u_nb.  kind 2 (VP): rho_bc = 1, u = u_presc. */
static REAL neq_extrap(const FN(orc_state) *s, int x, int y, int z, int q, int kind, int vaxis, REAL upresc) {
    int n = nb_idx(s, x, y, z, CX[q], CY[q], CZ[q], 0);
    REAL rho = R(0.0), ux = R(0.0), uy = R(0.0), uz = R(0.0), fnb = R(0.0);
    if (n >= 0) rho = s->rho[n], ux = s->ux[n], uy = s->uy[n], uz = s->uz[n], fnb = s->dst[(size_t)q * s->nlat + n];
    REAL wden = q == 0 ? R(3.0) : (q < 7 ? R(18.0) : R(36.0));
    REAL feq = feq_q(q, rho / R(3.0), rho / R(18.0), rho / R(36.0), ux, uy, uz);
    REAL tmp;
    if (kind == 1) {
        REAL one = R(1.0);
        tmp = feq_q(q, one / R(3.0), one / R(18.0), one / R(36.0), ux, uy, uz);
    } else {
        int cs = vaxis == 0 ? CX[q] : (vaxis == 1 ? CY[q] : CZ[q]);
        REAL rw = kind == 0 ? rho / wden : R(1.0) / wden;
        tmp = feq_bc_axis(rw, cs, upresc);
    }
    return tmp + (fnb - feq) * (R(1.0) - R(1.0) / s->tau);
}

static const int SET_YM[5] = {4, 8, 10, 16, 18};  /* c_y = -1 */
static const int SET_YP[5] = {3, 7, 9, 15, 17};   /* c_y = +1 */
static const int SET_XP[5] = {1, 7, 8, 11, 12};   /* c_x = +1 */
static const int SET_XM[5] = {2, 9, 10, 13, 14};  /* c_x = -1 */
static const int SET_ZM[5] = {6, 12, 14, 17, 18}; /* c_z = -1 */

/* boundary_stream on dst: ldc:373-458, pos:585-893, bif:639-1023, cor:555-944.
 * Two phases (gather into scratch, then write) = snapshot semantics. */
static void boundary_stream(FN(orc_state) *s) {
    const int nlat = s->nlat, nb = s->nb;
    const int ywrap = (s->case_id == CASE_POS || s->case_id == CASE_BIF);
    double scale = 1.0;
    if (s->pulse_amp != 0.0) scale = 1.0 + s->pulse_amp * sin(2.0 * M_PI * (double)s->step_count / s->pulse_period);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; b++) {
        int i = s->blist[b];
        size_t c = (size_t)s->cart[i];
        int g = s->geo[c];
        int x = (int)(c % s->nx), y = (int)((c / s->nx) % s->ny), z = (int)(c / ((size_t)s->nx * s->ny));
        REAL *out = s->bscratch + (size_t)19 * b;
        for (int q = 0; q < 19; q++) out[q] = s->dst[(size_t)q * nlat + i];
        if (g == 1) {
            if (s->case_id != CASE_LDC) wall_gather(s, s->dst, x, y, z, ywrap, out);
            continue;
        }
        const int *set = NULL;
        int kind = 0, vaxis = 1;
        REAL up = R(0.0);
        switch (s->case_id) {
        case CASE_LDC: /* lid, ldc:391-456 */
            if (g == 2) set = SET_YM, kind = 0, vaxis = 2, up = s->u_bc;
            break;
        case CASE_POS: /* pos:748-891; u at the BC node's own (i,k), pos:597 */
            if (g == 3) set = SET_YM, kind = 0, vaxis = 1, up = parabola(s, s->u_bc, x, z);
            if (g == 2) set = SET_YP, kind = 0, vaxis = 1, up = parabola(s, s->u_bc, x, z);
            break;
        case CASE_BIF: /* outlet P bif:877-948, inlet V bif:950-1021 */
            if (g == 3) set = SET_YM, kind = 1;
            if (g == 2) set = SET_YP, kind = 0, vaxis = 1, up = (REAL)(s->inlety[x + z * s->nx] * (REAL)scale);
            break;
        default: /* cor:716-942 */
            if (g == 2) set = SET_XP, kind = 2, vaxis = 0, up = (REAL)(s->cor_uin * (REAL)scale);
            if (g == 3) set = SET_XM, kind = 0, vaxis = 0, up = s->cor_uout;
            if (g == 5 || g == 6 || g == 7) set = SET_ZM, kind = 0, vaxis = 2, up = s->cor_usub;
            break;
        }
        if (set)
            for (int k = 0; k < 5; k++) out[set[k]] = neq_extrap(s, x, y, z, set[k], kind, vaxis, up);
    }
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; b++) {
        int i = s->blist[b];
        const REAL *out = s->bscratch + (size_t)19 * b;
        for (int q = 0; q < 19; q++) s->dst[(size_t)q * nlat + i] = out[q];
    }
}

/* ldc only: wall bounce on src before the fluid pull (ldc:75-202).  phase 1 gathers into bscratch,
 * phase 2 writes; with an order model (ldc_order != 0) the fluid update runs between the two. */
static void ldc_wall_bounce_src(FN(orc_state) *s, int phase) {
    const int nlat = s->nlat, nb = s->nb;
    if (phase != 2) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; b++) {
        int i = s->blist[b];
        size_t c = (size_t)s->cart[i];
        REAL *out = s->bscratch + (size_t)19 * b;
        for (int q = 0; q < 19; q++) out[q] = s->src[(size_t)q * nlat + i];
        if (s->geo[c] != 1) continue;
        int x = (int)(c % s->nx), y = (int)((c / s->nx) % s->ny), z = (int)(c / ((size_t)s->nx * s->ny));
        wall_gather(s, s->src, x, y, z, 0, out);
    }
    }
    if (phase == 1) return;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; b++) {
        int i = s->blist[b];
        const REAL *out = s->bscratch + (size_t)19 * b;
        for (int q = 0; q < 19; q++) s->src[(size_t)q * nlat + i] = out[q];
    }
}

/* one iteration of the main loop: ldc:654-666, bif:1249-1257 */
void FN(orc_step)(FN(orc_state) *s, int nsteps) {
    for (int it = 0; it < nsteps; it++) {
        const int ordered = s->case_id == CASE_LDC && s->ldc_order;
        if (s->case_id == CASE_LDC) ldc_wall_bounce_src(s, ordered ? 1 : 0);
        update_fluid(s);
        if (ordered) ldc_wall_bounce_src(s, 2);
        boundary_stream(s);
        REAL *t = s->src;
        s->src = s->dst;
        s->dst = t;
        s->step_count++;
    }
}

void FN(orc_get_fields)(const FN(orc_state) *s, REAL *rho, REAL *ux, REAL *uy, REAL *uz) {
    size_t n = (size_t)s->nlat * sizeof(REAL);
    memcpy(rho, s->rho, n), memcpy(ux, s->ux, n), memcpy(uy, s->uy, n), memcpy(uz, s->uz, n);
}
/* populations "as if in d_scr after the swap", 19*nlat q-major */
void FN(orc_get_populations)(const FN(orc_state) *s, REAL *f) {
    memcpy(f, s->src, (size_t)19 * s->nlat * sizeof(REAL));
}

/* ldc:460-466,662 : S = sum_i sqrt(ux^2+uy^2+uz^2) over all NLATTICE entries.
 * The reference sums floats with thrust (order unspecified); here sequential
 * in double, so comparisons are tolerance-based. */
double FN(orc_velsum)(const FN(orc_state) *s) {
    double acc = 0.0;
    for (int i = 0; i < s->nlat; i++) {
        REAL v = (REAL)sqrt((double)(s->ux[i] * s->ux[i] + s->uy[i] * s->uy[i] + s->uz[i] * s->uz[i]));
        acc += (double)v;
    }
    return acc;
}
/* bif:1158-1175 (label >= 4) / cor:1013-1030 (label == 4): sum of u^2 over the
 * trimmed box z in [1,NZ-2], y in [2,NY-3], x in [1,NX-2], long double. */
double FN(orc_calc_res)(const FN(orc_state) *s) {
    long double sum = 0.0L;
    for (int z = 1; z < s->nz - 1; z++)
        for (int y = 2; y < s->ny - 2; y++)
            for (int x = 1; x < s->nx - 1; x++) {
                size_t c = CIDS(x, y, z);
                int g = s->geo[c];
                int ok = s->case_id == CASE_COR ? (g == 4) : (g >= 4);
                if (!ok) continue;
                int i = s->index[c];
                REAL v = s->ux[i] * s->ux[i] + s->uy[i] * s->uy[i] + s->uz[i] * s->uz[i];
                sum += (long double)v;
            }
    return (double)sum;
}

/* CPU baseline: wall-clock seconds for nsteps iterations (threads = OMP_NUM_THREADS). */
double FN(orc_time_steps)(FN(orc_state) *s, int nsteps) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    FN(orc_step)(s, nsteps);
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
int FN(orc_num_fluid)(const FN(orc_state) *s) {
    int n = 0;
    for (int i = 0; i < s->nlat; i++) n += s->geo[s->cart[i]] == s->fluid_label;
    return n;
}
