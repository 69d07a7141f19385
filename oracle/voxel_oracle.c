/* CPU restatement of the STL voxeliser (TEST INFRASTRUCTURE ONLY -- used by tests/, never
 * by the product library; the product is csrc/lbm_voxel.cu).
 *
 * The reference ships the carotid surface (bifurcation/bif.stl) and its voxelisation
 * (bifurcation/geo.txt) but not the MATLAB pre-processing step in between
 * (bifurcation/README.md:1-5); this is the published ray-parity algorithm
 * (solid voxelisation, Schwarz & Seidel 2010, with the rasteriser's top-left tie rule), pinned by
 * reproducing the shipped geo.txt from the shipped bif.stl up to surface voxels
 * (tests/golden/make_voxel_fit.py, tests/test_voxel_cpu.py).  "parity partial": the original tool is absent.
 *
 * Voxel (i,j,k) has its centre at origin + (i+.5, j+.5, k+.5) * h.  A ray along +x through the centre
 * of every (j,k) row is intersected with every triangle; a crossing at abscissa xc toggles the marker
 * bit of the first voxel whose centre lies beyond it; a prefix XOR along the row turns markers into
 * inside/outside.  The surface only has to be closed around the ray direction (vessel openings along
 * y are fine).  Ties (centre exactly on a projected edge or vertex) are broken so that two triangles
 * sharing an edge claim its points exactly once when they lie on opposite sides of it in the
 * projection, and zero or two times when they lie on the same side (a silhouette edge).
 * A row with an ODD number of crossings (its ray passes the ragged rim of an open end and meets only one
 * of the two walls) has no inside/outside answer; it is cleared, so a leak never paints the row to the
 * edge of the box.
 *
 * All arithmetic in double, no FMA contraction (-ffp-contract=off): the CUDA kernel, compiled with
 * -fmad=false, produces the same bits.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double origin[3];
    double spacing;
    int32_t nx, ny, nz;
    int32_t reserved;
} vox_grid;

/* edge function of the directed edge p0->p1 at point (py,pz), evaluated with the endpoints in a
 * canonical order so that both triangles sharing the edge compute the same magnitude */
static double edge_fn(double p0y, double p0z, double p1y, double p1z, double py, double pz, int *tie_in) {
    int swap = (p1y < p0y) || (p1y == p0y && p1z < p0z);
    double ay = swap ? p1y : p0y, az = swap ? p1z : p0z, by = swap ? p0y : p1y, bz = swap ? p0z : p1z;
    double e = (py - ay) * (bz - az) - (pz - az) * (by - ay);
    if (swap) e = -e;
    /* tie rule on the ACTUAL direction d = p1 - p0: points on the edge belong to it iff d.y > 0, or
     * d.y == 0 and d.z > 0 -- complementary for the two directions of one edge */
    double dy = p1y - p0y, dz = p1z - p0z;
    *tie_in = (dy > 0.0) || (dy == 0.0 && dz > 0.0);
    return e;
}

/* does the +x ray through (py,pz) cross triangle v (3 vertices x 3 coords)?  if so *xc = abscissa */
static int ray_hits(const float *v, double py, double pz, double *xc) {
    double ax = v[0], ay = v[1], az = v[2], bx = v[3], by = v[4], bz = v[5], cx = v[6], cy = v[7], cz = v[8];
    double area = (by - ay) * (cz - az) - (bz - az) * (cy - ay); /* projected, signed */
    if (area == 0.0) return 0;                                   /* parallel to the ray */
    if (area < 0.0) {                                            /* make the projection counter-clockwise */
        double t;
        t = bx, bx = cx, cx = t;
        t = by, by = cy, cy = t;
        t = bz, bz = cz, cz = t;
    }
    int t0, t1, t2;
    double e0 = edge_fn(ay, az, by, bz, py, pz, &t0);
    double e1 = edge_fn(by, bz, cy, cz, py, pz, &t1);
    double e2 = edge_fn(cy, cz, ay, az, py, pz, &t2);
    /* counter-clockwise: inside is where every edge function is <= 0 with this sign convention */
    if (e0 > 0.0 || e1 > 0.0 || e2 > 0.0) return 0;
    if ((e0 == 0.0 && !t0) || (e1 == 0.0 && !t1) || (e2 == 0.0 && !t2)) return 0;
    /* plane through a with normal n = (b-a) x (c-a); n.x is the projected area (non-zero) */
    double nx = (by - ay) * (cz - az) - (bz - az) * (cy - ay);
    double ny = (bz - az) * (cx - ax) - (bx - ax) * (cz - az);
    double nz = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    *xc = ax - (ny * (py - ay) + nz * (pz - az)) / nx;
    return 1;
}

/* tri: [ntri][3][3] floats.  out: [z_end - z_begin][ny][nx] bytes (1 inside, 0 outside). */
int vox_oracle(const float *tri, int64_t ntri, const vox_grid *g, int32_t z_begin, int32_t z_end, uint8_t *out) {
    const int nx = g->nx, ny = g->ny, W = (nx + 31) / 32;
    const int nzl = z_end - z_begin;
    if (nzl <= 0) return 0;
    const double h = g->spacing, ox = g->origin[0], oy = g->origin[1], oz = g->origin[2];
    uint32_t *mark = (uint32_t *)calloc((size_t)nzl * ny * W, sizeof(uint32_t));
    if (!mark) return -1;
    for (int64_t t = 0; t < ntri; t++) {
        const float *v = tri + 9 * t;
        double ymin = fmin(v[1], fmin(v[4], v[7])), ymax = fmax(v[1], fmax(v[4], v[7]));
        double zmin = fmin(v[2], fmin(v[5], v[8])), zmax = fmax(v[2], fmax(v[5], v[8]));
        /* rows whose centre can lie inside the projected bounding box (one extra on each side) */
        int j0 = (int)floor((ymin - oy) / h - 0.5) - 1, j1 = (int)ceil((ymax - oy) / h - 0.5) + 1;
        int k0 = (int)floor((zmin - oz) / h - 0.5) - 1, k1 = (int)ceil((zmax - oz) / h - 0.5) + 1;
        if (j0 < 0) j0 = 0;
        if (j1 > ny - 1) j1 = ny - 1;
        if (k0 < z_begin) k0 = z_begin;
        if (k1 > z_end - 1) k1 = z_end - 1;
        for (int k = k0; k <= k1; k++)
            for (int j = j0; j <= j1; j++) {
                double py = oy + ((double)j + 0.5) * h, pz = oz + ((double)k + 0.5) * h, xc;
                if (!ray_hits(v, py, pz, &xc)) continue;
                /* first voxel whose centre ox + (i + .5) h lies strictly beyond the crossing */
                double fi = floor((xc - ox) / h - 0.5) + 1.0;
                if (fi >= (double)nx) continue;
                int i0 = fi < 0.0 ? 0 : (int)fi;
                mark[((size_t)(k - z_begin) * ny + j) * W + (i0 >> 5)] ^= 1u << (i0 & 31);
            }
    }
    for (int k = 0; k < nzl; k++)
        for (int j = 0; j < ny; j++) {
            const uint32_t *m = mark + ((size_t)k * ny + j) * W;
            uint8_t *o = out + ((size_t)k * ny + j) * nx;
            unsigned par = 0, total = 0;
            for (int w = 0; w < W; w++) total ^= (unsigned)__builtin_popcount(m[w]) & 1u;
            for (int i = 0; i < nx; i++) {
                par ^= (m[i >> 5] >> (i & 31)) & 1u;
                o[i] = (uint8_t)(total ? 0u : par);
            }
        }
    free(mark);
    return 0;
}
