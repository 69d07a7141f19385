// STL -> binary voxel field on the GPU: the front end that produces geo.txt-style inside/outside
// masks (the reference ships bifurcation/bif.stl and its voxelisation bifurcation/geo.txt, but not the
// MATLAB step in between: bifurcation/README.md:1-5).  Solid voxelisation by ray parity
// with the rasteriser's top-left tie rule; the algorithm and its tie rules are stated in
// oracle/voxel_oracle.c, which this file matches bit for bit (compiled with -fmad=false).
//
//   k_vox_mark : one thread per (triangle, row of its projected bounding box): +x ray through the row's
//                voxel centres; a crossing toggles the marker bit of the first voxel beyond it
//                (atomicXor -- order-independent, so the result is deterministic)
//   k_vox_fill : one warp per row: prefix XOR of the marker bits (in-word shifts + a warp scan of the
//                word parities) -> one byte per voxel, coalesced; rows with an odd number of crossings
//                (leaks at the ragged rim of an open end) are cleared
// A z-range can be voxelised on its own (what one rank of a z-slab run hands to lbm_set_flag_slab).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/lbm_b200.h"

namespace {

thread_local std::string g_vox_error;

struct Grid {
    double ox, oy, oz, h;
    int nx, ny, nz;
};

__device__ __forceinline__ double edge_fn(double p0y, double p0z, double p1y, double p1z, double py, double pz, bool &tie_in) {
    const bool swap = (p1y < p0y) || (p1y == p0y && p1z < p0z);
    const double ay = swap ? p1y : p0y, az = swap ? p1z : p0z, by = swap ? p0y : p1y, bz = swap ? p0z : p1z;
    double e = (py - ay) * (bz - az) - (pz - az) * (by - ay);
    if (swap) e = -e;
    const double dy = p1y - p0y, dz = p1z - p0z;
    tie_in = (dy > 0.0) || (dy == 0.0 && dz > 0.0);
    return e;
}

__device__ __forceinline__ bool ray_hits(const float *v, double py, double pz, double &xc) {
    double ax = v[0], ay = v[1], az = v[2], bx = v[3], by = v[4], bz = v[5], cx = v[6], cy = v[7], cz = v[8];
    const double area = (by - ay) * (cz - az) - (bz - az) * (cy - ay);
    if (area == 0.0) return false;
    if (area < 0.0) {
        double t;
        t = bx, bx = cx, cx = t;
        t = by, by = cy, cy = t;
        t = bz, bz = cz, cz = t;
    }
    bool t0, t1, t2;
    const double e0 = edge_fn(ay, az, by, bz, py, pz, t0);
    const double e1 = edge_fn(by, bz, cy, cz, py, pz, t1);
    const double e2 = edge_fn(cy, cz, ay, az, py, pz, t2);
    if (e0 > 0.0 || e1 > 0.0 || e2 > 0.0) return false;
    if ((e0 == 0.0 && !t0) || (e1 == 0.0 && !t1) || (e2 == 0.0 && !t2)) return false;
    const double nx = (by - ay) * (cz - az) - (bz - az) * (cy - ay);
    const double ny = (bz - az) * (cx - ax) - (bx - ax) * (cz - az);
    const double nz = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
    xc = ax - (ny * (py - ay) + nz * (pz - az)) / nx;
    return true;
}

// rows of a triangle's projected bounding box, clipped to the grid and the z range
struct RowBox {
    int j0, j1, k0, k1;
};
__device__ __forceinline__ RowBox row_box(const float *v, const Grid &g, int z_begin, int z_end) {
    const double ymin = fmin((double)v[1], fmin((double)v[4], (double)v[7])), ymax = fmax((double)v[1], fmax((double)v[4], (double)v[7]));
    const double zmin = fmin((double)v[2], fmin((double)v[5], (double)v[8])), zmax = fmax((double)v[2], fmax((double)v[5], (double)v[8]));
    RowBox b;
    b.j0 = (int)floor((ymin - g.oy) / g.h - 0.5) - 1, b.j1 = (int)ceil((ymax - g.oy) / g.h - 0.5) + 1;
    b.k0 = (int)floor((zmin - g.oz) / g.h - 0.5) - 1, b.k1 = (int)ceil((zmax - g.oz) / g.h - 0.5) + 1;
    if (b.j0 < 0) b.j0 = 0;
    if (b.j1 > g.ny - 1) b.j1 = g.ny - 1;
    if (b.k0 < z_begin) b.k0 = z_begin;
    if (b.k1 > z_end - 1) b.k1 = z_end - 1;
    return b;
}

// one warp per triangle; its lanes share the rows of the bounding box
__global__ void k_vox_mark(const float *tri, long long ntri, Grid g, int z_begin, int z_end, int W, uint32_t *mark) {
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= ntri) return;
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = tri[9 * t + i];
    const RowBox b = row_box(v, g, z_begin, z_end);
    if (b.j1 < b.j0 || b.k1 < b.k0) return;
    const int nj = b.j1 - b.j0 + 1;
    const long long rows = (long long)nj * (b.k1 - b.k0 + 1);
    for (long long r = lane; r < rows; r += 32) {
        const int k = b.k0 + (int)(r / nj), j = b.j0 + (int)(r % nj);
        const double py = g.oy + ((double)j + 0.5) * g.h, pz = g.oz + ((double)k + 0.5) * g.h;
        double xc;
        if (!ray_hits(v, py, pz, xc)) continue;
        const double fi = floor((xc - g.ox) / g.h - 0.5) + 1.0;  // first voxel whose centre lies beyond the crossing
        if (fi >= (double)g.nx) continue;
        const int i0 = fi < 0.0 ? 0 : (int)fi;
        atomicXor(mark + ((size_t)(k - z_begin) * g.ny + j) * W + (i0 >> 5), 1u << (i0 & 31));
    }
}

__global__ void k_vox_fill(const uint32_t *mark, Grid g, long long nrows, int W, uint8_t *out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    // a row with an odd number of crossings (ray through the ragged rim of an open end) is cleared
    unsigned total = 0;
    for (int w = lane; w < W; w += 32) total ^= (unsigned)__popc(mark[(size_t)row * W + w]) & 1u;
    total = __reduce_xor_sync(0xffffffffu, total);
    unsigned carry = 0;  // parity of all marker bits of the words before this group
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        uint32_t m = w < W ? mark[(size_t)row * W + w] : 0u;
        // inclusive prefix XOR inside the word
        m ^= m << 1, m ^= m << 2, m ^= m << 4, m ^= m << 8, m ^= m << 16;
        // exclusive prefix XOR of the word parities (= bit 31 of the in-word prefix) across the lanes
        unsigned p = m >> 31, incl = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl ^= up;
        }
        const unsigned before = (incl ^ p) ^ carry;
        if (before) m = ~m;
        if (total) m = 0u;
        carry ^= __shfl_sync(0xffffffffu, incl, 31);
        if (w < W) {
            uint8_t *o = out + (size_t)row * g.nx + (size_t)w * 32;
            const int n = g.nx - w * 32 < 32 ? g.nx - w * 32 : 32;
            for (int i = 0; i < n; i++) o[i] = (uint8_t)((m >> i) & 1u);
        }
    }
}

int fail(int code, const std::string &msg) {
    g_vox_error = msg;
    return code;
}

#define VCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            rc = fail(e_ == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA,                          \
                      std::string(#call) + " failed: " + cudaGetErrorString(e_));                              \
            goto done;                                                                                         \
        }                                                                                                      \
    } while (0)

// binary STL ("solid" ASCII files are recognised by their size mismatch and parsed as text)
int read_stl(const char *path, std::vector<float> &tri) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(LBM_ERR_IO, std::string("cannot open '") + path + "'");
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (raw.size() >= 84) {
        uint32_t n;
        memcpy(&n, raw.data() + 80, 4);
        if (84 + 50ull * n == raw.size()) {
            tri.resize(9ull * n);
            for (uint32_t t = 0; t < n; t++) memcpy(&tri[9ull * t], raw.data() + 84 + 50ull * t + 12, 36);
            return 0;
        }
    }
    // ASCII: every "vertex x y z" line
    tri.clear();
    std::string text(raw.begin(), raw.end());
    size_t pos = 0;
    while ((pos = text.find("vertex", pos)) != std::string::npos) {
        float x, y, z;
        if (sscanf(text.c_str() + pos + 6, "%f %f %f", &x, &y, &z) != 3) return fail(LBM_ERR_IO, std::string("bad vertex line in '") + path + "'");
        tri.push_back(x), tri.push_back(y), tri.push_back(z);
        pos += 6;
    }
    if (tri.empty() || tri.size() % 9) return fail(LBM_ERR_IO, std::string("'") + path + "' is neither a binary nor an ASCII STL");
    return 0;
}

}  // namespace

extern "C" {

const char *lbm_voxel_last_error(void) { return g_vox_error.c_str(); }

int lbm_voxelize_triangles(const float *tri, int64_t ntri, const lbm_voxel_grid *grid, int32_t z_begin, int32_t z_end,
                           int32_t device, uint8_t *flag_out) {
    if (!grid || !flag_out || (ntri > 0 && !tri)) return fail(LBM_ERR_ARG, "null argument");
    if (grid->nx <= 0 || grid->ny <= 0 || grid->nz <= 0 || !(grid->spacing > 0.0)) return fail(LBM_ERR_ARG, "bad grid");
    if (z_begin < 0 || z_end > grid->nz || z_end <= z_begin) return fail(LBM_ERR_ARG, "bad z range");
    Grid g{grid->origin[0], grid->origin[1], grid->origin[2], grid->spacing, grid->nx, grid->ny, grid->nz};
    const int W = (g.nx + 31) / 32;
    const long long nrows = (long long)(z_end - z_begin) * g.ny;
    int rc = 0;
    float *d_tri = nullptr;
    uint32_t *d_mark = nullptr;
    uint8_t *d_out = nullptr;
    VCK(cudaSetDevice(device));
    VCK(cudaMalloc((void **)&d_mark, (size_t)nrows * W * sizeof(uint32_t)));
    VCK(cudaMemset(d_mark, 0, (size_t)nrows * W * sizeof(uint32_t)));
    VCK(cudaMalloc((void **)&d_out, (size_t)nrows * g.nx));
    if (ntri > 0) {
        VCK(cudaMalloc((void **)&d_tri, (size_t)ntri * 9 * sizeof(float)));
        VCK(cudaMemcpy(d_tri, tri, (size_t)ntri * 9 * sizeof(float), cudaMemcpyHostToDevice));
        const long long threads = ntri * 32;
        k_vox_mark<<<(unsigned)((threads + 255) / 256), 256>>>(d_tri, ntri, g, z_begin, z_end, W, d_mark);
        VCK(cudaGetLastError());
    }
    k_vox_fill<<<(unsigned)((nrows * 32 + 255) / 256), 256>>>(d_mark, g, nrows, W, d_out);
    VCK(cudaGetLastError());
    VCK(cudaMemcpy(flag_out, d_out, (size_t)nrows * g.nx, cudaMemcpyDeviceToHost));
done:
    cudaFree(d_tri), cudaFree(d_mark), cudaFree(d_out);
    return rc;
}

int lbm_voxelize_stl(const char *stl_path, const lbm_voxel_grid *grid, int32_t z_begin, int32_t z_end, int32_t device,
                     uint8_t *flag_out) {
    if (!stl_path) return fail(LBM_ERR_ARG, "null path");
    std::vector<float> tri;
    int rc = read_stl(stl_path, tri);
    if (rc) return rc;
    return lbm_voxelize_triangles(tri.data(), (int64_t)(tri.size() / 9), grid, z_begin, z_end, device, flag_out);
}

}  // extern "C"
