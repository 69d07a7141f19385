// Kernel-variant microbenchmark for the fused D3Q19 step (development tool, not the product).
// Times stripped-down variants of the bulk path (no boundary epilogue, all nodes fluid in the
// interior) on a dense N^3 box so design choices can be measured in one gpurun call:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/kbench.cu -o gpurun_out/kbench
//   ./kbench [N=512] [fp64=1]
// Prints GB/s of algorithmic traffic (2*19*sizeof(T) per interior node) for each variant.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int Q = 19;
__host__ __device__ constexpr int cxq(int q) { constexpr int a[Q] = {0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, 0, 0, 0, 0}; return a[q]; }
__host__ __device__ constexpr int cyq(int q) { constexpr int a[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return a[q]; }
__host__ __device__ constexpr int czq(int q) { constexpr int a[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, 1, -1, -1}; return a[q]; }

template <typename T>
__device__ __forceinline__ void collide(T (&f)[Q], T inv_tau) {
    T a0 = f[1] + f[2], a1 = f[3] + f[4], a2 = f[5] + f[6];
    T d0 = f[7] + f[10], d1 = f[8] + f[9], d2 = f[11] + f[14], d3 = f[12] + f[13], d4 = f[15] + f[18], d5 = f[16] + f[17];
    T r = ((f[0] + a0) + (a1 + a2)) + ((d0 + d1) + (d2 + d3)) + (d4 + d5);
    T inv = T(1.0) / r;
    T ux = (((f[1] - f[2]) + (f[7] - f[10])) + ((f[8] - f[9]) + (f[11] - f[14])) + (f[12] - f[13])) * inv;
    T uy = (((f[3] - f[4]) + (f[7] - f[10])) + ((f[9] - f[8]) + (f[15] - f[18])) + (f[17] - f[16])) * inv;
    T uz = (((f[5] - f[6]) + (f[11] - f[14])) + ((f[13] - f[12]) + (f[15] - f[18])) + (f[16] - f[17])) * inv;
    const T om = inv_tau, om1 = T(1.0) - inv_tau;
    const T base = T(1.0) - T(1.5) * (ux * ux + uy * uy + uz * uz);
    const T k0 = om * r * T(1.0 / 3.0), k1 = om * r * T(1.0 / 18.0), k2 = om * r * T(1.0 / 36.0);
    f[0] = om1 * f[0] + k0 * base;
#define PAIR(qp, qm, cu, kw) { T cu_ = (cu); T ev = base + T(4.5) * cu_ * cu_; T od = T(3.0) * cu_; f[qp] = om1 * f[qp] + (kw) * (ev + od); f[qm] = om1 * f[qm] + (kw) * (ev - od); }
    PAIR(1, 2, ux, k1) PAIR(3, 4, uy, k1) PAIR(5, 6, uz, k1)
    PAIR(7, 10, ux + uy, k2) PAIR(8, 9, ux - uy, k2) PAIR(11, 14, ux + uz, k2) PAIR(12, 13, ux - uz, k2)
    PAIR(15, 18, uy + uz, k2) PAIR(17, 16, uy - uz, k2)
#undef PAIR
}

struct P {
    long long qs;      // q stride
    long long c0, c1;  // cell range
    int px;
    long long plane;
};

template <typename T> __device__ __forceinline__ T ld_plain(const T *p) { return *p; }
template <typename T> __device__ __forceinline__ T ld_nc(const T *p) { return __ldg(p); }
__device__ __forceinline__ double ld_stream_(const double *p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ld_stream_(const float *p) { float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ void st_cs(double *p, double v) { asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ void st_cs(float *p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v)); }

// ---- copy ceilings
template <typename T>
__global__ void __launch_bounds__(256) k_copy19(const T *__restrict__ src, T *__restrict__ dst, P p) {
    long long c = p.c0 + (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= p.c1) return;
    T f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = src[q * p.qs + c];
#pragma unroll
    for (int q = 0; q < Q; q++) dst[q * p.qs + c] = f[q];
}
template <typename T>
__global__ void __launch_bounds__(256) k_copy19_shift(const T *__restrict__ src, T *__restrict__ dst, P p) {
    long long c = p.c0 + (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= p.c1) return;
    T f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = src[q * p.qs + c - (cxq(q) + (long long)p.px * cyq(q) + p.plane * czq(q))];
#pragma unroll
    for (int q = 0; q < Q; q++) dst[q * p.qs + c] = f[q];
}
__global__ void __launch_bounds__(256) k_copy1(const double2 *__restrict__ src, double2 *__restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- V0: baseline of the product: one node per thread, 256 threads, __ldg
template <typename T, int BLOCK, int MINB, int LD, bool STCS>
__global__ void __launch_bounds__(BLOCK, MINB) k_step(const T *__restrict__ src, T *__restrict__ dst, P p, T inv_tau) {
    long long c = p.c0 + (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= p.c1) return;
    T f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const T *a = src + q * p.qs + c - (cxq(q) + (long long)p.px * cyq(q) + p.plane * czq(q));
        f[q] = LD == 0 ? ld_plain(a) : (LD == 1 ? ld_nc(a) : ld_stream_(a));
    }
    collide<T>(f, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) {
        if (STCS) st_cs(dst + q * p.qs + c, f[q]);
        else dst[q * p.qs + c] = f[q];
    }
}

// ---- V4: two nodes per thread, 128-bit (fp64) accesses; x-shifted populations via aligned load + shuffle
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_step_v2(const double *__restrict__ src, double *__restrict__ dst, P p, double inv_tau) {
    // each thread: cells c, c+1 (c even); a warp covers 64 consecutive cells of one row
    long long c = p.c0 + ((long long)blockIdx.x * BLOCK + threadIdx.x) * 2;
    if (c >= p.c1) return;
    const int lane = threadIdx.x & 31;
    double f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const long long off = (long long)p.px * cyq(q) + p.plane * czq(q);
        const double *row = src + q * p.qs + c - off;  // aligned pair of this thread in the source row
        double2 v = *reinterpret_cast<const double2 *>(row);
        if (cxq(q) == 0) {
            f0[q] = v.x, f1[q] = v.y;
        } else if (cxq(q) == 1) {  // need row[-1], row[0]
            double up = __shfl_up_sync(0xffffffffu, v.y, 1);
            if (lane == 0) up = row[-1];
            f0[q] = up, f1[q] = v.x;
        } else {  // need row[1], row[2]
            double dn = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31) dn = row[2];
            f0[q] = v.y, f1[q] = dn;
        }
    }
    collide<double>(f0, inv_tau);
    collide<double>(f1, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) *reinterpret_cast<double2 *>(dst + q * p.qs + c) = make_double2(f0[q], f1[q]);
}

// ---- two rows per thread: cells c and c + px, scalar coalesced accesses, 38 loads in flight per thread
template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_step_2rows(const T *__restrict__ src, T *__restrict__ dst, P p, T inv_tau) {
    // block covers BLOCK consecutive cells of row pair (2r, 2r+1)
    const long long per_pair = 2LL * p.px;
    const long long idx = (long long)blockIdx.x * BLOCK + threadIdx.x;  // index among first-row cells
    const long long pair = idx / p.px, x = idx - pair * p.px;
    const long long c = p.c0 + pair * per_pair + x;
    if (c + p.px >= p.c1) return;
    T f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const T *a = src + q * p.qs + c - (cxq(q) + (long long)p.px * cyq(q) + p.plane * czq(q));
        f0[q] = __ldg(a);
        f1[q] = __ldg(a + p.px);
    }
    collide<T>(f0, inv_tau);
    collide<T>(f1, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) {
        dst[q * p.qs + c] = f0[q];
        dst[q * p.qs + c + p.px] = f1[q];
    }
}

// ---- AA pattern: even step (purely local, aligned) and odd step (shifted reads AND writes), in place
template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_aa_even(T *__restrict__ f_, P p, T inv_tau) {
    long long c = p.c0 + (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= p.c1) return;
    constexpr int opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    T f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = f_[q * p.qs + c];
    collide<T>(f, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) f_[opp[q] * p.qs + c] = f[q];
}
template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_aa_odd(T *__restrict__ f_, P p, T inv_tau) {
    long long c = p.c0 + (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= p.c1) return;
    constexpr int opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    T f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = f_[opp[q] * p.qs + c - (cxq(q) + (long long)p.px * cyq(q) + p.plane * czq(q))];
    collide<T>(f, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) f_[q * p.qs + c + (cxq(q) + (long long)p.px * cyq(q) + p.plane * czq(q))] = f[q];
}

// ---- two cells per thread, 64-bit (fp32) / 128-bit (fp64) accesses: half the memory requests per byte
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
template <typename T> __device__ __forceinline__ typename Vec2<T>::type mk2(T a, T b) { typename Vec2<T>::type v; v.x = a, v.y = b; return v; }

template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_copy19_v2(const T *__restrict__ src, T *__restrict__ dst, P p) {
    using V = typename Vec2<T>::type;
    long long c = p.c0 + ((long long)blockIdx.x * BLOCK + threadIdx.x) * 2;
    if (c >= p.c1) return;
    V f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = *reinterpret_cast<const V *>(src + q * p.qs + c);
#pragma unroll
    for (int q = 0; q < Q; q++) *reinterpret_cast<V *>(dst + q * p.qs + c) = f[q];
}
template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_aa_even_v2(T *__restrict__ f_, P p, T inv_tau) {
    using V = typename Vec2<T>::type;
    long long c = p.c0 + ((long long)blockIdx.x * BLOCK + threadIdx.x) * 2;
    if (c >= p.c1) return;
    constexpr int opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    T f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        V v = *reinterpret_cast<const V *>(f_ + q * p.qs + c);
        f0[q] = v.x, f1[q] = v.y;
    }
    collide<T>(f0, inv_tau);
    collide<T>(f1, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) *reinterpret_cast<V *>(f_ + opp[q] * p.qs + c) = mk2<T>(f0[q], f1[q]);
}
template <typename T, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_aa_odd_v2(T *__restrict__ f_, P p, T inv_tau) {
    using V = typename Vec2<T>::type;
    long long c = p.c0 + ((long long)blockIdx.x * BLOCK + threadIdx.x) * 2;
    if (c >= p.c1) return;  // whole warps only (n is a multiple of 64 here)
    const int lane = threadIdx.x & 31;
    constexpr int opp[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    T f0[Q], f1[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const long long off = (long long)p.px * cyq(q) + p.plane * czq(q);
        const T *row = f_ + opp[q] * p.qs + c - off;  // this thread's aligned pair in the source row
        V v = *reinterpret_cast<const V *>(row);
        if (cxq(q) == 0) {
            f0[q] = v.x, f1[q] = v.y;
        } else if (cxq(q) == 1) {  // cells c, c+1 pull from c-1, c
            T up = __shfl_up_sync(0xffffffffu, v.y, 1);
            if (lane == 0) up = row[-1];
            f0[q] = up, f1[q] = v.x;
        } else {  // from c+1, c+2
            T dn = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31) dn = row[2];
            f0[q] = v.y, f1[q] = dn;
        }
    }
    collide<T>(f0, inv_tau);
    collide<T>(f1, inv_tau);
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const long long off = (long long)p.px * cyq(q) + p.plane * czq(q);
        T *row = f_ + q * p.qs + c + off;  // aligned pair in the target row
        if (cxq(q) == 0) {
            *reinterpret_cast<V *>(row) = mk2<T>(f0[q], f1[q]);
        } else if (cxq(q) == 1) {  // values go to c+1, c+2: the aligned pair (c, c+1) gets (left neighbour's f1, own f0)
            const T left = __shfl_up_sync(0xffffffffu, f1[q], 1);
            if (lane == 0) row[1] = f0[q];
            else *reinterpret_cast<V *>(row) = mk2<T>(left, f0[q]);
            if (lane == 31) row[2] = f1[q];
        } else {  // values go to c-1, c: the aligned pair (c, c+1) gets (own f1, right neighbour's f0)
            const T right = __shfl_down_sync(0xffffffffu, f0[q], 1);
            if (lane == 31) row[0] = f1[q];
            else *reinterpret_cast<V *>(row) = mk2<T>(f1[q], right);
            if (lane == 0) row[-1] = f0[q];
        }
    }
}

template <typename F>
float time_it(F launch, int reps = 10) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) launch(i + 3);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

template <typename T>
__global__ void k_fill(T *f, long long n, long long qs) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double w[Q] = {1 / 3., 1 / 18., 1 / 18., 1 / 18., 1 / 18., 1 / 18., 1 / 18., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36., 1 / 36.};
    for (int q = 0; q < Q; q++) f[q * qs + i] = (T)(w[q] * (1.0 + 1e-3 * ((i * 7 + q) % 13)));
}

static int g_only = 0;  // 1: only the AA 1-cell/thread b128 minb6 pair (for ncu)
template <typename T>
void run(int N, long long qpad) {
    const int px = ((N + 31) / 32) * 32;
    P p;
    p.px = px, p.plane = (long long)px * N;
    const long long cells = p.plane * N;
    p.qs = cells + qpad;
    p.c0 = p.plane, p.c1 = cells - p.plane;  // interior planes only: every access stays in range
    T *a, *b;
    CK(cudaMalloc(&a, sizeof(T) * p.qs * Q));
    CK(cudaMalloc(&b, sizeof(T) * p.qs * Q));
    k_fill<T><<<(unsigned)((cells + 255) / 256), 256>>>(a, cells, p.qs);
    k_fill<T><<<(unsigned)((cells + 255) / 256), 256>>>(b, cells, p.qs);
    CK(cudaDeviceSynchronize());
    const long long n = p.c1 - p.c0;
    const double gb = (double)n * 2 * Q * sizeof(T) / 1e9;
    const T it = (T)(1.0 / 0.55);
    auto rep = [&](const char *name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s  %8.1f MLUPS\n", name, ms, gb / (ms * 1e-3), n / (ms * 1e-3) / 1e6); fflush(stdout); };
    printf("N=%d %s  qstride pad=%lld elements  cells/launch=%lld  (%.2f GB algorithmic)\n", N, sizeof(T) == 8 ? "fp64" : "fp32", qpad, n, gb);
    unsigned g256 = (unsigned)((n + 255) / 256), g128 = (unsigned)((n + 127) / 128), g512 = (unsigned)((n + 511) / 512);
    if (g_only == 1) {
        rep("AA even+odd 1 cell/thread b128 minb6", time_it([&](int i) { if (i & 1) k_aa_odd<T, 128, 6><<<g128, 128>>>(a, p, it); else k_aa_even<T, 128, 6><<<g128, 128>>>(a, p, it); }));
        return;
    }
    {
        long long n2 = (long long)(sizeof(T) * p.qs * Q / sizeof(double2));
        float ms = time_it([&](int) { k_copy1<<<(unsigned)((n2 + 255) / 256), 256>>>((const double2 *)a, (double2 *)b, n2); });
        printf("%-44s %8.3f ms  %8.1f GB/s\n", "copy 1 stream double2 (whole buffer)", ms, 2.0 * n2 * 16 / 1e9 / (ms * 1e-3));
    }
    rep("copy19 (19 rd + 19 wr streams, no shift)", time_it([&](int i) { if (i & 1) k_copy19<T><<<g256, 256>>>(b, a, p); else k_copy19<T><<<g256, 256>>>(a, b, p); }));
    rep("copy19 shifted reads (pull, no math)", time_it([&](int i) { if (i & 1) k_copy19_shift<T><<<g256, 256>>>(b, a, p); else k_copy19_shift<T><<<g256, 256>>>(a, b, p); }));
#define STEP(NAME, BLOCK, MINB, LD, STCS, GRID) rep(NAME, time_it([&](int i) { if (i & 1) k_step<T, BLOCK, MINB, LD, STCS><<<GRID, BLOCK>>>(b, a, p, it); else k_step<T, BLOCK, MINB, LD, STCS><<<GRID, BLOCK>>>(a, b, p, it); }));
    STEP("step b256 minb1 ldg", 256, 1, 1, false, g256)
    STEP("step b256 minb2 ldg", 256, 2, 1, false, g256)
    STEP("step b256 minb3 ldg", 256, 3, 1, false, g256)
    STEP("step b256 minb4 ldg", 256, 4, 1, false, g256)
    STEP("step b128 minb4 ldg", 128, 4, 1, false, g128)
    STEP("step b128 minb6 ldg", 128, 6, 1, false, g128)
    STEP("step b512 minb1 ldg", 512, 1, 1, false, g512)
    STEP("step b256 minb2 plain ld", 256, 2, 0, false, g256)
    STEP("step b256 minb2 ld no_allocate", 256, 2, 2, false, g256)
    STEP("step b256 minb2 ldg + st.cs", 256, 2, 1, true, g256)
    STEP("step b256 minb3 ld no_allocate + st.cs", 256, 3, 2, true, g256)
    if constexpr (sizeof(T) == 8) {
        unsigned gv = (unsigned)((n / 2 + 127) / 128);
        rep("step 2 nodes/thread 128-bit b128 minb2", time_it([&](int i) { if (i & 1) k_step_v2<128, 2><<<gv, 128>>>((const double *)b, (double *)a, p, it); else k_step_v2<128, 2><<<gv, 128>>>((const double *)a, (double *)b, p, it); }));
        rep("step 2 nodes/thread 128-bit b128 minb3", time_it([&](int i) { if (i & 1) k_step_v2<128, 3><<<gv, 128>>>((const double *)b, (double *)a, p, it); else k_step_v2<128, 3><<<gv, 128>>>((const double *)a, (double *)b, p, it); }));
    }
    {
        unsigned g2 = (unsigned)((n / 2 + 127) / 128), g2b = (unsigned)((n / 2 + 255) / 256);
        rep("step 2 rows/thread b128 minb3", time_it([&](int i) { if (i & 1) k_step_2rows<T, 128, 3><<<g2, 128>>>(b, a, p, it); else k_step_2rows<T, 128, 3><<<g2, 128>>>(a, b, p, it); }));
        rep("step 2 rows/thread b128 minb4", time_it([&](int i) { if (i & 1) k_step_2rows<T, 128, 4><<<g2, 128>>>(b, a, p, it); else k_step_2rows<T, 128, 4><<<g2, 128>>>(a, b, p, it); }));
        rep("step 2 rows/thread b128 minb6", time_it([&](int i) { if (i & 1) k_step_2rows<T, 128, 6><<<g2, 128>>>(b, a, p, it); else k_step_2rows<T, 128, 6><<<g2, 128>>>(a, b, p, it); }));
        rep("step 2 rows/thread b256 minb2", time_it([&](int i) { if (i & 1) k_step_2rows<T, 256, 2><<<g2b, 256>>>(b, a, p, it); else k_step_2rows<T, 256, 2><<<g2b, 256>>>(a, b, p, it); }));
        STEP("step b128 minb8 ldg", 128, 8, 1, false, g128)
        STEP("step b128 minb10 ldg", 128, 10, 1, false, g128)
    }
    rep("AA even (local rd/wr, in place)", time_it([&](int) { k_aa_even<T, 256, 2><<<g256, 256>>>(a, p, it); }));
    rep("AA odd (shifted rd + shifted wr, in place)", time_it([&](int) { k_aa_odd<T, 256, 2><<<g256, 256>>>(a, p, it); }));
    rep("AA even+odd average", time_it([&](int i) { if (i & 1) k_aa_odd<T, 256, 2><<<g256, 256>>>(a, p, it); else k_aa_even<T, 256, 2><<<g256, 256>>>(a, p, it); }));
    {
        unsigned h128 = (unsigned)((n / 2 + 127) / 128), h256 = (unsigned)((n / 2 + 255) / 256);
        rep("copy19, 2 cells/thread vector accesses", time_it([&](int i) { if (i & 1) k_copy19_v2<T, 256, 2><<<h256, 256>>>(b, a, p); else k_copy19_v2<T, 256, 2><<<h256, 256>>>(a, b, p); }));
#define AA2(NAME, BLOCK, MINB, GRID)                                                                                          \
    rep("AA even 2 cells/thread " NAME, time_it([&](int) { k_aa_even_v2<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); }));      \
    rep("AA odd  2 cells/thread " NAME, time_it([&](int) { k_aa_odd_v2<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); }));       \
    rep("AA even+odd 2 cells/thread " NAME, time_it([&](int i) { if (i & 1) k_aa_odd_v2<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); else k_aa_even_v2<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); }));
        AA2("b128 minb3", 128, 3, h128)
        AA2("b128 minb4", 128, 4, h128)
        AA2("b128 minb5", 128, 5, h128)
        AA2("b128 minb6", 128, 6, h128)
        AA2("b256 minb2", 256, 2, h256)
        AA2("b256 minb3", 256, 3, h256)
#define AA1(NAME, BLOCK, MINB, GRID) rep("AA even+odd 1 cell/thread " NAME, time_it([&](int i) { if (i & 1) k_aa_odd<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); else k_aa_even<T, BLOCK, MINB><<<GRID, BLOCK>>>(a, p, it); }));
        AA1("b128 minb6", 128, 6, g128)
        AA1("b128 minb8", 128, 8, g128)
        AA1("b128 minb10", 128, 10, g128)
        AA1("b128 minb12", 128, 12, g128)
    }
    CK(cudaFree(a));
    CK(cudaFree(b));
}

int main(int argc, char **argv) {
    int N = argc > 1 ? atoi(argv[1]) : 512;
    int fp64 = argc > 2 ? atoi(argv[2]) : 1;
    long long qpad = argc > 3 ? atoll(argv[3]) : 64;
    g_only = argc > 4 ? atoi(argv[4]) : 0;
    if (fp64) run<double>(N, qpad);
    else run<float>(N, qpad);
    return 0;
}
