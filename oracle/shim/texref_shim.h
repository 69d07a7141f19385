/* Stand-in for the legacy texture-reference API (texture<>, tex1Dfetch,
 * cudaBindTexture, cudaUnbindTexture) that CUDA 12 removed and that
 * Poiseulle.cu:49-50, bifurcation.cu:27-34 and coronary.cu:28-29 still use.
 * Force-included (-include) when building oracle/_ref so the reference sources
 * compile UNMODIFIED from /root/reference.  A "texture" becomes a __device__
 * struct holding a plain pointer; a fetch is a read-only load.  Test
 * infrastructure only. */
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

template <class T, int Dim = 1, int Mode = 0>
struct texref_shim {
    const T *ptr;
};

template <class T, int D, int M>
__device__ __forceinline__ T tex1Dfetch(const texref_shim<T, D, M> &t, int i) {
    return __ldg(t.ptr + i);
}
template <class T, int D, int M>
inline cudaError_t cudaBindTexture(size_t *offset, texref_shim<T, D, M> &t, const void *devptr) {
    if (offset) *offset = 0;
    return cudaMemcpyToSymbol(t, &devptr, sizeof(devptr));
}
template <class T, int D, int M>
inline cudaError_t cudaUnbindTexture(texref_shim<T, D, M> &) {
    return cudaSuccess;
}
/* `texture<int,cudaTextureType1D,cudaReadModeElementType> name;` at file scope
 * now declares a device-side pointer holder. */
#define texture __device__ texref_shim
