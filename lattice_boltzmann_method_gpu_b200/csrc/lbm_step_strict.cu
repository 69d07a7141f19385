// STRICT arithmetic instantiation of the fused step.  This file is compiled with
// -fmad=false so that every expression keeps the reference's evaluation order
// with IEEE rounding after each operation -- the results are bit-identical to
// the CPU oracle (gcc -ffp-contract=off).  Used by the parity tests to separate
// "indexing / boundary logic" (must be exact) from "arithmetic reformulation"
// (FAST vs STRICT, tolerance).
#include "step_dense.cuh"

namespace lbm {
template <typename T>
cudaError_t launch_step_dense_strict(const StepParams<T> &p, bool moments, bool resid, int storage, cudaStream_t s) {
    return launch_step_dense_impl<T, true>(p, moments, resid, storage, s);
}
template cudaError_t launch_step_dense_strict<float>(const StepParams<float> &, bool, bool, int, cudaStream_t);
template cudaError_t launch_step_dense_strict<double>(const StepParams<double> &, bool, bool, int, cudaStream_t);
}  // namespace lbm

#include "step_sparse.cuh"
namespace lbm {
template <typename T>
cudaError_t launch_step_sparse_strict(const SparseParams<T> &p, bool moments, bool resid, cudaStream_t s) {
    return launch_step_sparse_impl<T, true>(p, moments, resid, s);
}
template cudaError_t launch_step_sparse_strict<float>(const SparseParams<float> &, bool, bool, cudaStream_t);
template cudaError_t launch_step_sparse_strict<double>(const SparseParams<double> &, bool, bool, cudaStream_t);
}  // namespace lbm

#include "step_sparse_aa.cuh"
namespace lbm {
template <typename T>
cudaError_t launch_step_sparse_aa_strict(const SparseParams<T> &p, bool moments, bool resid, cudaStream_t s) {
    return launch_step_sparse_aa_impl<T, true>(p, moments, resid, s);
}
template cudaError_t launch_step_sparse_aa_strict<float>(const SparseParams<float> &, bool, bool, cudaStream_t);
template cudaError_t launch_step_sparse_aa_strict<double>(const SparseParams<double> &, bool, bool, cudaStream_t);
}  // namespace lbm
