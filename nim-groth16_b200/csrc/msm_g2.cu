// G2 (Fp2) instantiation of the MSM back end; see msm_impl.cuh.
// One out-of-line call per Fp2 multiply / square (field.cuh): 6.81 vs 6.88 ms for the 2^20 accumulation.
#define G16_FP2_WHOLE_CALL
#include "msm_impl.cuh"

namespace g16 {

template class MsmAccumulator<Fp2>;
template class Msm<Fp2>;
template void msm_build_table<Fp2>(const Affine<Fp2>*, size_t, size_t, int, Affine<Fp2>*, cudaStream_t);
template void xyzz_sum_to_affine<Fp2>(const XYZZ<Fp2>*, int, Affine<Fp2>*, cudaStream_t);

}  // namespace g16
