// G1 (Fp) instantiation of the MSM back end; see msm_impl.cuh.
#include "msm_impl.cuh"

namespace g16 {

template class MsmAccumulator<Fp>;
template class Msm<Fp>;
template void msm_build_table<Fp>(const Affine<Fp>*, size_t, size_t, int, Affine<Fp>*, cudaStream_t);
template void xyzz_sum_to_affine<Fp>(const XYZZ<Fp>*, int, Affine<Fp>*, cudaStream_t);

}  // namespace g16
