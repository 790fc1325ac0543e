// G2 (Fp2) instantiation of the MSM pipeline; see msm_impl.cuh.
#include "msm_impl.cuh"

namespace g16 {

template class Msm<Fp2>;
template void xyzz_sum_to_affine<Fp2>(const XYZZ<Fp2>*, int, Affine<Fp2>*, cudaStream_t);
template void affine_sum_to_xyzz<Fp2>(const Affine<Fp2>*, int, XYZZ<Fp2>*, cudaStream_t);

}  // namespace g16
