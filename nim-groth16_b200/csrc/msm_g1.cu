// G1 (Fp) instantiation of the MSM pipeline; see msm_impl.cuh.
#include "msm_impl.cuh"

namespace g16 {

int msm_pick_window(size_t n, bool g2) {
  (void)g2;
  double best = 1e300;
  int best_c = 4;
  for (int c = 4; c <= 18; c++) {
    double W = (double)msm_num_windows(c);
    double nb = (double)((size_t)1 << (c - 1));
    // mixed adds + running-sum adds (two full adds per bucket) + a small per-window latency term
    double cost = W * (10.0 * (double)n + 28.0 * nb) + W * 4000.0;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

template class Msm<Fp>;
template void xyzz_sum_to_affine<Fp>(const XYZZ<Fp>*, int, Affine<Fp>*, cudaStream_t);
template void affine_sum_to_xyzz<Fp>(const Affine<Fp>*, int, XYZZ<Fp>*, cudaStream_t);

}  // namespace g16
