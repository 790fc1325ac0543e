// Library-internal interface of ntt.cu (device pointers, asynchronous on `stream`).
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace g16 {

// builds (once per device and size) the twiddle / coset tables for a 2^log_n domain
void ntt_prepare(int log_n, cudaStream_t stream);
void ntt_release_tables();

// natural-order in, natural-order out (ntt.nim:55 forwardNTT / ntt.nim:139 inverseNTT incl. 1/n).
// `work` is n elements of scratch (may alias `in`); `out` must be distinct from both.
void ntt_natural(const Fr* in, Fr* out, Fr* work, int log_n, bool inverse, cudaStream_t stream);

// abc = [Az | Bz | n scratch elements]; clobbers abc, writes the n H-scalars to qs.
// flavour 0 = JensGroth (prover.nim:118-148), 1 = Snarkjs (prover.nim:158-181).
void quotient(Fr* abc, Fr* qs, int log_n, int flavour, cudaStream_t stream);

void pointwise_mul(const Fr* a, const Fr* b, Fr* c, size_t n, cudaStream_t stream);

}  // namespace g16
