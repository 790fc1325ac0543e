// Library-internal interface of abc.cu: sparse A.w / B.w evaluation (prover.nim:56-73 buildABC).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "field.cuh"

namespace g16 {

// coefficient record formats accepted at the boundary (see include/g16b200.h)
enum { COEFF_PACKED44_R2 = 0, COEFF_STRUCT48_MONT = 1 };

// Sorted sparse structure on the device: entries grouped by `key`, ptr[k] .. ptr[k+1].
struct SparseCsr {
  DevBuf ptr;    // nkeys + 1 (u32)
  DevBuf other;  // nnz (u32): the index that is not the key (column for buildABC)
  DevBuf vals;   // nnz Fr
  size_t nnz = 0;
  size_t nkeys = 0;
};

// zkey section-4 style coefficient list -> rows of A (keys 0..n-1) then rows of B (keys n..2n-1);
// values stored R^2-encoded so that montmul(value, standard-form witness) is the Montgomery product.
// Raises on a matrix-C entry (prover.nim:67) or an out-of-range row/column (zkey.nim:186-187).
void coeffs_to_csr(SparseCsr& out, const void* dev_records, size_t nnz, int format, int log_n, size_t nvars,
                   cudaStream_t stream);

// generic COO -> CSR by key (used by the fake setup: key = column)
void coo_to_csr(SparseCsr& out, const uint32_t* dev_keys, const uint32_t* dev_other, const Fr* dev_vals, size_t nnz,
                size_t nkeys, cudaStream_t stream);

// abc = [Az | Bz | Cz], each n = 2^log_n elements (Montgomery form); witness in standard form.
void build_abc(const SparseCsr& csr, const Fr* witness_std, Fr* abc, int log_n, cudaStream_t stream);

// elementwise conversions
void fr_from_mont(const Fr* in, Fr* out, size_t n, cudaStream_t stream);
void fr_to_mont(const Fr* in, Fr* out, size_t n, cudaStream_t stream);
// in place: standard-form values >= r are reduced mod r (the reference's fromBig, io.nim:141-145)
void fr_reduce_std(Fr* x, size_t n, cudaStream_t stream);

}  // namespace g16
