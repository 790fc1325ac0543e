/* g16b200 -- C ABI of the B200 (sm_100a) Groth16 proving backend for codex-storage/nim-groth16.
 *
 * The reference (/root/reference, pure Nim) has no FFI of its own: the boundary is the set of Nim
 * procs on the hot path.  Every entry point below names the reference proc it replaces (file:line,
 * relative to the reference root); INTEGRATION.md shows the `importc` binding for each.
 *
 * Conventions:
 *   - Field elements are 32 little-endian bytes = 4 x uint64 limbs.  "mont" = Montgomery residue with
 *     R = 2^256 (constantine's in-memory Fr/Fp, and the point sections of a .zkey, io.nim:103-131);
 *     "std" = the plain integer (.wtns / .r1cs, io.nim:141-145).
 *   - G1 affine = x, y (2 x Fp, 64 bytes); G2 affine = x.c0, x.c1, y.c0, y.c1 (128 bytes); infinity is
 *     the all-zero encoding (curves.nim:49-50).  Points are always Montgomery.
 *   - All calls are synchronous; the caller owns every buffer; inputs are only read during the call.
 *   - Return value 0 = ok; non-zero = failure, text from g16_last_error() (the Nim shim raises
 *     AssertionDefect, matching the reference's assert()s, e.g. msm.nim:97, prover.nim:236).
 *   - One caller thread per context.  `nthreads` of the reference signatures is a host-side hint with
 *     no meaning on the GPU and is therefore absent here.
 *   - There is no CPU fallback: every function fails if no sm_100 device is usable.
 */
#ifndef G16B200_H
#define G16B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G16_OK 0
#define G16_ERR_ARG 1      /* precondition violated (the reference would raise AssertionDefect) */
#define G16_ERR_CUDA 2     /* CUDA runtime failure, or no usable device */

#define G16_FORM_MONT 0    /* scalars are Montgomery residues (Nim seq[Fr] payload) */
#define G16_FORM_STD 1     /* scalars are standard-form integers (.wtns payload) */

#define G16_FLAVOUR_JENSGROTH 0   /* zkey_types.nim:11 */
#define G16_FLAVOUR_SNARKJS 1     /* zkey_types.nim:12 */

#define G16_COEFF_PACKED44_R2 0   /* .zkey section 4 record: u32 m,row,col + 32 B value*R^2 (zkey.nim:169-188) */
#define G16_COEFF_STRUCT48_MONT 1 /* g16_coeff below: value*R (the reference's in-memory Coeff, zkey_types.nim:48-52) */

#define G16_MEM_HOST 0
#define G16_MEM_DEVICE 1

/* g16_zkey_view.flags */
#define G16_ZKEY_TRUSTED 1u   /* skip the on-curve validation of the prover points (io.nim:228-236 -> curves.nim:95-107
                                 mkG1/mkG2 asserts); default: every point is checked on the GPU at g16_ctx_create */
#define G16_ZKEY_ONE_SHOT 2u  /* the context will serve one or a few proofs (cli_main.nim:193-210): keep the plain points
                                 instead of building the 13x window tables -- context creation is an upload, the MSMs
                                 run in the reference-shaped plain layout (about 2x slower per proof) */

const char* g16_last_error(void);
/* ABI version, device selection (one process per GPU; default device = cudaGetDevice()). */
int g16_version(void);               /* 2: g16_zkey_view.flags, 400-byte g16_partials, g16_shard_plan */
int g16_set_device(int device);
int g16_device_count(int* count);
/* Page-locks a caller-owned host range (for example the read-only mmap of a .zkey, files/zkey.nim:196-224) so that
 * the uploads of g16_ctx_create / the witness copies of g16_prove run as DMA transfers at PCIe speed instead of
 * staged pageable copies.  Optional; undo with g16_host_unregister before unmapping. */
int g16_host_register(const void* ptr, size_t bytes);
int g16_host_unregister(const void* ptr);
/* Device memory freed by destroyed contexts stays cached in the library's stream-ordered pools (a context per proof
 * then costs no cudaMalloc / cudaFree); this returns the cached memory of every device to the driver.  Environment
 * G16_POOL_KEEP_MB=<n> bounds the cache instead (0 = keep nothing). */
int g16_release_cached_memory(void);

/* ------------------------------------------------------------------------------------------------
 * Fine-grained level: one call per reference proc, host buffers in and out.
 * ---------------------------------------------------------------------------------------------- */

/* msmMultiThreadedG1 (groth16/bn128/msm.nim:89-124) / msmConstantineG1 (msm.nim:35-59).
 * out = sum_i scalars[i] * points[i], affine Montgomery, (0,0) for infinity.  n may be 0. */
int g16_msm_g1(const uint64_t* scalars, int scalar_form, const uint64_t* points, size_t n, uint64_t out[8]);

/* msmMultiThreadedG2 (msm.nim:128-158) / msmConstantineG2 (msm.nim:63-83). points: n x 16 limbs. */
int g16_msm_g2(const uint64_t* scalars, int scalar_form, const uint64_t* points, size_t n, uint64_t out[16]);

/* forwardNTT (groth16/math/ntt.nim:55-77) / inverseNTT (ntt.nim:139-161): natural order in and out,
 * Montgomery in and out, domain = createDomain(2^log_n) (domain.nim:28-46), inverse includes 1/n. */
int g16_ntt_fr(const uint64_t* in, uint64_t* out, int log_n, int inverse);

/* computeSnarkjsScalarCoeffs (groth16/prover.nim:158-181) for flavour = G16_FLAVOUR_SNARKJS,
 * computeQuotientPointwise (prover.nim:118-148) for G16_FLAVOUR_JENSGROTH.
 * az, bz: 2^log_n Montgomery elements each; Cz = Az o Bz is derived inside (prover.nim:69-71). */
int g16_quotient(const uint64_t* az, const uint64_t* bz, int log_n, int flavour, uint64_t* qs_out);

typedef struct g16_coeff {   /* mirrors Coeff (zkey_types.nim:48-52) */
  uint32_t matrix;           /* 0 = A, 1 = B; 2 (= C) is rejected like prover.nim:67 */
  uint32_t row;
  uint32_t col;
  uint32_t reserved;
  uint64_t value[4];         /* Montgomery */
} g16_coeff;

/* buildABC (groth16/prover.nim:56-73).  coeffs: nnz records of `coeff_format`; witness: m elements of
 * `witness_form`; az/bz/cz: 2^log_n Montgomery elements each. */
int g16_build_abc(const void* coeffs, size_t nnz, int coeff_format, const uint64_t* witness, int witness_form,
                  size_t m, int log_n, uint64_t* az, uint64_t* bz, uint64_t* cz);

/* ------------------------------------------------------------------------------------------------
 * Coarse level: resident prover context = generateProofWithMask (groth16/prover.nim:215-304).
 * ---------------------------------------------------------------------------------------------- */

typedef struct g16_zkey_view {        /* ZKey (zkey_types.nim:54-60) as raw arrays */
  uint32_t nvars;                     /* GrothHeader.nvars  (zkey_types.nim:19) */
  uint32_t npubs;                     /* GrothHeader.npubs  */
  uint32_t log_domain;                /* GrothHeader.logDomainSize */
  uint32_t flavour;                   /* G16_FLAVOUR_* */
  uint32_t coeff_format;              /* G16_COEFF_* */
  uint32_t mem_kind;                  /* G16_MEM_HOST / G16_MEM_DEVICE for the pointers below */
  uint32_t flags;                     /* G16_ZKEY_* */
  uint32_t reserved0;
  uint64_t ncoeffs;
  const void* coeffs;                 /* ZKey.coeffs */
  const uint64_t* points_a1;          /* nvars G1            (ProverPoints, zkey_types.nim:34-40) */
  const uint64_t* points_b1;          /* nvars G1 */
  const uint64_t* points_b2;          /* nvars G2 */
  const uint64_t* points_c1;          /* nvars - npubs - 1 G1 */
  const uint64_t* points_h1;          /* 2^log_domain G1 */
  uint64_t alpha1[8];                 /* SpecPoints (zkey_types.nim:24-31) */
  uint64_t beta1[8];
  uint64_t beta2[16];
  uint64_t delta1[8];
  uint64_t delta2[16];
} g16_zkey_view;

typedef struct g16_proof {            /* Proof (prover.nim:38-43) minus publicIO (= witness[0..npubs]) */
  uint64_t pi_a[8];
  uint64_t pi_b[16];
  uint64_t pi_c[8];
} g16_proof;

typedef struct g16_stats {            /* device-side milliseconds (CUDA events on the launch streams); for a multi-device
                                         context the maximum over its devices */
  float ms_h2d;                       /* witness upload */
  float ms_abc;                       /* "building 'ABC'"                    prover.nim:244 */
  float ms_quotient;                  /* "computing the quotient (FFTs)"     prover.nim:249 */
  float ms_sort_witness;              /* digit decomposition + sort shared by the four witness MSMs */
  float ms_msm_g1_witness;            /* pi_A, rho and the C part of pi_C (A1, B1, C1 fused)  prover.nim:279-288,302 */
  float ms_msm_b2;                    /* "computing pi_B (G2 MSM)"           prover.nim:291 */
  float ms_msm_h;                     /* the H part of "computing pi_C (2x G1 MSM)" incl. its sort  prover.nim:301 */
  float reserved_ms;
  float ms_assemble;
  float ms_total;
  uint32_t kernel_launches;           /* launches of this library's own kernels during the call */
  uint32_t reserved;
} g16_stats;

/* partial MSM results of one shard, affine Montgomery: A1, B1, H1, C1 (G1) and B2 (G2); infinity for an MSM the
 * shard owns no points of.  tag[0] = 1 when the record was produced after g16_ctx_set_mask (msm_c1 then holds
 * C_k + s*A_k + r*B1_k), tag[1] = a hash of that (r, s): g16_prove_finish refuses a mix of conventions. */
typedef struct g16_partials {
  uint64_t msm_a1[8];
  uint64_t msm_b1[8];
  uint64_t msm_h1[8];
  uint64_t msm_c1[8];
  uint64_t msm_b2[16];
  uint64_t tag[2];
} g16_partials;

typedef struct g16_ctx g16_ctx;

/* Uploads (once) the prover points and the coefficient list of a zkey, validates the points (unless
 * G16_ZKEY_TRUSTED) and builds the resident window tables (unless G16_ZKEY_ONE_SHOT).
 *   shard_count == 1, shard_index == 0   the whole key on the current device; g16_prove* work on it.  With the
 *                                        environment variable G16_NGPUS=N (N > 1) this is the next case.
 *   shard_count == -N (N >= 1)           the whole key spread over N devices of THIS process (devices
 *                                        shard_index .. shard_index+N-1, or the list in G16_DEVICES="0,1,..."), one
 *                                        shard each: g16_prove / g16_prove_submit / g16_prove_wait / g16_prove_dev work
 *                                        unchanged -- witness slices go to each device, the 400-byte partial records
 *                                        come back by peer copy ordered with events (no host synchronisation), device
 *                                        shard_index assembles.  This is msm.nim:96-124's internal parallelism
 *                                        behind the unchanged generateProofWithMask call.
 *   shard_count == G > 1                 shard shard_index of G on the current device (one process per GPU): use
 *                                        g16_prove_partials* / g16_prove_finish* and exchange the records yourself.
 * Which points a shard owns: g16_shard_plan. */
int g16_ctx_create(const g16_zkey_view* zkey, int shard_index, int shard_count, g16_ctx** out);
/* Another proof slot over the SAME resident key (tables, CSR rows, spec points are shared and read-only; only
 * the per-proof scratch and streams are new): use it with g16_prove_submit to keep several proofs in flight.
 * The key's device memory is released when the last context referring to it is destroyed. */
int g16_ctx_clone(g16_ctx* ctx, g16_ctx** out);
void g16_ctx_destroy(g16_ctx* ctx);

/* generateProofWithMask (prover.nim:215-304).  witness: nvars elements of `witness_form` (host);
 * r, s: the mask (prover.nim:211-213) as standard-form integers.  Requires shard_count == 1. */
int g16_prove(g16_ctx* ctx, const uint64_t* witness, int witness_form, const uint64_t r_std[4],
              const uint64_t s_std[4], g16_proof* proof, g16_stats* stats);

/* Asynchronous form of g16_prove: submit enqueues the whole proof on the context's streams and returns;
 * wait blocks until the proof is in host memory.  One proof in flight per context; several contexts (of the
 * same or of different zkeys) may be in flight at once, which overlaps the latency-bound tail of one proof
 * (bucket-reduction levels, assembly) with the accumulation kernels of the next.  The witness buffer must
 * stay valid until g16_prove_wait returns. */
int g16_prove_submit(g16_ctx* ctx, const void* witness, int witness_form, int witness_mem_kind,
                     const uint64_t r_std[4], const uint64_t s_std[4]);
int g16_prove_wait(g16_ctx* ctx, g16_proof* proof, g16_stats* stats);

/* Same with the witness already resident in device memory (standard form). */
int g16_prove_dev(g16_ctx* ctx, const void* witness_std_dev, const uint64_t r_std[4], const uint64_t s_std[4],
                  g16_proof* proof, g16_stats* stats);

/* Multi-GPU split of the same computation (msm.nim:107-119 across devices):
 *   every rank:  g16_prove_partials -> its five partial sums (device buffer of sizeof(g16_partials)),
 *   exchange:    all-gather of those 384-byte records (NCCL / peer copy, done by the host side),
 *   any rank:    g16_prove_finish over the gathered records -> the proof. */
/* The point ranges shard shard_index of shard_count owns, out = {a1_lo, a1_hi, b1_lo, b1_hi, c1_lo, c1_hi, b2_lo,
 * b2_hi, h_lo, h_hi}: a contiguous range (msm.nim:107-111) of each of the five MSMs -- witness indices for A1, B1,
 * B2 and C1 (C1[j - npubs - 1] multiplies witness[j]), domain indices for H.  Default policy: the MSMs themselves
 * are placed (whole MSMs, or a tail / head of one, per rank) by a cost model in nvars and the domain size, and
 * the ranks that own H points are the ones that run buildABC and the quotient.  Environment G16_SHARD_POLICY =
 * "uniform": the reference's equal chunks of every array; "g2own": the G2 MSM alone on the last shard.
 * Pure host arithmetic (no device needed); every rank computes the same plan. */
int g16_shard_plan(uint64_t nvars, uint64_t npubs, uint64_t domain_size, int shard_index, int shard_count,
                   uint64_t out[10]);
/* Optional, before g16_prove_partials*: announce the blinding scalars of the proof about to be computed.  The
 * rank then multiplies its OWN partial sums by them -- s * A_k + r * B1_k (prover.nim:298-299 by linearity), folded
 * into the c1 field of its record and overlapped with its remaining MSM work -- and g16_prove_finish*, called
 * with the same r, s, is additions and three affine conversions only.  Every rank of a proof must make the
 * same choice (all call it, or none). */
int g16_ctx_set_mask(g16_ctx* ctx, const uint64_t r_std[4], const uint64_t s_std[4]);
/* witness: the full nvars-element array; only the intervals the shard reads are copied (everything on a shard that
 * owns H points, else the ranges of its MSM pieces). */
int g16_prove_partials(g16_ctx* ctx, const uint64_t* witness, int witness_form, int witness_mem_kind,
                       void* partials_dev, g16_stats* stats);
/* asynchronous forms (see g16_prove_submit) */
int g16_prove_partials_submit(g16_ctx* ctx, const void* witness, int witness_form, int witness_mem_kind,
                              void* partials_dev);
int g16_prove_partials_wait(g16_ctx* ctx, g16_stats* stats);
int g16_prove_finish_submit(g16_ctx* ctx, const void* gathered_partials_dev, int count, const uint64_t r_std[4],
                            const uint64_t s_std[4]);     /* completed by g16_prove_wait */
/* Device-side ordering against a caller-owned stream (a cudaStream_t), so that the exchange between the partial sums
 * and the finish needs no host synchronisation: direction 0 makes `stream` wait for the record of the last
 * g16_prove_partials_submit; direction 1 makes the context wait for everything enqueued on `stream` so far (the
 * all-gather) before g16_prove_finish_submit's kernels run. */
int g16_ctx_order_stream(g16_ctx* ctx, void* stream, int direction);
/* Which layout the context ended up with (1 = resident window tables, 0 = plain points: G16_ZKEY_ONE_SHOT, or the tables
 * would not have fit the device -- environment G16_TABLE_BUDGET_MB overrides the budget of 80 % of the free memory) and
 * how much device memory it holds (all devices of a multi-device context). */
int g16_ctx_layout(g16_ctx* ctx, int* window_tables, uint64_t* device_bytes);
/* bytes of witness the last g16_prove* / g16_prove_partials* call copied to this context's device(s) */
int g16_ctx_last_witness_bytes(g16_ctx* ctx, uint64_t* bytes);
/* the five MSM sums of the most recent g16_prove* call on this context, as affine records (diagnostics) */
int g16_ctx_last_partials(g16_ctx* ctx, void* partials_dev);
int g16_prove_finish(g16_ctx* ctx, const void* gathered_partials_dev, int count, const uint64_t r_std[4],
                     const uint64_t s_std[4], g16_proof* proof);

/* ------------------------------------------------------------------------------------------------
 * Device-resident variants used by the benchmarks (inputs already in HBM; `stream` is a cudaStream_t).
 * ---------------------------------------------------------------------------------------------- */
typedef struct g16_msm_plan g16_msm_plan;   /* reusable workspace for one MSM shape */
int g16_msm_plan_create(int g2, size_t max_n, int window_bits /*0 = auto*/, g16_msm_plan** out);
void g16_msm_plan_destroy(g16_msm_plan* plan);
/* result_dev: XYZZ accumulator (128 B for G1, 256 B for G2) */
int g16_msm_dev(g16_msm_plan* plan, const void* scalars_dev, int scalar_form, const void* points_dev, size_t n,
                void* result_xyzz_dev, void* stream);
/* Resident-key layout used by g16_ctx: a table of 2^(c*w) * P_i for every window w, built once per point
 * array (table_dev: ceil(255/c) * n affine points), then MSMs against it share one bucket set. */
int g16_msm_plan_build_table(g16_msm_plan* plan, const void* points_dev, size_t n, void** table_dev_out);
int g16_msm_dev_table(g16_msm_plan* plan, const void* scalars_dev, int scalar_form, size_t n, void* result_xyzz_dev,
                      void* stream);
int g16_msm_result_to_affine(int g2, const void* result_xyzz_dev, int count, uint64_t* out_host);
int g16_msm_plan_info(const g16_msm_plan* plan, int* window_bits, int* num_windows, size_t* workspace_bytes);
/* per-kernel timing of the dominant kernel (bucket accumulation) with CUDA events on the launch stream */
int g16_msm_plan_profile(g16_msm_plan* plan, int enable);
int g16_msm_plan_last_profile(const g16_msm_plan* plan, float* accumulate_ms, float* total_ms, uint64_t* pairs);
/* device timer on the context's main stream: brackets every kernel and copy issued for the context */
int g16_ctx_timer_start(g16_ctx* ctx);
int g16_ctx_timer_stop(g16_ctx* ctx, float* elapsed_ms);

/* in_dev/out_dev/work_dev: 2^log_n elements each (out distinct from the others) */
int g16_ntt_fr_dev(const void* in_dev, void* out_dev, void* work_dev, int log_n, int inverse, void* stream);
/* abc_dev: 3 * 2^log_n elements [Az | Bz | scratch], clobbered; qs_dev: 2^log_n elements */
int g16_quotient_dev(void* abc_dev, void* qs_dev, int log_n, int flavour, void* stream);
int g16_ntt_prepare(int log_n);

/* ------------------------------------------------------------------------------------------------
 * Fake trusted setup on the GPU (groth16/fake_setup.nim:201-326 fakeCircuitSetup) -- fixture generator
 * and `--setup` backend.  Matrices are COO triplets (row, col, standard-form value) of the R1CS
 * (r1cs.nim:64-80), *without* the dummy public rows (they are added inside, fake_setup.nim:182-185).
 * ---------------------------------------------------------------------------------------------- */
typedef struct g16_r1cs_view {
  uint32_t nvars;          /* cfg.nWires */
  uint32_t npubs;          /* nPubIn + nPubOut */
  uint32_t neqs;           /* constraints.len */
  uint32_t flavour;
  uint64_t nnz[3];         /* A, B, C */
  const uint32_t* rows[3];
  const uint32_t* cols[3];
  const uint64_t* vals[3]; /* nnz x 4 limbs, standard form */
} g16_r1cs_view;

typedef struct g16_toxic {   /* ToxicWaste (fake_setup.nim:24-30), standard form */
  uint64_t alpha[4], beta[4], gamma[4], delta[4], tau[4];
} g16_toxic;

typedef struct g16_setup_out {       /* host buffers sized by the caller; any pointer may be NULL */
  uint64_t* points_a1;               /* nvars G1 */
  uint64_t* points_b1;               /* nvars G1 */
  uint64_t* points_b2;               /* nvars G2 */
  uint64_t* points_c1;               /* nvars - npubs - 1 G1 */
  uint64_t* points_h1;               /* domain G1 */
  uint64_t* points_ic;               /* npubs + 1 G1 */
  uint64_t* spec;                    /* alpha1, beta1, delta1 (G1) then beta2, gamma2, delta2 (G2): 3*8 + 3*16 limbs */
  uint64_t* dlog_a;                  /* nvars standard-form scalars a_j  (closed-form oracle) */
  uint64_t* dlog_b;                  /* nvars */
  uint64_t* dlog_k;                  /* nvars - npubs - 1 */
  uint64_t* dlog_h;                  /* domain */
  uint64_t* dlog_ic;                 /* npubs + 1 */
} g16_setup_out;

int g16_fake_setup(const g16_r1cs_view* r1cs, const g16_toxic* toxic, uint32_t* log_domain_out, g16_setup_out* out);

/* k_i * g1 / k_i * g2 for standard-form scalars (curves.nim:182-196 `**` on the generators) */
int g16_fixed_base_g1(const uint64_t* scalars_std, size_t n, uint64_t* points_out);
int g16_fixed_base_g2(const uint64_t* scalars_std, size_t n, uint64_t* points_out);

/* ------------------------------------------------------------------------------------------------
 * Diagnostics
 * ---------------------------------------------------------------------------------------------- */
/* device field/curve self-test against the library's portable host arithmetic; 0 = pass */
int g16_selftest(uint32_t seed, uint32_t cases);
/* integer-pipe microbenchmark: sustained 32-bit multiply-add rate of all SMs.
 * kind 0 = mad.lo.u32, 1 = mad.hi.u32, 2 = lo/hi carry pairs (IMAD.WIDE.X), 3 = full Montgomery multiplies.
 * Reports operations (MAC32, or modmuls for kind 3) per second.
 * FP64-pipe co-issue experiment: 4 = DFMA Montgomery multiplies (field_fp64.cuh) on the odd warps only,
 * 5 = IMAD Montgomery multiplies on the even warps only, 6 = both at once (modmuls/s of the even warps; the odd
 * warps do the same number of multiplies -- compare `ms` with kinds 4 and 5); 7 = DFMA multiplies on all warps;
 * 8 / 9 = 2 of 8 / 6 of 8 warps on DFMA, the rest on IMAD (total modmuls/s).
 * Instruction-mix models (multiplies/s): 10 = 128 IMAD.WIDE + 40 IADD3 (today's fmul), 11 = 112 + 152, 12 = 112 + 200.
 * 13 = Montgomery multiplies without the final conditional subtraction (operands and results in [0, 2p)). */
int g16_bench_int_pipe(int kind, double* ops_per_sec, float* ms);
/* GLV split of a scalar k < r (standard form): k = (+-k1) + (+-k2) * lambda (mod r), both magnitudes below 2^127,
 * lambda = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd the eigenvalue of (x, y) -> (beta x, y) on G1.  What the
 * prover does with the blinding scalars before s ** pi_A and r ** rho (prover.nim:298-299, csrc/glv.h); exported so that
 * the host arithmetic is testable without a device. */
int g16_glv_decompose(const uint64_t k_std[4], uint64_t k1_abs[2], uint64_t k2_abs[2], int* neg1, int* neg2);
/* number of kernels this library has launched in this process */
uint64_t g16_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* G16B200_H */
