// extern "C" boundary of libg16b200.so (declared in include/g16b200.h).
// Plain pointers and sizes only; every function returns a status code and never throws.
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <string.h>
#include "../../include/g16b200.h"
#include "abc.cuh"
#include "common.cuh"
#include "ec.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "prover.cuh"
#include "multi.cuh"
#include "glv.h"
#include <stdlib.h>

namespace g16 {

static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
const char* get_last_error() { return t_last_error.c_str(); }

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
uint64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

// per-device internal stream of the stream-ordered allocator (DevBuf); the pool keeps freed memory cached
static std::mutex g_pool_mu;
static cudaStream_t g_pool_streams[64] = {nullptr};
cudaStream_t pool_stream() {
  std::mutex& mu = g_pool_mu;
  cudaStream_t* streams = g_pool_streams;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!streams[dev]) {
    G16_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    cudaMemPool_t pool = nullptr;
    G16_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = ~0ull;
    if (const char* e = getenv("G16_POOL_KEEP_MB")) keep = (uint64_t)atoll(e) << 20;   // 0: give memory back on free
    G16_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    // experiment knob: L2 fetch granularity in bytes (32 / 64 / 128).  The bucket accumulation gathers 64-byte table
    // entries at random; ncu shows 2x the algorithmic DRAM bytes for it (profiles/ncu_traffic.json)
    if (const char* e = getenv("G16_L2_FETCH")) {
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)) != cudaSuccess) cudaGetLastError();
    }
  }
  return streams[dev];
}

void pool_trim() {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool = nullptr;
  cudaDeviceSynchronize();
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaGetLastError();
}

void fake_setup(const g16_r1cs_view& r, const g16_toxic& toxic, uint32_t* log_domain_out, g16_setup_out& out);
template <class F>
void fixed_base_mul(const Fr* scalars_dev, size_t n, Affine<F>* out_dev, cudaStream_t stream);
int selftest_run(uint32_t seed, uint32_t cases);
void bench_int_pipe(int kind, double* ops_per_sec, float* ms);

static void require_device() {
  static std::once_flag once;
  static int ok = 0;
  static std::string why;
  std::call_once(once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      why = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
      return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      why = "cudaGetDeviceProperties failed";
      return;
    }
    if (p.major != 10) {
      why = std::string("device '") + p.name + "' is not sm_100 (compute capability " + std::to_string(p.major) +
            "." + std::to_string(p.minor) + "); the library is built for sm_100a only";
      return;
    }
    ok = 1;
  });
  if (!ok) throw Error(G16_ERR_CUDA, why);
}

template <class Fn>
static int guard(Fn&& fn) {
  try {
    require_device();
    fn();
    return G16_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return G16_ERR_CUDA;
  }
}

// per-thread scratch of the fine-grained host-buffer calls (grow-only, reused between calls)
struct Scratch {
  DevBuf a, b, c, d, res;
  Msm<Fp> msm1;
  Msm<Fp2> msm2;
  cudaStream_t stream = nullptr;
  cudaStream_t s() {
    if (!stream) G16_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return stream;
  }
};
static Scratch& scratch() {
  static thread_local std::unique_ptr<Scratch> s;
  if (!s) s.reset(new Scratch());
  return *s;
}

template <class F>
static void msm_host(Msm<F>& eng, const uint64_t* scalars, int form, const uint64_t* points, size_t n, uint64_t* out) {
  G16_REQUIRE(form == G16_FORM_MONT || form == G16_FORM_STD, "unknown scalar form");
  G16_REQUIRE(out != nullptr, "output is null");
  G16_REQUIRE(n == 0 || (scalars != nullptr && points != nullptr), "incompatible sequence lengths");   // msm.nim:97
  Scratch& sc = scratch();
  cudaStream_t s = sc.s();
  sc.a.ensure(n * sizeof(Fr) + 16);
  sc.b.ensure(n * sizeof(Affine<F>) + 16);
  sc.res.ensure(sizeof(XYZZ<F>) + sizeof(Affine<F>));
  if (n) {
    G16_CUDA(cudaMemcpyAsync(sc.a.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    G16_CUDA(cudaMemcpyAsync(sc.b.p, points, n * sizeof(Affine<F>), cudaMemcpyHostToDevice, s));
  }
  XYZZ<F>* r = sc.res.as<XYZZ<F>>();
  Affine<F>* ra = reinterpret_cast<Affine<F>*>(r + 1);
  eng.run(sc.a.as<Fr>(), form == G16_FORM_MONT, sc.b.as<Affine<F>>(), n, r, s, 0);
  xyzz_sum_to_affine<F>(r, 1, ra, s);
  G16_CUDA(cudaMemcpyAsync(out, ra, sizeof(Affine<F>), cudaMemcpyDeviceToHost, s));
  G16_CUDA(cudaStreamSynchronize(s));
}

template <class F>
static void fixed_base_host(const uint64_t* scalars, size_t n, uint64_t* out) {
  G16_REQUIRE(n == 0 || (scalars && out), "null argument");
  if (!n) return;
  Scratch& sc = scratch();
  cudaStream_t s = sc.s();
  sc.a.ensure(n * sizeof(Fr));
  sc.b.ensure(n * sizeof(Affine<F>));
  G16_CUDA(cudaMemcpyAsync(sc.a.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
  fixed_base_mul<F>(sc.a.as<Fr>(), n, sc.b.as<Affine<F>>(), s);
  G16_CUDA(cudaMemcpyAsync(out, sc.b.p, n * sizeof(Affine<F>), cudaMemcpyDeviceToHost, s));
  G16_CUDA(cudaStreamSynchronize(s));
}
}  // namespace g16

using namespace g16;

struct g16_ctx {
  std::unique_ptr<Prover> prover;          // one shard (or the whole key) on one device
  std::unique_ptr<MultiProver> multi;      // the whole key over several devices of this process
  uint64_t launches0 = 0, launches1 = 0;
};

// devices of an in-library multi-GPU context: G16_DEVICES="0,1,.." if set, else first .. first+n-1
static std::vector<int> device_list(int first, int n) {
  std::vector<int> d;
  if (const char* e = getenv("G16_DEVICES")) {
    for (const char* p = e; *p;) {
      char* end = nullptr;
      long v = strtol(p, &end, 10);
      if (end == p) break;
      d.push_back((int)v);
      p = (*end == ',') ? end + 1 : end;
    }
    if ((int)d.size() >= n) {
      d.resize((size_t)n);
      return d;
    }
    d.clear();
  }
  for (int i = 0; i < n; i++) d.push_back(first + i);
  return d;
}

struct g16_msm_plan {
  int g2 = 0;
  int window = 0;
  size_t max_n = 0;
  Msm<Fp> m1;
  Msm<Fp2> m2;
  DevBuf table;          // resident window table (g16_msm_plan_build_table)
  size_t table_n = 0;
  int table_c = 0;
};

extern "C" {

const char* g16_last_error(void) { return get_last_error(); }
int g16_version(void) { return 2; }
uint64_t g16_kernel_launch_count(void) { return g_launches.load(); }

int g16_set_device(int device) {
  return guard([&] { G16_CUDA(cudaSetDevice(device)); });
}
int g16_device_count(int* count) {
  return guard([&] {
    G16_REQUIRE(count != nullptr, "count is null");
    G16_CUDA(cudaGetDeviceCount(count));
  });
}

int g16_release_cached_memory(void) {
  return guard([&] {
    int n = 0, cur = 0;
    G16_CUDA(cudaGetDeviceCount(&n));
    cudaGetDevice(&cur);
    ntt_release_tables();
    for (int d = 0; d < n && d < 64; d++) {
      if (!g_pool_streams[d]) continue;              // only devices this library has allocated on
      cudaSetDevice(d);
      pool_trim();
    }
    cudaSetDevice(cur);
  });
}
int g16_glv_decompose(const uint64_t k_std[4], uint64_t k1_abs[2], uint64_t k2_abs[2], int* neg1, int* neg2) {
  // pure host arithmetic: no device needed, so no guard()
  if (!k_std || !k1_abs || !k2_abs || !neg1 || !neg2) {
    set_last_error("g16_glv_decompose: null argument");
    return G16_ERR_ARG;
  }
  bool ok = false;
  GlvSplit sp = glv_decompose(k_std, &ok);
  k1_abs[0] = sp.k1[0];
  k1_abs[1] = sp.k1[1];
  k2_abs[0] = sp.k2[0];
  k2_abs[1] = sp.k2[1];
  *neg1 = (int)sp.neg1;
  *neg2 = (int)sp.neg2;
  if (!ok) {
    set_last_error("g16_glv_decompose: scalar out of range");
    return G16_ERR_ARG;
  }
  return G16_OK;
}
int g16_host_register(const void* ptr, size_t bytes) {
  return guard([&] {
    G16_REQUIRE(ptr != nullptr && bytes > 0, "null range");
    // read-only mappings need the read-only flag; writable ranges take the default
    cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterReadOnly | cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
    }
    if (e != cudaSuccess) cudaGetLastError();
    G16_CUDA(e);
  });
}
int g16_host_unregister(const void* ptr) {
  return guard([&] {
    G16_REQUIRE(ptr != nullptr, "null range");
    G16_CUDA(cudaHostUnregister(const_cast<void*>(ptr)));
  });
}

int g16_msm_g1(const uint64_t* scalars, int scalar_form, const uint64_t* points, size_t n, uint64_t out[8]) {
  return guard([&] { msm_host<Fp>(scratch().msm1, scalars, scalar_form, points, n, out); });
}
int g16_msm_g2(const uint64_t* scalars, int scalar_form, const uint64_t* points, size_t n, uint64_t out[16]) {
  return guard([&] { msm_host<Fp2>(scratch().msm2, scalars, scalar_form, points, n, out); });
}

int g16_ntt_fr(const uint64_t* in, uint64_t* out, int log_n, int inverse) {
  return guard([&] {
    G16_REQUIRE(in != nullptr && out != nullptr, "input must have the same size as the domain");   // ntt.nim:57
    G16_REQUIRE(log_n >= 0 && log_n <= 26, "domain must have a power-of-two size");                  // ntt.nim:56
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) {                      // ntt.nim:24-26: the size-1 transform is the identity
      memcpy(out, in, sizeof(Fr));
      return;
    }
    Scratch& sc = scratch();
    cudaStream_t s = sc.s();
    sc.a.ensure(n * sizeof(Fr));
    sc.b.ensure(n * sizeof(Fr));
    G16_CUDA(cudaMemcpyAsync(sc.a.p, in, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ntt_natural(sc.a.as<Fr>(), sc.b.as<Fr>(), sc.a.as<Fr>(), log_n, inverse != 0, s);
    G16_CUDA(cudaMemcpyAsync(out, sc.b.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    G16_CUDA(cudaStreamSynchronize(s));
  });
}

int g16_quotient(const uint64_t* az, const uint64_t* bz, int log_n, int flavour, uint64_t* qs_out) {
  return guard([&] {
    G16_REQUIRE(az && bz && qs_out, "null vector");
    G16_REQUIRE(log_n >= 1 && log_n <= 26, "domain must be 2^1 .. 2^26 (prover.nim:101 needs n >= 2)");
    G16_REQUIRE(flavour == G16_FLAVOUR_JENSGROTH || flavour == G16_FLAVOUR_SNARKJS, "unknown flavour");
    size_t n = (size_t)1 << log_n;
    Scratch& sc = scratch();
    cudaStream_t s = sc.s();
    sc.a.ensure(3 * n * sizeof(Fr));
    sc.b.ensure(n * sizeof(Fr));
    G16_CUDA(cudaMemcpyAsync(sc.a.p, az, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    G16_CUDA(cudaMemcpyAsync(sc.a.as<Fr>() + n, bz, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    quotient(sc.a.as<Fr>(), sc.b.as<Fr>(), log_n, flavour, s);
    G16_CUDA(cudaMemcpyAsync(qs_out, sc.b.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    G16_CUDA(cudaStreamSynchronize(s));
  });
}

int g16_build_abc(const void* coeffs, size_t nnz, int coeff_format, const uint64_t* witness, int witness_form,
                  size_t m, int log_n, uint64_t* az, uint64_t* bz, uint64_t* cz) {
  return guard([&] {
    G16_REQUIRE(log_n >= 0 && log_n <= 26, "domain must have a power-of-two size");
    G16_REQUIRE(nnz == 0 || coeffs != nullptr, "coefficient list is null");
    G16_REQUIRE(m == 0 || witness != nullptr, "witness is null");
    G16_REQUIRE(witness_form == G16_FORM_MONT || witness_form == G16_FORM_STD, "unknown witness form");
    G16_REQUIRE(coeff_format == G16_COEFF_PACKED44_R2 || coeff_format == G16_COEFF_STRUCT48_MONT,
                "unknown coefficient record format");
    size_t n = (size_t)1 << log_n;
    size_t rec = coeff_format == G16_COEFF_PACKED44_R2 ? 44 : 48;
    Scratch& sc = scratch();
    cudaStream_t s = sc.s();
    sc.a.ensure(nnz * rec + 16);
    sc.b.ensure(m * sizeof(Fr) + 32);
    sc.c.ensure(3 * n * sizeof(Fr));
    if (nnz) G16_CUDA(cudaMemcpyAsync(sc.a.p, coeffs, nnz * rec, cudaMemcpyHostToDevice, s));
    if (m) G16_CUDA(cudaMemcpyAsync(sc.b.p, witness, m * sizeof(Fr), cudaMemcpyHostToDevice, s));
    if (witness_form == G16_FORM_MONT) fr_from_mont(sc.b.as<Fr>(), sc.b.as<Fr>(), m, s);
    SparseCsr csr;
    coeffs_to_csr(csr, sc.a.p, nnz, coeff_format, log_n, m, s);
    build_abc(csr, sc.b.as<Fr>(), sc.c.as<Fr>(), log_n, s);
    if (az) G16_CUDA(cudaMemcpyAsync(az, sc.c.as<Fr>(), n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    if (bz) G16_CUDA(cudaMemcpyAsync(bz, sc.c.as<Fr>() + n, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    if (cz) G16_CUDA(cudaMemcpyAsync(cz, sc.c.as<Fr>() + 2 * n, n * sizeof(Fr), cudaMemcpyDeviceToHost, s));
    G16_CUDA(cudaStreamSynchronize(s));
  });
}

// ---------------------------------------------------------------------------------------------
int g16_ctx_create(const g16_zkey_view* zkey, int shard_index, int shard_count, g16_ctx** out) {
  return guard([&] {
    G16_REQUIRE(zkey != nullptr && out != nullptr, "null argument");
    std::unique_ptr<g16_ctx> c(new g16_ctx());
    int ndev = shard_count < 0 ? -shard_count : 0;
    if (shard_count == 1 && shard_index == 0)
      if (const char* e = getenv("G16_NGPUS")) {
        int v = atoi(e);
        if (v > 1) ndev = v;
      }
    if (ndev >= 1) {
      G16_REQUIRE(shard_index >= 0, "bad first device");
      c->multi.reset(new MultiProver(*zkey, device_list(shard_index, ndev)));
    } else {
      c->prover.reset(new Prover(*zkey, shard_index, shard_count));
    }
    *out = c.release();
  });
}
int g16_ctx_clone(g16_ctx* ctx, g16_ctx** out) {
  return guard([&] {
    G16_REQUIRE(ctx && (ctx->prover || ctx->multi) && out, "null argument");
    std::unique_ptr<g16_ctx> c(new g16_ctx());
    if (ctx->multi) c->multi.reset(new MultiProver(*ctx->multi));
    else c->prover.reset(new Prover(ctx->prover->resident()));
    *out = c.release();
  });
}
void g16_ctx_destroy(g16_ctx* ctx) { delete ctx; }

static void prove_submit(g16_ctx* ctx, const void* witness, int form, int mem_kind, const uint64_t r[4],
                         const uint64_t s[4]) {
  G16_REQUIRE(ctx && (ctx->prover || ctx->multi), "context is null");
  ctx->launches0 = g_launches.load();
  if (ctx->multi) {
    ctx->multi->submit(witness, form, mem_kind, r, s);
    ctx->launches1 = g_launches.load();
    return;
  }
  G16_REQUIRE(ctx->prover->shard_count() == 1, "g16_prove needs an unsharded context; use g16_prove_partials");
  G16_REQUIRE(!ctx->prover->in_flight(), "a proof is already in flight on this context");
  Prover& p = *ctx->prover;
  p.start_mask(r, s);
  p.load_witness(witness, form, mem_kind);
  p.run_msms(nullptr);
  p.finish_async();
  ctx->launches1 = g_launches.load();
}
static void prove_wait(g16_ctx* ctx, g16_proof* proof, g16_stats* stats) {
  G16_REQUIRE(ctx && (ctx->prover || ctx->multi), "context is null");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (ctx->multi) ctx->multi->wait(proof, stats);
  else ctx->prover->wait(proof, stats);
  if (stats) stats->kernel_launches = (uint32_t)(ctx->launches1 - ctx->launches0);
}

int g16_prove(g16_ctx* ctx, const uint64_t* witness, int witness_form, const uint64_t r_std[4],
              const uint64_t s_std[4], g16_proof* proof, g16_stats* stats) {
  return guard([&] {
    prove_submit(ctx, witness, witness_form, G16_MEM_HOST, r_std, s_std);
    prove_wait(ctx, proof, stats);
  });
}
int g16_prove_dev(g16_ctx* ctx, const void* witness_std_dev, const uint64_t r_std[4], const uint64_t s_std[4],
                  g16_proof* proof, g16_stats* stats) {
  return guard([&] {
    prove_submit(ctx, witness_std_dev, G16_FORM_STD, G16_MEM_DEVICE, r_std, s_std);
    prove_wait(ctx, proof, stats);
  });
}
int g16_prove_submit(g16_ctx* ctx, const void* witness, int witness_form, int witness_mem_kind,
                     const uint64_t r_std[4], const uint64_t s_std[4]) {
  return guard([&] { prove_submit(ctx, witness, witness_form, witness_mem_kind, r_std, s_std); });
}
int g16_prove_wait(g16_ctx* ctx, g16_proof* proof, g16_stats* stats) {
  return guard([&] { prove_wait(ctx, proof, stats); });
}

int g16_shard_plan(uint64_t nvars, uint64_t npubs, uint64_t domain_size, int shard_index, int shard_count,
                   uint64_t out[10]) {
  // pure host arithmetic: no device needed, so no guard()
  if (out == nullptr || shard_count < 1 || shard_index < 0 || shard_index >= shard_count) {
    set_last_error("g16_shard_plan: bad argument");
    return G16_ERR_ARG;
  }
  ShardPlan p;
  shard_plan((size_t)nvars, (size_t)npubs, (size_t)domain_size, shard_index, shard_count, p);
  const size_t v[10] = {p.a1_lo, p.a1_hi, p.b1_lo, p.b1_hi, p.c1_lo, p.c1_hi, p.b2_lo, p.b2_hi, p.h_lo, p.h_hi};
  for (int i = 0; i < 10; i++) out[i] = v[i];
  return G16_OK;
}
int g16_ctx_order_stream(g16_ctx* ctx, void* stream, int direction) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "g16_ctx_order_stream needs a single-shard context");
    G16_REQUIRE(direction == 0 || direction == 1, "direction must be 0 or 1");
    ctx->prover->order_stream(reinterpret_cast<cudaStream_t>(stream), direction);
  });
}
int g16_ctx_layout(g16_ctx* ctx, int* window_tables, uint64_t* device_bytes) {
  return guard([&] {
    G16_REQUIRE(ctx && (ctx->prover || ctx->multi), "context is null");
    Prover& p = ctx->multi ? ctx->multi->shard(0) : *ctx->prover;
    if (window_tables) *window_tables = p.table_layout() ? 1 : 0;
    if (device_bytes) *device_bytes = ctx->multi ? ctx->multi->resident_bytes() : p.resident_bytes();
  });
}
int g16_ctx_last_witness_bytes(g16_ctx* ctx, uint64_t* bytes) {
  return guard([&] {
    G16_REQUIRE(ctx && (ctx->prover || ctx->multi) && bytes, "null argument");
    *bytes = ctx->multi ? ctx->multi->last_witness_bytes() : ctx->prover->last_witness_bytes();
  });
}
int g16_ctx_set_mask(g16_ctx* ctx, const uint64_t r_std[4], const uint64_t s_std[4]) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null (or a multi-device context, which announces the masks itself)");
    G16_REQUIRE(!ctx->prover->in_flight(), "a proof is already in flight on this context");
    ctx->prover->set_mask(r_std, s_std);
  });
}
int g16_prove_partials(g16_ctx* ctx, const uint64_t* witness, int witness_form, int witness_mem_kind,
                       void* partials_dev, g16_stats* stats) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null");
    G16_REQUIRE(partials_dev != nullptr, "partials buffer is null");
    G16_REQUIRE(!ctx->prover->in_flight(), "a proof is already in flight on this context");
    uint64_t l0 = g_launches.load();
    if (stats) memset(stats, 0, sizeof(*stats));
    Prover& p = *ctx->prover;
    p.load_witness(witness, witness_form, witness_mem_kind);
    p.run_msms(nullptr);
    p.partials_to_affine_async(partials_dev);
    p.partials_wait(stats);
    if (stats) stats->kernel_launches = (uint32_t)(g_launches.load() - l0);
  });
}
int g16_prove_partials_submit(g16_ctx* ctx, const void* witness, int witness_form, int witness_mem_kind,
                              void* partials_dev) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null");
    G16_REQUIRE(partials_dev != nullptr, "partials buffer is null");
    G16_REQUIRE(!ctx->prover->in_flight(), "a proof is already in flight on this context");
    Prover& p = *ctx->prover;
    p.load_witness(witness, witness_form, witness_mem_kind);
    p.run_msms(nullptr);
    p.partials_to_affine_async(partials_dev);
  });
}
int g16_prove_partials_wait(g16_ctx* ctx, g16_stats* stats) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null");
    if (stats) memset(stats, 0, sizeof(*stats));
    ctx->prover->partials_wait(stats);
  });
}

int g16_ctx_last_partials(g16_ctx* ctx, void* partials_dev) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover && partials_dev, "null argument");
    ctx->prover->partials_to_affine(partials_dev);
  });
}

int g16_prove_finish(g16_ctx* ctx, const void* gathered_partials_dev, int count, const uint64_t r_std[4],
                     const uint64_t s_std[4], g16_proof* proof) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null");
    G16_REQUIRE(gathered_partials_dev != nullptr, "partials buffer is null");
    Prover& p = *ctx->prover;
    if (p.masked_partials()) G16_REQUIRE(p.same_mask(r_std, s_std), "masks differ from g16_ctx_set_mask");
    else p.start_mask(r_std, s_std);
    p.sum_partials(gathered_partials_dev, count);
    p.finish(proof, nullptr);
  });
}
int g16_prove_finish_submit(g16_ctx* ctx, const void* gathered_partials_dev, int count, const uint64_t r_std[4],
                            const uint64_t s_std[4]) {
  return guard([&] {
    G16_REQUIRE(ctx && ctx->prover, "context is null");
    G16_REQUIRE(gathered_partials_dev != nullptr, "partials buffer is null");
    Prover& p = *ctx->prover;
    if (p.masked_partials()) G16_REQUIRE(p.same_mask(r_std, s_std), "masks differ from g16_ctx_set_mask");
    else p.start_mask(r_std, s_std);
    p.sum_partials(gathered_partials_dev, count);
    p.finish_async();
  });
}

// ---------------------------------------------------------------------------------------------
int g16_msm_plan_create(int g2, size_t max_n, int window_bits, g16_msm_plan** out) {
  return guard([&] {
    G16_REQUIRE(out != nullptr, "null argument");
    G16_REQUIRE(window_bits == 0 || (window_bits >= 2 && window_bits <= 22), "window must be 0 (auto) or 2..22");
    std::unique_ptr<g16_msm_plan> p(new g16_msm_plan());
    p->g2 = g2 ? 1 : 0;
    p->window = window_bits;
    p->max_n = max_n;
    *out = p.release();
  });
}
void g16_msm_plan_destroy(g16_msm_plan* plan) { delete plan; }

int g16_msm_dev(g16_msm_plan* plan, const void* scalars_dev, int scalar_form, const void* points_dev, size_t n,
                void* result_xyzz_dev, void* stream) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr && result_xyzz_dev != nullptr, "null argument");
    G16_REQUIRE(scalar_form == G16_FORM_MONT || scalar_form == G16_FORM_STD, "unknown scalar form");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (plan->g2)
      plan->m2.run(reinterpret_cast<const Fr*>(scalars_dev), scalar_form == G16_FORM_MONT,
                   reinterpret_cast<const G2Affine*>(points_dev), n, reinterpret_cast<G2XYZZ*>(result_xyzz_dev), s,
                   plan->window);
    else
      plan->m1.run(reinterpret_cast<const Fr*>(scalars_dev), scalar_form == G16_FORM_MONT,
                   reinterpret_cast<const G1Affine*>(points_dev), n, reinterpret_cast<G1XYZZ*>(result_xyzz_dev), s,
                   plan->window);
  });
}

int g16_msm_plan_build_table(g16_msm_plan* plan, const void* points_dev, size_t n, void** table_dev_out) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr && points_dev != nullptr && n > 0, "bad argument");
    int c = plan->window ? plan->window : msm_pick_window(n, true);
    size_t elem = plan->g2 ? sizeof(G2Affine) : sizeof(G1Affine);
    plan->table.ensure((size_t)msm_num_windows(c) * n * elem);
    cudaStream_t s = scratch().s();
    if (plan->g2)
      msm_build_table<Fp2>(reinterpret_cast<const G2Affine*>(points_dev), n, 0, c, plan->table.as<G2Affine>(), s);
    else
      msm_build_table<Fp>(reinterpret_cast<const G1Affine*>(points_dev), n, 0, c, plan->table.as<G1Affine>(), s);
    G16_CUDA(cudaStreamSynchronize(s));
    plan->table_n = n;
    plan->table_c = c;
    if (table_dev_out) *table_dev_out = plan->table.p;
  });
}

int g16_msm_dev_table(g16_msm_plan* plan, const void* scalars_dev, int scalar_form, size_t n, void* result_xyzz_dev,
                      void* stream) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr && result_xyzz_dev != nullptr, "null argument");
    G16_REQUIRE(plan->table_n == n && n > 0, "no table of this size: call g16_msm_plan_build_table first");
    G16_REQUIRE(scalar_form == G16_FORM_MONT || scalar_form == G16_FORM_STD, "unknown scalar form");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (plan->g2)
      plan->m2.run_precomp(reinterpret_cast<const Fr*>(scalars_dev), scalar_form == G16_FORM_MONT,
                           plan->table.as<G2Affine>(), n, plan->table_c, reinterpret_cast<G2XYZZ*>(result_xyzz_dev), s);
    else
      plan->m1.run_precomp(reinterpret_cast<const Fr*>(scalars_dev), scalar_form == G16_FORM_MONT,
                           plan->table.as<G1Affine>(), n, plan->table_c, reinterpret_cast<G1XYZZ*>(result_xyzz_dev), s);
  });
}

int g16_msm_result_to_affine(int g2, const void* result_xyzz_dev, int count, uint64_t* out_host) {
  return guard([&] {
    G16_REQUIRE(result_xyzz_dev && out_host && count >= 1, "bad argument");
    Scratch& sc = scratch();
    cudaStream_t s = sc.s();
    sc.res.ensure(sizeof(G2XYZZ) + sizeof(G2Affine));
    G16_CUDA(cudaDeviceSynchronize());
    if (g2) {
      G2Affine* ra = sc.res.as<G2Affine>();
      xyzz_sum_to_affine<Fp2>(reinterpret_cast<const G2XYZZ*>(result_xyzz_dev), count, ra, s);
      G16_CUDA(cudaMemcpyAsync(out_host, ra, sizeof(G2Affine), cudaMemcpyDeviceToHost, s));
    } else {
      G1Affine* ra = sc.res.as<G1Affine>();
      xyzz_sum_to_affine<Fp>(reinterpret_cast<const G1XYZZ*>(result_xyzz_dev), count, ra, s);
      G16_CUDA(cudaMemcpyAsync(out_host, ra, sizeof(G1Affine), cudaMemcpyDeviceToHost, s));
    }
    G16_CUDA(cudaStreamSynchronize(s));
  });
}

int g16_msm_plan_info(const g16_msm_plan* plan, int* window_bits, int* num_windows, size_t* workspace_bytes) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr, "null argument");
    if (window_bits) *window_bits = plan->g2 ? plan->m2.last_c : plan->m1.last_c;
    if (num_windows) *num_windows = plan->g2 ? plan->m2.last_nwin : plan->m1.last_nwin;
    if (workspace_bytes) *workspace_bytes = plan->g2 ? plan->m2.workspace_bytes() : plan->m1.workspace_bytes();
  });
}

int g16_msm_plan_profile(g16_msm_plan* plan, int enable) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr, "null argument");
    plan->m1.set_profile(enable != 0);
    plan->m2.set_profile(enable != 0);
  });
}
int g16_msm_plan_last_profile(const g16_msm_plan* plan, float* accumulate_ms, float* total_ms, uint64_t* pairs) {
  return guard([&] {
    G16_REQUIRE(plan != nullptr, "null argument");
    if (accumulate_ms) *accumulate_ms = plan->g2 ? plan->m2.last_accum_ms() : plan->m1.last_accum_ms();
    if (total_ms) *total_ms = plan->g2 ? plan->m2.last_total_ms() : plan->m1.last_total_ms();
    if (pairs) *pairs = plan->g2 ? plan->m2.last_pairs : plan->m1.last_pairs;
  });
}
int g16_ctx_timer_start(g16_ctx* ctx) {
  return guard([&] {
    G16_REQUIRE(ctx && (ctx->prover || ctx->multi), "context is null");
    if (ctx->multi) ctx->multi->timer_start();
    else ctx->prover->timer_start();
  });
}
int g16_ctx_timer_stop(g16_ctx* ctx, float* elapsed_ms) {
  return guard([&] {
    G16_REQUIRE(ctx && (ctx->prover || ctx->multi) && elapsed_ms, "null argument");
    *elapsed_ms = ctx->multi ? ctx->multi->timer_stop() : ctx->prover->timer_stop();
  });
}

int g16_ntt_prepare(int log_n) {
  return guard([&] {
    Scratch& sc = scratch();
    ntt_prepare(log_n, sc.s());
  });
}
int g16_ntt_fr_dev(const void* in_dev, void* out_dev, void* work_dev, int log_n, int inverse, void* stream) {
  return guard([&] {
    G16_REQUIRE(in_dev && out_dev && work_dev, "null argument");
    G16_REQUIRE(log_n >= 1 && log_n <= 26, "domain must be 2^1 .. 2^26");
    ntt_natural(reinterpret_cast<const Fr*>(in_dev), reinterpret_cast<Fr*>(out_dev), reinterpret_cast<Fr*>(work_dev),
                log_n, inverse != 0, reinterpret_cast<cudaStream_t>(stream));
  });
}
int g16_quotient_dev(void* abc_dev, void* qs_dev, int log_n, int flavour, void* stream) {
  return guard([&] {
    G16_REQUIRE(abc_dev && qs_dev, "null argument");
    G16_REQUIRE(log_n >= 1 && log_n <= 26, "domain must be 2^1 .. 2^26");
    G16_REQUIRE(flavour == G16_FLAVOUR_JENSGROTH || flavour == G16_FLAVOUR_SNARKJS, "unknown flavour");
    quotient(reinterpret_cast<Fr*>(abc_dev), reinterpret_cast<Fr*>(qs_dev), log_n, flavour,
             reinterpret_cast<cudaStream_t>(stream));
  });
}

// ---------------------------------------------------------------------------------------------
int g16_fake_setup(const g16_r1cs_view* r1cs, const g16_toxic* toxic, uint32_t* log_domain_out, g16_setup_out* out) {
  return guard([&] {
    G16_REQUIRE(r1cs && toxic && out, "null argument");
    fake_setup(*r1cs, *toxic, log_domain_out, *out);
  });
}

int g16_fixed_base_g1(const uint64_t* scalars_std, size_t n, uint64_t* points_out) {
  return guard([&] { fixed_base_host<Fp>(scalars_std, n, points_out); });
}
int g16_fixed_base_g2(const uint64_t* scalars_std, size_t n, uint64_t* points_out) {
  return guard([&] { fixed_base_host<Fp2>(scalars_std, n, points_out); });
}

int g16_selftest(uint32_t seed, uint32_t cases) {
  int rc = -1;
  int st = guard([&] { rc = selftest_run(seed, cases); });
  if (st != G16_OK) return st;
  return rc;
}
int g16_bench_int_pipe(int kind, double* ops_per_sec, float* ms) {
  return guard([&] {
    G16_REQUIRE(kind >= 0 && kind <= 13 && ops_per_sec && ms, "bad argument");
    bench_int_pipe(kind, ops_per_sec, ms);
  });
}

}  // extern "C"
