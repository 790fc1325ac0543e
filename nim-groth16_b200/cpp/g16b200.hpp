// g16b200.hpp -- C++ host side over the C ABI (include/g16b200.h), mirroring nim-groth16's own interface for the
// hot path: same proc names, argument meaning and error behaviour, so that code written against the reference
// reads the same.  The reference's host language is Nim (shim: nim-groth16_b200/nim/g16b200.nim); there is no Nim
// toolchain in this build environment, so this header is the compiled-language host mirror that is actually
// built and exercised (cpp/g16prove.cpp, tests/test_cpp_host.py).
//
//   reference (file:line)                                   here
//   groth16/files/zkey.nim:241-246     parseZKey            groth16::parseZKey      (zero-copy: the file is mmap'ed, sections are views)
//   groth16/files/witness.nim:71-76    parseWitness         groth16::parseWitness
//   groth16/prover.nim:215-304         generateProofWithMask        groth16::generateProofWithMask
//   groth16/prover.nim:308             generateProofWithTrivialMask groth16::generateProofWithTrivialMask
//   groth16/prover.nim:312-319         generateProof                groth16::generateProof
//   groth16/bn128/msm.nim:89,128       msmMultiThreadedG1/G2        groth16::msmMultiThreadedG1/G2
//   groth16/math/ntt.nim:55,139        forwardNTT / inverseNTT      groth16::forwardNTT / inverseNTT
//   groth16/prover.nim:56,118,158      buildABC, computeQuotientPointwise, computeSnarkjsScalarCoeffs
//   groth16/files/export_json.nim:25-80 exportPublicIO / exportProof groth16::exportPublicIO / exportProof
//
// Failures raise groth16::AssertionDefect with the reference's message (the reference asserts).
// Layout: Fr / Fp = 4 x u64 little-endian limbs; curve points and Az/Bz/Cz/qs are Montgomery
// residues (R = 2^256) exactly as in a .zkey; witness values are the standard-form integers of the .wtns file.
#pragma once
#include <stdint.h>
#include <string.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/random.h>
#include <sys/stat.h>
#include <unistd.h>

#include <array>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/g16b200.h"

namespace groth16 {

struct AssertionDefect : std::runtime_error {
  explicit AssertionDefect(const std::string& m) : std::runtime_error(m) {}
};
inline void check(int status) {
  if (status != 0) throw AssertionDefect(std::string("g16b200: ") + g16_last_error());
}
inline void doAssert(bool cond, const char* msg) {
  if (!cond) throw AssertionDefect(msg);
}

struct Fr { uint64_t limb[4]; };                 // scalar field element
struct Fp { uint64_t limb[4]; };                 // base field element (Montgomery in points)
struct G1 { Fp x, y; };                          // affine, infinity = (0,0)   curves.nim:33,49
struct Fp2 { Fp c0, c1; };
struct G2 { Fp2 x, y; };                         // curves.nim:34,50
static_assert(sizeof(G1) == 64 && sizeof(G2) == 128 && sizeof(Fr) == 32, "boundary layout");

enum Flavour { JensGroth = 0, Snarkjs = 1 };     // zkey_types.nim:10-13

namespace detail {
// BN254 moduli, little-endian limbs, and -m^-1 mod 2^64
static const uint64_t P_MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t R_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t P_INV = 0x87d20782e4866389ull, R_INV = 0xc2e1f593efffffffull;

// x * 2^-256 mod m (Montgomery reduction of a 4-limb value): Montgomery -> standard form
inline void from_mont(const uint64_t x[4], const uint64_t m[4], uint64_t inv, uint64_t out[4]) {
  uint64_t t[5] = {x[0], x[1], x[2], x[3], 0};
  for (int i = 0; i < 4; i++) {
    uint64_t q = t[0] * inv;
    unsigned __int128 c = (unsigned __int128)q * m[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (unsigned __int128)q * m[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = (uint64_t)(c >> 64);
  }
  // conditional subtraction
  uint64_t u[4];
  unsigned __int128 b = 0;
  for (int j = 0; j < 4; j++) {
    unsigned __int128 d = (unsigned __int128)t[j] - m[j] - (uint64_t)b;
    u[j] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
  bool ge = t[4] || !b;
  for (int j = 0; j < 4; j++) out[j] = ge ? u[j] : t[j];
}
inline bool less_than(const uint64_t a[4], const uint64_t m[4]) {
  for (int j = 3; j >= 0; j--)
    if (a[j] != m[j]) return a[j] < m[j];
  return false;
}
inline std::string to_decimal(const uint64_t x[4]) {
  uint64_t t[4] = {x[0], x[1], x[2], x[3]};
  std::string s;
  while (t[0] | t[1] | t[2] | t[3]) {
    unsigned __int128 rem = 0;
    for (int j = 3; j >= 0; j--) {
      unsigned __int128 cur = (rem << 64) | t[j];
      t[j] = (uint64_t)(cur / 10);
      rem = cur % 10;
    }
    s.push_back((char)('0' + (int)rem));
  }
  if (s.empty()) s = "0";
  return std::string(s.rbegin(), s.rend());
}
inline std::string fp_decimal(const Fp& a) {       // Montgomery point coordinate -> decimal string
  uint64_t v[4];
  from_mont(a.limb, P_MOD, P_INV, v);
  return to_decimal(v);
}
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Section { const uint8_t* p = nullptr; size_t len = 0; };
// container.nim:75-93: magic, version, nsections, then (id u32, len u64, payload)*
// A read-only memory mapping of a whole file: the .zkey point sections and the .wtns payload are consumed where
// the page cache holds them -- no read() into a heap buffer, no per-element parsing.  With
// pin() the pages are registered with the CUDA driver (g16_host_register), so that the uploads of g16_ctx_create
// are DMA transfers at PCIe speed instead of staged pageable copies; worth it when the mapping feeds several
// contexts (one per GPU) or several proofs.
class FileMap {
 public:
  explicit FileMap(const std::string& fname) {
    fd_ = ::open(fname.c_str(), O_RDONLY);
    if (fd_ < 0) throw AssertionDefect("cannot open file `" + fname + "`");
    struct stat st;
    if (fstat(fd_, &st) != 0) {
      ::close(fd_);
      throw AssertionDefect("cannot read file `" + fname + "`");
    }
    len_ = (size_t)st.st_size;
    if (len_) {
      void* m = mmap(nullptr, len_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (m == MAP_FAILED) {
        ::close(fd_);
        throw AssertionDefect("cannot map file `" + fname + "`");
      }
      p_ = static_cast<const uint8_t*>(m);
      madvise(const_cast<uint8_t*>(p_), len_, MADV_SEQUENTIAL | MADV_WILLNEED);
    }
  }
  ~FileMap() {
    if (pinned_) g16_host_unregister(p_);
    if (p_) munmap(const_cast<uint8_t*>(p_), len_);
    if (fd_ >= 0) ::close(fd_);
  }
  FileMap(const FileMap&) = delete;
  FileMap& operator=(const FileMap&) = delete;
  const uint8_t* data() const { return p_; }
  size_t size() const { return len_; }
  bool pin() {                         // false when the platform refuses read-only registration: the uploads still work
    if (!pinned_ && p_) pinned_ = g16_host_register(p_, len_) == 0;
    return pinned_;
  }

 private:
  int fd_ = -1;
  const uint8_t* p_ = nullptr;
  size_t len_ = 0;
  bool pinned_ = false;
};

inline std::vector<Section> parse_container(const uint8_t* buf, size_t size, const char magic[4], uint32_t version,
                                            int max_id) {
  doAssert(size >= 12 && memcmp(buf, magic, 4) == 0, "not a file of the expected kind (bad magic)");
  doAssert(rd32(buf + 4) == version, "unexpected container version");
  uint32_t nsec = rd32(buf + 8);
  std::vector<Section> out(max_id + 1);
  size_t pos = 12;
  for (uint32_t i = 0; i < nsec; i++) {
    doAssert(pos + 12 <= size, "truncated file");
    uint32_t id = rd32(buf + pos);
    uint64_t len = rd64(buf + pos + 4);
    pos += 12;
    doAssert(len <= size && pos + len <= size, "truncated file");
    if (id <= (uint32_t)max_id) out[id] = Section{buf + pos, (size_t)len};
    pos += len;
  }
  return out;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------ types
struct GrothHeader {                             // zkey_types.nim:15-22
  std::string curve = "bn128";
  Flavour flavour = Snarkjs;
  int nvars = 0, npubs = 0, domainSize = 0, logDomainSize = 0;
};
struct SpecPoints { G1 alpha1, beta1, delta1; G2 beta2, gamma2, delta2; };   // zkey_types.nim:24-31
// The point arrays are views into the file buffer the ZKey owns (no per-element parsing: the sections are already
// little-endian Montgomery records).
struct ZKey {                                    // zkey_types.nim:54-60
  GrothHeader header;
  SpecPoints specPoints;
  const G1* pointsIC = nullptr;                  // npubs + 1            (VerifierPoints)
  const G1* pointsA1 = nullptr;                  // nvars                (ProverPoints, zkey_types.nim:34-40)
  const G1* pointsB1 = nullptr;
  const G2* pointsB2 = nullptr;
  const G1* pointsC1 = nullptr;                  // nvars - npubs - 1
  const G1* pointsH1 = nullptr;                  // domainSize
  const uint8_t* coeffs = nullptr;               // ncoeffs packed 44-byte records, value * R^2 (zkey.nim:169-188)
  size_t ncoeffs = 0;
  std::shared_ptr<detail::FileMap> file;         // owns the mapping the views point into
};
struct Witness {                                 // witness.nim:28-32
  std::string curve = "bn128";
  int nvars = 0;
  const Fr* values = nullptr;                    // standard form, view into `file`
  std::shared_ptr<detail::FileMap> file;
};
struct Mask { Fr r{}, s{}; };                    // prover.nim:211-213 (standard-form integers)
struct Proof {                                   // prover.nim:38-43
  std::string curve = "bn128";
  std::vector<Fr> publicIO;                      // standard form, publicIO[0] = 1
  G1 pi_a{};
  G2 pi_b{};
  G1 pi_c{};
};

// ------------------------------------------------------------------------------------------------ files
// pin: register the mapping with the CUDA driver (FileMap::pin); also taken from the environment G16_PIN_ZKEY=1
inline ZKey parseZKey(const std::string& fname, bool pin = false) {  // zkey.nim:241-246
  using namespace detail;
  ZKey zk;
  zk.file = std::make_shared<FileMap>(fname);
  if (const char* e = getenv("G16_PIN_ZKEY")) pin = pin || e[0] == '1';
  if (pin) zk.file->pin();
  auto sec = parse_container(zk.file->data(), zk.file->size(), "zkey", 1, 10);
  doAssert(sec[1].len == 4 && rd32(sec[1].p) == 1, "expecting `.zkey` file for a Groth16 prover");   // zkey.nim:110
  const Section& s2 = sec[2];
  doAssert(s2.len == 2 * 4 + 32 + 32 + 3 * 4 + 3 * 64 + 3 * 128, "unexpected section length");      // zkey.nim:122
  doAssert(rd32(s2.p) == 32 && memcmp(s2.p + 4, P_MOD, 32) == 0, "expecting the alt-bn128 curve");   // zkey.nim:134
  doAssert(rd32(s2.p + 36) == 32 && memcmp(s2.p + 40, R_MOD, 32) == 0, "expecting the alt-bn128 curve");
  zk.header.nvars = (int)rd32(s2.p + 72);
  zk.header.npubs = (int)rd32(s2.p + 76);
  zk.header.domainSize = (int)rd32(s2.p + 80);
  int lg = 0;
  while ((1 << lg) < zk.header.domainSize) lg++;
  doAssert((1 << lg) == zk.header.domainSize, "domain size should be a power of two");              // zkey.nim:143
  zk.header.logDomainSize = lg;
  zk.header.flavour = Snarkjs;                                                                      // zkey.nim:129
  const uint8_t* sp = s2.p + 84;                 // alpha1, beta1, beta2, gamma2, delta1, delta2
  memcpy(&zk.specPoints.alpha1, sp, 64);
  memcpy(&zk.specPoints.beta1, sp + 64, 64);
  memcpy(&zk.specPoints.beta2, sp + 128, 128);
  memcpy(&zk.specPoints.gamma2, sp + 256, 128);
  memcpy(&zk.specPoints.delta1, sp + 384, 64);
  memcpy(&zk.specPoints.delta2, sp + 448, 128);
  doAssert(sec[4].len >= 4, "unexpected section length");
  zk.ncoeffs = rd32(sec[4].p);
  doAssert(sec[4].len == 4 + zk.ncoeffs * 44, "unexpected section length");                         // zkey.nim:171
  zk.coeffs = sec[4].p + 4;
  const size_t nv = (size_t)zk.header.nvars, np1 = (size_t)zk.header.npubs + 1;
  doAssert(nv >= np1, "unexpected section length");
  const size_t want[10] = {0, 0, 0, np1 * 64, 0, nv * 64, nv * 64, nv * 128, (nv - np1) * 64,
                           (size_t)zk.header.domainSize * 64};
  for (int id : {3, 5, 6, 7, 8, 9}) doAssert(sec[id].len == want[id], "unexpected section length"); // zkey.nim:198-224
  zk.pointsIC = reinterpret_cast<const G1*>(sec[3].p);
  zk.pointsA1 = reinterpret_cast<const G1*>(sec[5].p);
  zk.pointsB1 = reinterpret_cast<const G1*>(sec[6].p);
  zk.pointsB2 = reinterpret_cast<const G2*>(sec[7].p);
  zk.pointsC1 = reinterpret_cast<const G1*>(sec[8].p);
  zk.pointsH1 = reinterpret_cast<const G1*>(sec[9].p);
  return zk;
}

inline Witness parseWitness(const std::string& fname) {              // witness.nim:71-76
  using namespace detail;
  Witness w;
  w.file = std::make_shared<FileMap>(fname);
  auto sec = parse_container(w.file->data(), w.file->size(), "wtns", 2, 2);
  const Section& s1 = sec[1];
  doAssert(s1.len == 4 + 32 + 4, "unexpected section length");                                      // witness.nim:44
  doAssert(rd32(s1.p) == 32, "expecting 256 bit prime");                                            // witness.nim:46
  doAssert(memcmp(s1.p + 4, R_MOD, 32) == 0, "expecting the alt-bn128 curve");                      // witness.nim:47
  w.nvars = (int)rd32(s1.p + 36);
  doAssert(sec[2].len == (size_t)w.nvars * 32, "unexpected section length");                        // witness.nim:59
  w.values = reinterpret_cast<const Fr*>(sec[2].p);
  return w;
}

// ------------------------------------------------------------------------------------------------ fine-grained procs
// coeffs: field VALUES as Montgomery residues (the reference's in-memory seq[Fr], msm.nim:44 toBig)
inline G1 msmMultiThreadedG1(int /*nthreads_hint*/, const std::vector<Fr>& coeffs, const std::vector<G1>& points) {
  doAssert(coeffs.size() == points.size(), "incompatible sequence lengths");                        // msm.nim:97
  G1 out;
  check(g16_msm_g1(reinterpret_cast<const uint64_t*>(coeffs.data()), G16_FORM_MONT,
                   reinterpret_cast<const uint64_t*>(points.data()), coeffs.size(), reinterpret_cast<uint64_t*>(&out)));
  return out;
}
inline G2 msmMultiThreadedG2(int /*nthreads_hint*/, const std::vector<Fr>& coeffs, const std::vector<G2>& points) {
  doAssert(coeffs.size() == points.size(), "incompatible sequence lengths");                        // msm.nim:136
  G2 out;
  check(g16_msm_g2(reinterpret_cast<const uint64_t*>(coeffs.data()), G16_FORM_MONT,
                   reinterpret_cast<const uint64_t*>(points.data()), coeffs.size(), reinterpret_cast<uint64_t*>(&out)));
  return out;
}
inline int ceilingLog2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }
inline std::vector<Fr> forwardNTT(const std::vector<Fr>& src) {      // ntt.nim:55-77, domain = createDomain(src.len)
  doAssert(src.size() >= 2 && (src.size() & (src.size() - 1)) == 0, "input must have the same size as the domain");
  std::vector<Fr> dst(src.size());
  check(g16_ntt_fr(reinterpret_cast<const uint64_t*>(src.data()), reinterpret_cast<uint64_t*>(dst.data()),
                   ceilingLog2(src.size()), 0));
  return dst;
}
inline std::vector<Fr> inverseNTT(const std::vector<Fr>& src) {      // ntt.nim:139-161
  doAssert(src.size() >= 2 && (src.size() & (src.size() - 1)) == 0, "input must have the same size as the domain");
  std::vector<Fr> dst(src.size());
  check(g16_ntt_fr(reinterpret_cast<const uint64_t*>(src.data()), reinterpret_cast<uint64_t*>(dst.data()),
                   ceilingLog2(src.size()), 1));
  return dst;
}
struct ABC { std::vector<Fr> valuesAz, valuesBz, valuesCz; };        // prover.nim:49-52
inline ABC buildABC(const ZKey& zkey, const Fr* witness_std, size_t nwitness) {   // prover.nim:56-73
  const size_t n = (size_t)zkey.header.domainSize;
  ABC abc{std::vector<Fr>(n), std::vector<Fr>(n), std::vector<Fr>(n)};
  check(g16_build_abc(zkey.coeffs, zkey.ncoeffs, G16_COEFF_PACKED44_R2, reinterpret_cast<const uint64_t*>(witness_std),
                      G16_FORM_STD, nwitness, zkey.header.logDomainSize, reinterpret_cast<uint64_t*>(abc.valuesAz.data()),
                      reinterpret_cast<uint64_t*>(abc.valuesBz.data()), reinterpret_cast<uint64_t*>(abc.valuesCz.data())));
  return abc;
}
inline std::vector<Fr> computeSnarkjsScalarCoeffs(int /*nthreads*/, const ABC& abc) {   // prover.nim:158-181
  std::vector<Fr> qs(abc.valuesAz.size());
  check(g16_quotient(reinterpret_cast<const uint64_t*>(abc.valuesAz.data()),
                     reinterpret_cast<const uint64_t*>(abc.valuesBz.data()), ceilingLog2(qs.size()), G16_FLAVOUR_SNARKJS,
                     reinterpret_cast<uint64_t*>(qs.data())));
  return qs;
}
inline std::vector<Fr> computeQuotientPointwise(int /*nthreads*/, const ABC& abc) {     // prover.nim:118-148
  std::vector<Fr> qs(abc.valuesAz.size());
  check(g16_quotient(reinterpret_cast<const uint64_t*>(abc.valuesAz.data()),
                     reinterpret_cast<const uint64_t*>(abc.valuesBz.data()), ceilingLog2(qs.size()), G16_FLAVOUR_JENSGROTH,
                     reinterpret_cast<uint64_t*>(qs.data())));
  return qs;
}

// ------------------------------------------------------------------------------------------------ resident prover
// generateProofWithMask with the zkey kept in HBM between proofs (window tables + CSR rows built once)
class Prover {
 public:
  // flags: G16_ZKEY_* (TRUSTED skips the on-curve checks of io.nim:228-236; ONE_SHOT keeps plain points);
  // devices = N > 1: the key is spread over N GPUs of this process (also: environment G16_NGPUS)
  explicit Prover(const ZKey& zkey, uint32_t flags = 0, int devices = 0)
      : npubs_(zkey.header.npubs), nvars_(zkey.header.nvars) {
    g16_zkey_view v;
    memset(&v, 0, sizeof(v));
    v.flags = flags;
    v.nvars = (uint32_t)zkey.header.nvars;
    v.npubs = (uint32_t)zkey.header.npubs;
    v.log_domain = (uint32_t)zkey.header.logDomainSize;
    v.flavour = (uint32_t)zkey.header.flavour;
    v.coeff_format = G16_COEFF_PACKED44_R2;
    v.mem_kind = G16_MEM_HOST;
    v.ncoeffs = zkey.ncoeffs;
    v.coeffs = zkey.coeffs;
    v.points_a1 = reinterpret_cast<const uint64_t*>(zkey.pointsA1);
    v.points_b1 = reinterpret_cast<const uint64_t*>(zkey.pointsB1);
    v.points_b2 = reinterpret_cast<const uint64_t*>(zkey.pointsB2);
    v.points_c1 = reinterpret_cast<const uint64_t*>(zkey.pointsC1);
    v.points_h1 = reinterpret_cast<const uint64_t*>(zkey.pointsH1);
    memcpy(v.alpha1, &zkey.specPoints.alpha1, 64);
    memcpy(v.beta1, &zkey.specPoints.beta1, 64);
    memcpy(v.beta2, &zkey.specPoints.beta2, 128);
    memcpy(v.delta1, &zkey.specPoints.delta1, 64);
    memcpy(v.delta2, &zkey.specPoints.delta2, 128);
    check(g16_ctx_create(&v, 0, devices > 1 ? -devices : 1, &ctx_));
  }
  ~Prover() { if (ctx_) g16_ctx_destroy(ctx_); }
  Prover(const Prover&) = delete;
  Prover& operator=(const Prover&) = delete;

  Proof prove(const Witness& wtns, const Mask& mask, g16_stats* stats = nullptr) {
    doAssert(wtns.nvars == nvars_, "wrong witness length");                                        // prover.nim:236
    g16_proof raw;
    check(g16_prove(ctx_, reinterpret_cast<const uint64_t*>(wtns.values), G16_FORM_STD, mask.r.limb, mask.s.limb, &raw,
                    stats));
    Proof prf;
    prf.publicIO.assign(wtns.values, wtns.values + npubs_ + 1);                                    // prover.nim:239-240
    memcpy(&prf.pi_a, raw.pi_a, 64);
    memcpy(&prf.pi_b, raw.pi_b, 128);
    memcpy(&prf.pi_c, raw.pi_c, 64);
    return prf;
  }

 private:
  g16_ctx* ctx_ = nullptr;
  int npubs_, nvars_;
};

// ------------------------------------------------------------------------------------------------ the reference's prover procs
// The reference's signature: the zkey arrives with every call (cli_main.nim:193-210), so the context made here
// serves exactly one proof and is created ONE_SHOT (an upload, no window tables).  A host that proves repeatedly
// against one key keeps a groth16::Prover instead.
inline Proof generateProofWithMask(int /*nthreads*/, bool printTimings, const ZKey& zkey, const Witness& wtns,
                                   const Mask& mask, int devices = 0) {                            // prover.nim:215
  doAssert(zkey.header.curve == wtns.curve, "zkey.header.curve != wtns.curve");                    // prover.nim:224
  doAssert(zkey.header.nvars == wtns.nvars, "wrong witness length");                               // prover.nim:236
  Prover prover(zkey, G16_ZKEY_ONE_SHOT, devices);
  g16_stats st;
  Proof prf = prover.prove(wtns, mask, &st);
  if (printTimings)                                                                                // prover.nim:221
    fprintf(stderr,
            "building 'ABC' %.3f ms | quotient %.3f ms | witness sort %.3f ms | pi_A, rho, C %.3f ms | pi_B (G2) %.3f ms | "
            "H %.3f ms | total %.3f ms (device)\n",
            st.ms_abc, st.ms_quotient, st.ms_sort_witness, st.ms_msm_g1_witness, st.ms_msm_b2, st.ms_msm_h, st.ms_total);
  return prf;
}
inline Proof generateProofWithTrivialMask(int nthreads, bool printTimings, const ZKey& zkey, const Witness& wtns,
                                          int devices = 0) {
  return generateProofWithMask(nthreads, printTimings, zkey, wtns, Mask{}, devices);               // prover.nim:308
}
// The blinding scalars are what makes the proof zero-knowledge: unlike the reference's std/random (rnd.nim) they are
// drawn from the kernel CSPRNG (getrandom), 254 bits with rejection sampling, i.e. uniform in [0, r).
inline Fr randFr() {
  Fr x;
  do {
    size_t got = 0;
    while (got < sizeof(x.limb)) {
      ssize_t k = getrandom(reinterpret_cast<char*>(x.limb) + got, sizeof(x.limb) - got, 0);
      if (k < 0) throw AssertionDefect("getrandom failed");
      got += (size_t)k;
    }
    x.limb[3] &= 0x3fffffffffffffffull;
  } while (!detail::less_than(x.limb, detail::R_MOD));
  return x;
}
inline Proof generateProof(int nthreads, bool printTimings, const ZKey& zkey, const Witness& wtns,
                           int devices = 0) {                                                      // prover.nim:312-319
  Mask m;
  m.r = randFr();
  m.s = randFr();
  return generateProofWithMask(nthreads, printTimings, zkey, wtns, m, devices);
}

// ------------------------------------------------------------------------------------------------ export_json.nim:25-80
inline std::string publicIOJson(const Proof& prf) {
  doAssert(!prf.publicIO.empty(), "empty public IO");
  const uint64_t one[4] = {1, 0, 0, 0};
  doAssert(memcmp(prf.publicIO[0].limb, one, 32) == 0, "the first public input must be 1");        // export_json.nim:30-31
  std::string s;
  for (size_t i = 1; i < prf.publicIO.size(); i++)
    s += std::string(i == 1 ? "[ " : ", ") + "\"" + detail::to_decimal(prf.publicIO[i].limb) + "\"\n";
  s += "] \n";
  return s;
}
inline std::string proofJson(const Proof& prf) {
  using detail::fp_decimal;
  auto g1 = [](const G1& p) {
    return "    [ \"" + fp_decimal(p.x) + "\"\n    , \"" + fp_decimal(p.y) + "\"\n    , \"1\"\n    ]\n";
  };
  auto fp2 = [](const char* lead, const std::string& a, const std::string& b) {
    return std::string("    ") + lead + " [ \"" + a + "\"\n      , \"" + b + "\"\n      ]\n";
  };
  std::string s = "{ \"protocol\": \"groth16\"\n, \"curve\":    \"bn128\"\n, \"pi_a\":\n" + g1(prf.pi_a) + ", \"pi_b\":\n";
  s += fp2("[", fp_decimal(prf.pi_b.x.c0), fp_decimal(prf.pi_b.x.c1));
  s += fp2(",", fp_decimal(prf.pi_b.y.c0), fp_decimal(prf.pi_b.y.c1));
  s += fp2(",", "1", "0");
  s += "    ]\n, \"pi_c\":\n" + g1(prf.pi_c) + "}\n";
  return s;
}
inline void writeText(const std::string& fpath, const std::string& text) {
  std::ofstream f(fpath);
  if (!f) throw AssertionDefect("cannot write file `" + fpath + "`");
  f << text;
}
inline void exportPublicIO(const std::string& fpath, const Proof& prf) { writeText(fpath, publicIOJson(prf)); }
inline void exportProof(const std::string& fpath, const Proof& prf) { writeText(fpath, proofJson(prf)); }

}  // namespace groth16
