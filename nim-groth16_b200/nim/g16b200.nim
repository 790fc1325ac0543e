#
# g16b200.nim -- `importc` shim that puts the B200 (sm_100a) CUDA backend behind nim-groth16's own procs.
#
# Drop this file into the reference as `groth16/g16b200.nim`, link with `-lg16b200`, and replace the bodies of
#   groth16/bn128/msm.nim:   msmMultiThreadedG1 (:89), msmMultiThreadedG2 (:128)
#   groth16/math/ntt.nim:    forwardNTT (:55), inverseNTT (:139)
#   groth16/prover.nim:      buildABC (:56), computeSnarkjsScalarCoeffs (:158), computeQuotientPointwise (:118),
#                            generateProofWithMask (:215)
# with the one-line calls shown at the bottom (INTEGRATION.md has the full diff).  Every signature, the
# `Proof` / `Mask` / `ZKey` / `Witness` types, the CLI and verifier.nim stay untouched.
#
# NOTE: there is no Nim toolchain in the build environment of this repository, so this file has not been
# compiled; it is kept deliberately small and mechanical.  The identical C ABI is exercised by the Python
# ctypes binding (nim-groth16_b200/g16b200/_lib.py) in the test-suite.
#
# Layout assumptions (SURVEY.md 8b), asserted at start-up by `g16CheckLayout`:
#   Fr / Fp        = 4 x uint64 little-endian limbs, Montgomery residue, R = 2^256   (constantine BigInt[254])
#   G1 = (x, y)    = 64 bytes;  G2 = (x.c0, x.c1, y.c0, y.c1) = 128 bytes; infinity = all zero (curves.nim:49-50)
#

import constantine/math/arithmetic except Fp, Fr
import constantine/math/io/io_bigints

import groth16/bn128
import groth16/zkey_types
import groth16/files/witness
import std/tables

{.passL: "-lg16b200".}

const
  G16_FORM_MONT* = 0.cint
  G16_FORM_STD*  = 1.cint
  G16_COEFF_STRUCT48_MONT = 1'u32
  G16_MEM_HOST = 0'u32

type
  G16Coeff {.bycopy.} = object          # g16_coeff  (include/g16b200.h)
    matrix, row, col, reserved: uint32
    value: array[4, uint64]

  G16ZkeyView {.bycopy.} = object       # g16_zkey_view
    nvars, npubs, logDomain, flavour, coeffFormat, memKind: uint32
    flags, reserved0: uint32            # G16_ZKEY_TRUSTED = 1, G16_ZKEY_ONE_SHOT = 2
    ncoeffs: uint64
    coeffs: pointer
    pointsA1, pointsB1, pointsB2, pointsC1, pointsH1: pointer
    alpha1, beta1: array[8, uint64]
    beta2: array[16, uint64]
    delta1: array[8, uint64]
    delta2: array[16, uint64]

  G16ProofRaw {.bycopy.} = object       # g16_proof
    piA: array[8, uint64]
    piB: array[16, uint64]
    piC: array[8, uint64]

  G16Stats {.bycopy.} = object          # g16_stats
    msH2d, msAbc, msQuotient, msSortWitness, msMsmG1Witness, msMsmB2, msMsmH, reservedMs, msAssemble, msTotal: cfloat
    kernelLaunches, reserved: uint32

  G16Ctx = distinct pointer

proc g16_last_error(): cstring {.importc, cdecl.}
proc g16_msm_g1(scalars: pointer, form: cint, points: pointer, n: csize_t, res: pointer): cint {.importc, cdecl.}
proc g16_msm_g2(scalars: pointer, form: cint, points: pointer, n: csize_t, res: pointer): cint {.importc, cdecl.}
proc g16_ntt_fr(src, dst: pointer, logN, inverse: cint): cint {.importc, cdecl.}
proc g16_quotient(az, bz: pointer, logN, flavour: cint, qs: pointer): cint {.importc, cdecl.}
proc g16_build_abc(coeffs: pointer, nnz: csize_t, coeffFormat: cint, witness: pointer, witnessForm: cint,
                   m: csize_t, logN: cint, az, bz, cz: pointer): cint {.importc, cdecl.}
proc g16_ctx_create(zk: ptr G16ZkeyView, shardIndex, shardCount: cint, ctx: ptr G16Ctx): cint {.importc, cdecl.}
proc g16_ctx_destroy(ctx: G16Ctx) {.importc, cdecl.}
proc g16_prove(ctx: G16Ctx, witness: pointer, witnessForm: cint, r, s: pointer, proof: ptr G16ProofRaw,
               stats: ptr G16Stats): cint {.importc, cdecl.}
# throughput-oriented hosts: several proofs in flight over one resident key
proc g16_ctx_clone(ctx: G16Ctx, res: ptr G16Ctx): cint {.importc, cdecl.}
proc g16_prove_submit(ctx: G16Ctx, witness: pointer, witnessForm, witnessMemKind: cint, r, s: pointer): cint {.importc, cdecl.}
proc g16_prove_wait(ctx: G16Ctx, proof: ptr G16ProofRaw, stats: ptr G16Stats): cint {.importc, cdecl.}
# multi-GPU, one process: g16_ctx_create(zk, 0, -N) (or G16_NGPUS=N) and the calls above -- nothing else changes.
# multi-GPU, one process per device: one context per rank, 400-byte partial records exchanged by the host
proc g16_ctx_set_mask(ctx: G16Ctx, r, s: pointer): cint {.importc, cdecl.}
proc g16_prove_partials(ctx: G16Ctx, witness: pointer, witnessForm, witnessMemKind: cint, partialsDev: pointer,
                        stats: ptr G16Stats): cint {.importc, cdecl.}
proc g16_prove_finish(ctx: G16Ctx, gatheredPartialsDev: pointer, count: cint, r, s: pointer,
                      proof: ptr G16ProofRaw): cint {.importc, cdecl.}

# the reference signals every failure with assert()/AssertionDefect (msm.nim:97, prover.nim:224,236,270-276)
template check(status: cint) =
  if status != 0:
    raise newException(AssertionDefect, "g16b200: " & $g16_last_error())

proc payload[T](xs: seq[T]): pointer =
  (if xs.len == 0: nil else: unsafeAddr xs[0])

proc g16CheckLayout*() =
  ## raw limbs of oneFr / oneFp must be the Montgomery constants of io.nim:87,91
  doAssert sizeof(Fr) == 32 and sizeof(Fp) == 32 and sizeof(G1) == 64 and sizeof(G2) == 128
  var one = oneFr
  doAssert cast[ptr array[4, uint64]](addr one)[][0] == 0xac96341c4ffffffb'u64   # low limb of frMontR

#-------------------------------------------------------------------------------
# fine-grained replacements
#-------------------------------------------------------------------------------

proc msmMultiThreadedG1*(nthreads_hint: int, coeffs: seq[Fr], points: seq[G1]): G1 =   # msm.nim:89
  assert(coeffs.len == points.len, "incompatible sequence lengths")
  check g16_msm_g1(payload(coeffs), G16_FORM_MONT, payload(points), csize_t(coeffs.len), addr result)

proc msmMultiThreadedG2*(nthreads_hint: int, coeffs: seq[Fr], points: seq[G2]): G2 =   # msm.nim:128
  assert(coeffs.len == points.len, "incompatible sequence lengths")
  check g16_msm_g2(payload(coeffs), G16_FORM_MONT, payload(points), csize_t(coeffs.len), addr result)

proc g16ForwardNTT*(src: seq[Fr], logDomainSize: int): seq[Fr] =                        # ntt.nim:55
  result = newSeq[Fr](src.len)
  check g16_ntt_fr(payload(src), payload(result), cint(logDomainSize), 0)

proc g16InverseNTT*(src: seq[Fr], logDomainSize: int): seq[Fr] =                        # ntt.nim:139
  result = newSeq[Fr](src.len)
  check g16_ntt_fr(payload(src), payload(result), cint(logDomainSize), 1)

proc g16Quotient*(valuesAz, valuesBz: seq[Fr], logDomainSize: int, flavour: Flavour): seq[Fr] =
  ## computeSnarkjsScalarCoeffs (prover.nim:158) / computeQuotientPointwise (prover.nim:118)
  result = newSeq[Fr](valuesAz.len)
  check g16_quotient(payload(valuesAz), payload(valuesBz), cint(logDomainSize), cint(ord(flavour)), payload(result))

proc packCoeffs(coeffs: seq[Coeff]): seq[G16Coeff] =
  result = newSeq[G16Coeff](coeffs.len)
  for i, c in coeffs:
    result[i].matrix = uint32(ord(c.matrix))
    result[i].row    = uint32(c.row)
    result[i].col    = uint32(c.col)
    copyMem(addr result[i].value, unsafeAddr c.coeff, 32)

#-------------------------------------------------------------------------------
# coarse replacement: resident context = generateProofWithMask (prover.nim:215-304)
#-------------------------------------------------------------------------------

type
  G16Prover* = ref object
    ctx: G16Ctx
    npubs: int

proc close*(p: G16Prover) =
  if pointer(p.ctx) != nil:
    g16_ctx_destroy(p.ctx)
    p.ctx = G16Ctx(nil)

const
  G16_ZKEY_TRUSTED*  = 1'u32   # skip the on-curve checks (the reference's loader already ran them, io.nim:228-236)
  G16_ZKEY_ONE_SHOT* = 2'u32   # plain points instead of window tables: context creation is an upload

proc newG16Prover*(zkey: ZKey, flags: uint32 = G16_ZKEY_TRUSTED, gpus: int = 1): G16Prover =
  ## gpus > 1 (or the environment variable G16_NGPUS): the key is spread over that many devices of this process,
  ## everything else -- `prove`, `close` -- is unchanged.
  g16CheckLayout()
  let packed = packCoeffs(zkey.coeffs)
  var v: G16ZkeyView
  v.nvars = uint32(zkey.header.nvars)
  v.npubs = uint32(zkey.header.npubs)
  v.logDomain = uint32(zkey.header.logDomainSize)
  v.flavour = uint32(ord(zkey.header.flavour))
  v.coeffFormat = G16_COEFF_STRUCT48_MONT
  v.memKind = G16_MEM_HOST
  v.ncoeffs = uint64(packed.len)
  v.coeffs = payload(packed)
  v.pointsA1 = payload(zkey.pPoints.pointsA1)
  v.pointsB1 = payload(zkey.pPoints.pointsB1)
  v.pointsB2 = payload(zkey.pPoints.pointsB2)
  v.pointsC1 = payload(zkey.pPoints.pointsC1)
  v.pointsH1 = payload(zkey.pPoints.pointsH1)
  copyMem(addr v.alpha1, unsafeAddr zkey.specPoints.alpha1, 64)
  copyMem(addr v.beta1,  unsafeAddr zkey.specPoints.beta1,  64)
  copyMem(addr v.beta2,  unsafeAddr zkey.specPoints.beta2, 128)
  copyMem(addr v.delta1, unsafeAddr zkey.specPoints.delta1, 64)
  copyMem(addr v.delta2, unsafeAddr zkey.specPoints.delta2, 128)
  v.flags = flags
  new(result)
  result.npubs = zkey.header.npubs
  check g16_ctx_create(addr v, 0, (if gpus > 1: cint(-gpus) else: cint(1)), addr result.ctx)

# One resident prover per zkey: generateProofWithMask receives the ZKey with every call (prover.nim:215), the
# device copy must not be rebuilt each time.  Keyed by the address of the H-point payload, which identifies a loaded
# ZKey for as long as it is alive.  A key starts on a ONE_SHOT context (an upload: ~80 ms at 2^20 including its first
# proof); the window tables cost ~0.5 s once and save ~6 ms per proof, i.e. they pay off after ~90 proofs, so the key
# is promoted to a resident context at its 64th proof (the classic rent-or-buy rule: never worse than twice the best
# choice made with hindsight).  A host that knows it will prove many times calls newG16Prover(zkey) itself.
var g16Cache {.threadvar.}: Table[pointer, tuple[prover: G16Prover, uses: int, resident: bool]]

proc g16ProverFor*(zkey: ZKey): G16Prover =
  let key = payload(zkey.pPoints.pointsH1)
  if key notin g16Cache:
    g16Cache[key] = (newG16Prover(zkey, G16_ZKEY_TRUSTED or G16_ZKEY_ONE_SHOT), 0, false)
  var e = g16Cache[key]
  inc e.uses
  if e.uses == 64 and not e.resident:
    e.prover.close()
    e.prover = newG16Prover(zkey, G16_ZKEY_TRUSTED)
    e.resident = true
  g16Cache[key] = e
  e.prover

proc g16ReleaseProvers*() =
  for e in g16Cache.mvalues: e.prover.close()
  g16Cache.clear()

proc prove*(p: G16Prover, witness: seq[Fr], mask_r, mask_s: Fr): (G1, G2, G1) =
  ## the masks travel as plain integers (toBig), the witness as the in-memory Montgomery seq[Fr]
  var raw: G16ProofRaw
  var r = mask_r.toBig()
  var s = mask_s.toBig()
  check g16_prove(p.ctx, payload(witness), G16_FORM_MONT, addr r, addr s, addr raw, nil)
  copyMem(addr result[0], addr raw.piA, 64)
  copyMem(addr result[1], addr raw.piB, 128)
  copyMem(addr result[2], addr raw.piC, 64)

#-------------------------------------------------------------------------------
# What changes inside the reference (each body becomes one call):
#
#   # groth16/prover.nim:215
#   proc generateProofWithMask*( nthreads: int, printTimings: bool, zkey: ZKey, wtns: Witness, mask: Mask ): Proof =
#     assert( zkey.header.curve == wtns.curve )                          # prover.nim:224 (kept)
#     assert( zkey.header.nvars == wtns.values.len , "wrong witness length" )   # prover.nim:236 (kept)
#     let prover = g16ProverFor(zkey)          # cached per ZKey: the key stays resident in HBM between proofs
#     let (pi_a, pi_b, pi_c) = prover.prove(wtns.values, mask.r, mask.s)
#     var pubIO = newSeq[Fr](zkey.header.npubs + 1)
#     for i in 0..zkey.header.npubs: pubIO[i] = wtns.values[i]           # prover.nim:239-240
#     return Proof( curve:"bn128", publicIO:pubIO, pi_a:pi_a, pi_b:pi_b, pi_c:pi_c )
#-------------------------------------------------------------------------------
