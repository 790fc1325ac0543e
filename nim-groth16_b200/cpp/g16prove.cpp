// g16prove -- the prover half of nim-groth16's command line (cli/cli_main.nim: -p -z -w -o -i -n -t) on top of
// g16b200.hpp.  Verification, setup and the debug switches stay with the reference's own CLI.
//
//   g16prove -z circuit.zkey -w witness.wtns [-o proof.json] [-i public.json] [-n] [-t]
//            [--mask-r HEX --mask-s HEX]     fixed blinding scalars (testing; default: random, -n: none)
//            [--gpus N]                      spread the key over N GPUs of this process (the library's in-library
//                                            multi-GPU context; also: environment G16_NGPUS)
//            [--info]                        parse the inputs and print their headers only (needs no GPU)
//            [-d | --debug]                  print the intermediates of the fine-grained procs (buildABC, the
//                                            quotient, forward/inverse NTT, the H and pi_B MSMs) as JSON on stdout
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>

#include "g16b200.hpp"

using namespace groth16;

static void printHelp() {
  printf("usage: g16prove -z <file.zkey> -w <file.wtns> [-o <proof.json>] [-i <public.json>] [-n] [-t]\n"
         "                [--mask-r <hex> --mask-s <hex>] [--gpus <n>] [--info]\n"
         " -z, --zkey      the circuit's proving key (snarkjs .zkey, Groth16, bn128)\n"
         " -w, --wtns      the witness (.wtns)\n"
         " -o, --output    where to write the proof (default: proof.json)\n"
         " -i, --io        where to write the public inputs/outputs (default: public.json)\n"
         " -n, --nomask    no masking (r = s = 0)\n"
         " -t, --time      print timings\n"
         " -d, --debug     print the intermediates (Az/Bz/Cz, qs, MSM(qs,H), MSM(w,B2), an NTT round trip) as JSON\n"
         "     --gpus      number of GPUs of this box to spread the proof over (default 1)\n"
         "     --info      print the headers of the inputs and exit (no GPU needed)\n");
}

static Fr parseHexFr(const char* s) {
  Fr x{};
  if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
  size_t n = strlen(s);
  doAssert(n >= 1 && n <= 64, "mask must be 1..64 hex digits");
  for (size_t i = 0; i < n; i++) {
    char c = s[n - 1 - i];
    int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
    doAssert(d >= 0, "mask is not hexadecimal");
    x.limb[i / 16] |= (uint64_t)d << (4 * (i % 16));
  }
  return x;
}

int main(int argc, char** argv) {
  std::string zkey_file, wtns_file, out_file = "proof.json", io_file = "public.json";
  bool nomask = false, timing = false, info = false, debug = false, have_r = false, have_s = false;
  Mask mask;
  int gpus = 0;
  try {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      auto value = [&]() -> const char* {
        doAssert(i + 1 < argc, "missing value for an option");
        return argv[++i];
      };
      if (a == "-h" || a == "--help") { printHelp(); return 0; }
      else if (a == "-z" || a == "--zkey") zkey_file = value();
      else if (a == "-w" || a == "--wtns" || a == "--witness") wtns_file = value();
      else if (a == "-o" || a == "--output") out_file = value();
      else if (a == "-i" || a == "--io" || a == "--input") io_file = value();
      else if (a == "-n" || a == "--nomask") nomask = true;
      else if (a == "-t" || a == "--time") timing = true;
      else if (a == "-p" || a == "--prove") {}                    // the only action of this tool
      else if (a == "--mask-r" || a == "--mask-s") {
        Fr x = parseHexFr(value());
        doAssert(detail::less_than(x.limb, detail::R_MOD), "mask must be below the group order");
        if (a == "--mask-r") { mask.r = x; have_r = true; }
        else { mask.s = x; have_s = true; }
      }
      else if (a == "--gpus") gpus = atoi(value());
      else if (a == "--info") info = true;
      else if (a == "--to-decimal") {               // host arithmetic check: Montgomery Fp hex -> decimal string
        Fr x = parseHexFr(value());
        Fp m;
        memcpy(&m, &x, 32);
        printf("%s\n", detail::fp_decimal(m).c_str());
        return 0;
      }
      else if (a == "-d" || a == "--debug") debug = true;
      else throw AssertionDefect("unknown option `" + a + "`");
    }
    if (zkey_file.empty() || wtns_file.empty()) { printHelp(); return 2; }
    auto t0 = std::chrono::steady_clock::now();
    ZKey zkey = parseZKey(zkey_file);
    Witness wtns = parseWitness(wtns_file);
    auto t1 = std::chrono::steady_clock::now();
    if (info) {
      printf("zkey: curve=%s flavour=%s nvars=%d npubs=%d domainSize=%d logDomainSize=%d ncoeffs=%zu\n",
             zkey.header.curve.c_str(), zkey.header.flavour == Snarkjs ? "Snarkjs" : "JensGroth", zkey.header.nvars,
             zkey.header.npubs, zkey.header.domainSize, zkey.header.logDomainSize, zkey.ncoeffs);
      printf("wtns: curve=%s nvars=%d\n", wtns.curve.c_str(), wtns.nvars);
      return 0;
    }
    if (debug) {                                   // the fine-grained procs, one call each (prover.nim:245-301)
      auto hexFr = [](const Fr& m) {               // Montgomery -> standard -> hex
        uint64_t v[4];
        detail::from_mont(m.limb, detail::R_MOD, detail::R_INV, v);
        char buf[80];
        snprintf(buf, sizeof buf, "\"0x%016llx%016llx%016llx%016llx\"", (unsigned long long)v[3], (unsigned long long)v[2],
                 (unsigned long long)v[1], (unsigned long long)v[0]);
        return std::string(buf);
      };
      auto list = [&](const std::vector<Fr>& xs) {
        std::string s = "[";
        for (size_t i = 0; i < xs.size(); i++) s += (i ? ", " : "") + hexFr(xs[i]);
        return s + "]";
      };
      ABC abc = buildABC(zkey, wtns.values, (size_t)wtns.nvars);
      std::vector<Fr> qs = computeSnarkjsScalarCoeffs(0, abc);
      std::vector<Fr> back = forwardNTT(inverseNTT(abc.valuesAz));
      bool roundtrip = memcmp(back.data(), abc.valuesAz.data(), back.size() * sizeof(Fr)) == 0;
      std::vector<G1> hpts(zkey.pointsH1, zkey.pointsH1 + zkey.header.domainSize);
      G1 msmH = msmMultiThreadedG1(0, qs, hpts);
      // pi_B's MSM straight through the C entry point: the .wtns values are standard-form integers
      std::vector<G2> b2(zkey.pointsB2, zkey.pointsB2 + zkey.header.nvars);
      G2 msmB2;
      check(g16_msm_g2(reinterpret_cast<const uint64_t*>(wtns.values), G16_FORM_STD,
                       reinterpret_cast<const uint64_t*>(b2.data()), b2.size(), reinterpret_cast<uint64_t*>(&msmB2)));
      printf("{ \"Az\": %s,\n  \"Bz\": %s,\n  \"Cz\": %s,\n  \"qs\": %s,\n  \"ntt_roundtrip\": %s,\n",
             list(abc.valuesAz).c_str(), list(abc.valuesBz).c_str(), list(abc.valuesCz).c_str(), list(qs).c_str(),
             roundtrip ? "true" : "false");
      printf("  \"msmH\": [\"%s\", \"%s\"],\n", detail::fp_decimal(msmH.x).c_str(), detail::fp_decimal(msmH.y).c_str());
      printf("  \"msmB2\": [[\"%s\", \"%s\"], [\"%s\", \"%s\"]] }\n", detail::fp_decimal(msmB2.x.c0).c_str(),
             detail::fp_decimal(msmB2.x.c1).c_str(), detail::fp_decimal(msmB2.y.c0).c_str(),
             detail::fp_decimal(msmB2.y.c1).c_str());
      return 0;
    }
    Proof prf;
    if (nomask) prf = generateProofWithTrivialMask(0, timing, zkey, wtns, gpus);
    else if (have_r || have_s) prf = generateProofWithMask(0, timing, zkey, wtns, mask, gpus);
    else prf = generateProof(0, timing, zkey, wtns, gpus);
    auto t2 = std::chrono::steady_clock::now();
    exportProof(out_file, prf);
    exportPublicIO(io_file, prf);
    if (timing) {
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      fprintf(stderr, "parsing the zkey and the witness: %.1f ms; context + proof: %.1f ms\n", ms(t0, t1), ms(t1, t2));
    }
    return 0;
  } catch (const AssertionDefect& e) {
    fprintf(stderr, "fatal error: %s\n", e.what());
    return 1;
  }
}
