"""The C++ host side over the C ABI (nim-groth16_b200/cpp/g16b200.hpp, mirror of the reference's procs) and its
command line g16prove (the prover half of cli/cli_main.nim): parsing and error behaviour on the CPU, proofs and
the snarkjs JSON export on the GPU against the golden fixture of the reference's own test circuit."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nim-groth16_b200")
EXE = os.path.join(PKG, "cpp", "g16prove")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(os.path.join(PKG, "libg16b200.so")):
        pytest.skip("libg16b200.so not built")
    # the same command as the Makefile target, run directly so that file times never trigger an nvcc rebuild here
    srcs = [os.path.join(PKG, "cpp", "g16prove.cpp"), os.path.join(PKG, "cpp", "g16b200.hpp"),
            os.path.join(ROOT, "include", "g16b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(p) for p in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, srcs[0], "-L" + PKG, "-lg16b200",
                               "-Wl,-rpath,$ORIGIN/.."])
    return EXE


@pytest.fixture()
def files(tmp_path, kat):
    z, w = tmp_path / "circuit.zkey", tmp_path / "witness.wtns"
    z.write_bytes(bytes.fromhex(kat["snarkjs"]["zkey_hex"]))
    w.write_bytes(bytes.fromhex(kat["wtns_hex"]))
    return str(z), str(w), tmp_path


def _run(*args):
    return subprocess.run(list(args), capture_output=True, text=True, timeout=600)


def test_cli_parses_the_snarkjs_containers_like_the_python_mirror(exe, files):
    from g16b200 import files as pf
    z, w, _ = files
    zk = pf.parse_zkey(z)
    r = _run(exe, "-z", z, "-w", w, "--info")
    assert r.returncode == 0, r.stderr
    assert ("nvars=%d npubs=%d domainSize=%d logDomainSize=%d ncoeffs=%d"
            % (zk.nvars, zk.npubs, zk.domainSize, zk.logDomainSize, zk.coeffs.shape[0])) in r.stdout
    assert "wtns: curve=bn128 nvars=%d" % zk.nvars in r.stdout


def test_cli_error_behaviour(exe, files):
    z, w, tmp = files
    bad = tmp / "bad.zkey"
    bad.write_bytes(b"zkex" + bytes(64))
    r = _run(exe, "-z", str(bad), "-w", w, "--info")
    assert r.returncode == 1 and "fatal error" in r.stderr
    data = bytearray(open(z, "rb").read())
    trunc = tmp / "trunc.zkey"
    trunc.write_bytes(bytes(data[: len(data) // 2]))
    r = _run(exe, "-z", str(trunc), "-w", w, "--info")
    assert r.returncode == 1 and "truncated" in r.stderr
    r = _run(exe, "-z", z, "-w", w, "--mask-r", "zz")
    assert r.returncode == 1 and "hexadecimal" in r.stderr
    r = _run(exe, "-z", z, "-w", z, "--info")                       # a zkey where a witness is expected
    assert r.returncode == 1
    assert _run(exe, "--help").returncode == 0 and _run(exe).returncode == 2


def test_host_montgomery_to_decimal(exe):
    """detail::from_mont + to_decimal of g16b200.hpp (what exportProof prints) against Python integers."""
    import random
    import g16_oracle as o
    rnd = random.Random(5)
    vals = [0, 1, o.P - 1, 3] + [rnd.randrange(o.P) for _ in range(20)]
    for v in vals:
        mont = v * (1 << 256) % o.P
        r = _run(exe, "--to-decimal", "%x" % mont)
        assert r.returncode == 0 and r.stdout.strip() == str(v), (v, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cli_proof_and_json_export_match_the_golden_fixture(exe, files, kat):
    """generateProofWithMask / generateProofWithTrivialMask + exportProof / exportPublicIO through the C++ host:
    byte-identical files to the Python mirror's export of the golden proofs (fixed masks and no mask)."""
    import numpy as np
    from g16b200 import encoding as e, export_json
    from g16b200.zkey_types import Proof
    z, w, tmp = files

    def golden(which):
        g = kat["snarkjs"][which]
        h = lambda v: int(v, 16)
        pa = e.g1_array([(h(g["pi_a"][0]), h(g["pi_a"][1]))])[0]
        pb = e.g2_array([((h(g["pi_b"][0][0]), h(g["pi_b"][0][1])), (h(g["pi_b"][1][0]), h(g["pi_b"][1][1])))])[0]
        pc = e.g1_array([(h(g["pi_c"][0]), h(g["pi_c"][1]))])[0]
        pub = e.fr_std([int(v, 16) if isinstance(v, str) else int(v) for v in kat["witness"][:3]])
        return Proof(publicIO=pub, pi_a=pa, pi_b=pb, pi_c=pc)

    out, io = str(tmp / "proof.json"), str(tmp / "public.json")
    r = _run(exe, "-p", "-z", z, "-w", w, "-o", out, "-i", io, "-t", "--mask-r", kat["mask"]["r"], "--mask-s",
             kat["mask"]["s"])
    assert r.returncode == 0, r.stderr
    assert "total" in r.stderr                                        # -t prints the phase timings
    assert open(out).read() == export_json.proof_json(golden("fixed"))
    assert open(io).read() == export_json.public_io_json(golden("fixed"))
    assert json.loads(open(out).read())["protocol"] == "groth16"      # and it is valid JSON
    r = _run(exe, "-z", z, "-w", w, "-o", out, "-i", io, "-n")
    assert r.returncode == 0, r.stderr
    assert open(out).read() == export_json.proof_json(golden("trivial"))
    # random masks (generateProof): a different, still well-formed proof with the same public inputs
    r = _run(exe, "-z", z, "-w", w, "-o", out, "-i", io)
    assert r.returncode == 0, r.stderr
    assert open(out).read() != export_json.proof_json(golden("trivial"))
    assert open(io).read() == export_json.public_io_json(golden("fixed"))
    # a witness of the wrong length is refused like prover.nim:236
    from g16b200 import files as pf
    from g16b200.zkey_types import Witness
    short = tmp / "short.wtns"
    pf.write_witness(str(short), Witness(values=pf.parse_witness(w).values[:-1]))
    r = _run(exe, "-z", z, "-w", str(short), "-n")
    assert r.returncode == 1 and "wrong witness length" in r.stderr


@pytest.mark.gpu
def test_cli_debug_intermediates_match_the_golden_fixture(exe, files, kat):
    """The fine-grained procs of g16b200.hpp (buildABC, computeSnarkjsScalarCoeffs, forwardNTT / inverseNTT,
    msmMultiThreadedG1, g16_msm_g2) called one by one from C++: Az/Bz/Cz, qs and the MSM results of the golden
    fixture (SURVEY.md Appendix C values)."""
    z, w, _ = files
    r = _run(exe, "-z", z, "-w", w, "-d")
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    g = kat["snarkjs"]
    h = lambda v: int(v, 16)
    for key in ("Az", "Bz", "Cz", "qs"):
        assert [h(v) for v in got[key]] == [h(v) for v in g[key]], key
    assert got["ntt_roundtrip"] is True
    assert [int(v) for v in got["msmH"]] == [h(v) for v in g["fixed"]["msmH"]]
    assert [[int(a) for a in c] for c in got["msmB2"]] == [[h(a) for a in c] for c in g["fixed"]["msmB2"]]


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 4])
def test_cli_multi_gpu_reproduces_the_golden_proof(exe, files, kat, gpus):
    """g16prove --gpus N: the compiled host calls generateProofWithMask unchanged and the library spreads the proof
    over N devices of the process (g16_ctx_create with shard_count = -N).  Same bytes as the golden proof.  With
    fewer than N GPUs on the box the shards share the devices that exist (G16_DEVICES), same code path."""
    from g16b200 import export_json, _lib
    import ctypes as C
    z, w, tmp = files
    n = C.c_int()
    _lib.check(_lib.load().g16_device_count(C.byref(n)))
    env = dict(os.environ, G16_DEVICES=",".join(str(k % n.value) for k in range(gpus)))
    out, io = str(tmp / "proof_m.json"), str(tmp / "public_m.json")
    one, io1 = str(tmp / "proof_1.json"), str(tmp / "public_1.json")
    args = ["-z", z, "-w", w, "--mask-r", kat["mask"]["r"], "--mask-s", kat["mask"]["s"]]
    r = subprocess.run([exe] + args + ["-o", one, "-i", io1], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe] + args + ["-o", out, "-i", io, "--gpus", str(gpus)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(one).read() and open(io).read() == open(io1).read()
    # the mmap'ed zkey page-locked for the upload (g16_host_register; falls back silently where the platform refuses)
    r = subprocess.run([exe] + args + ["-o", out, "-i", io], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, G16_PIN_ZKEY="1"))
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(one).read()
    # and through the environment alone (the Nim shim's route: nothing but G16_NGPUS changes)
    env2 = dict(env, G16_NGPUS=str(gpus))
    r = subprocess.run([exe] + args + ["-o", out, "-i", io], capture_output=True, text=True, timeout=600, env=env2)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(one).read()
