"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel family of the library at
sizes that finish under the sanitizer -- NTT through the radix-8 register kernel and the radix-2 one, both quotient
flavours, G1 / G2 MSM (finer work items, fixups), a validated context, a one-shot context, a three-shard in-library
context, sharded contexts with masked records.  Results are checked against each other, not against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200"))
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib, encoding as E
lib = _lib.load()
for lg in (5, 11, 12, 14):
    x = E.fr_mont(E.fr_from_std(E.random_fr_std(1 << lg, 6)))
    D = g.create_domain(1 << lg)
    y = g.forward_ntt(x, D)
    assert np.array_equal(g.inverse_ntt(y, D), x), lg
    az, bz = x, y
    for q in (g.compute_snarkjs_scalar_coeffs, g.compute_quotient_pointwise):
        q(1, az, bz)
print("ntt/quotient ok", flush=True)
n = 6000
sc = E.random_fr_std(n, 4)
for g2 in (0, 1):
    pts = (g.fixed_base_g2 if g2 else g.fixed_base_g1)(E.random_fr_std(n, 5))
    a = (g.msm_multi_threaded_g2 if g2 else g.msm_multi_threaded_g1)(0, sc, pts, form=E.FORM_STD)
    b = (g.msm_multi_threaded_g2 if g2 else g.msm_multi_threaded_g1)(0, sc[::-1].copy(), pts[::-1].copy(), form=E.FORM_STD)
    assert np.array_equal(a, b)
print("msm ok", flush=True)
r1cs, wit = g.synthetic_chain_circuit(1500, seed=3)
zk, _ = g.fake_circuit_setup(r1cs, g.ToxicWaste(11, 22, 33, 44, 55), 1)
wit = np.ascontiguousarray(wit)
m = g.Mask(12345678901234567890123, 98765432109876543210987)
c = g.ProverContext(zk)
want = c.prove(wit, m)
c.close()
c = g.ProverContext(zk, one_shot=True)
got = c.prove(wit, m)
c.close()
assert np.array_equal(got.pi_c, want.pi_c) and np.array_equal(got.pi_b, want.pi_b)
os.environ["G16_DEVICES"] = "0,0,0"
c = g.ProverContext(zk, devices=3, trusted=True)
got = c.prove(wit, m)
c.close()
assert np.array_equal(got.pi_c, want.pi_c) and np.array_equal(got.pi_a, want.pi_a)
G = 4
ctxs = [g.ProverContext(zk, k, G) for k in range(G)]
parts = torch.zeros((G, _lib.PARTIALS_BYTES), dtype=torch.uint8, device="cuda")
for k, ctx in enumerate(ctxs):
    ctx.set_mask(m)
    ctx.prove_partials(wit.ctypes.data, E.FORM_STD, 0, parts[k].data_ptr())
raw = ctxs[0].prove_finish(parts.data_ptr(), G, m)
assert bytes(raw.pi_c) == want.pi_c.tobytes()
for ctx in ctxs:
    ctx.close()
print("prover ok", flush=True)
