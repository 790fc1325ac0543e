import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import g16b200 as g
import g16_oracle as o
from g16b200 import _lib, encoding as E
lib = _lib.load()
seed = 40
for neqs in (1000, 5000):
    r1cs, wit = g.synthetic_chain_circuit(neqs, seed=3)
    tox_o = o.ToxicWaste(*[o.Rng(seed + i).fr() or 1 for i in range(5)])
    tox = g.ToxicWaste(tox_o.alpha, tox_o.beta, tox_o.gamma, tox_o.delta, tox_o.tau)
    zk, sc = g.fake_circuit_setup(r1cs, tox, 1, want_scalars=True)
    w_i = E.fr_from_std(wit); b_i = E.fr_from_std(sc.b); a_i = E.fr_from_std(sc.a)
    mB = sum(a*b for a,b in zip(w_i,b_i)) % o.R
    r, s = o.Rng(seed+10).fr(), o.Rng(seed+11).fr()
    wantB = o.g2_mul((tox_o.beta + s * tox_o.delta + mB) % o.R, o.GEN2)
    wantM = o.g2_mul(mB, o.GEN2)
    parts = torch.zeros(400, dtype=torch.uint8, device="cuda")
    bad = 0
    for rep in range(60):
        ctx = g.ProverContext(zk)
        for call in range(2):
            prf = ctx.prove(wit, g.Mask(r, s))
            ok = E.g2_from_array(prf.pi_b)[0] == wantB
            if not ok:
                bad += 1
                _lib.check(lib.g16_ctx_last_partials(ctx._h, parts.data_ptr()))
                pb = parts.cpu().numpy()[256:384].view(np.uint64)
                msm_ok = E.g2_from_array(pb)[0] == wantM
                print("neqs", neqs, "rep", rep, "call", call, "pi_b WRONG; msm_b2 partial ok =", msm_ok, ctx.last_stats, flush=True)
        ctx.close()
        # dirty the allocator a bit
        junk = torch.randint(0, 255, (rep * 1000003 % 7000000 + 10,), dtype=torch.uint8, device="cuda"); del junk
    print("neqs", neqs, "bad", bad, "of 120", flush=True)
