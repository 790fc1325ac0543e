"""Busy time of single shard SHAPES on one GPU (two proofs in flight, no exchange): the data the cost model of
`shard_plan` (prover.cu) is fitted to.  A shape is "name=a1lo,a1hi,b1lo,b1hi,c1lo,c1hi,b2lo,b2hi,hlo,hhi" (fractions),
installed through the experiment knob G16_SHARD_SHAPE.
    python tools/shape_probe.py LOG [shape ...]      (no shapes: the built-in list)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nim-groth16_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import g16b200 as g
from g16b200 import _lib
import bench

def shape(a1=(0, 0), b1=(0, 0), c1=(0, 0), b2=(0, 0), h=(0, 0)):
    return ",".join("%g,%g" % t for t in (a1, b1, c1, b2, h))

BUILTIN = []
for f in (0.125, 0.25, 0.5, 0.75, 1.0):
    BUILTIN.append(("c1 %.3f" % f, shape(c1=(0, f))))
for f in (0.25, 0.5, 1.0):
    BUILTIN.append(("a1 %.3f" % f, shape(a1=(0, f))))
for f in (0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 1.0):
    BUILTIN.append(("b2 %.3f" % f, shape(b2=(0, f))))
for f in (0.333, 0.5, 1.0):
    BUILTIN.append(("h %.3f" % f, shape(h=(0, f))))
BUILTIN += [
    ("a1 1 + b1 1 (one sort)", shape(a1=(0, 1), b1=(0, 1))),
    ("a1 1 + b1 1 + c1 1 (one sort)", shape(a1=(0, 1), b1=(0, 1), c1=(0, 1))),
    ("a1 .66 + b1 1 + c1 .18", shape(a1=(0.34, 1), b1=(0, 1), c1=(0, 0.18))),
    ("a1 .5 + b1 1", shape(a1=(0.5, 1), b1=(0, 1))),
    ("b1 1 + c1 .5", shape(b1=(0, 1), c1=(0, 0.5))),
    ("c1 .82 + b2 .36", shape(c1=(0.18, 1), b2=(0, 0.36))),
    ("c1 1 + b2 .25", shape(c1=(0, 1), b2=(0, 0.25))),
    ("c1 .5 + b2 .5", shape(c1=(0.5, 1), b2=(0, 0.5))),
    ("h 1 + a1 .34", shape(h=(0, 1), a1=(0, 0.34))),
    ("h .5 + a1 .25", shape(h=(0, 0.5), a1=(0, 0.25))),
    ("h 1 + a1 1 + b1 1 + c1 .18", shape(h=(0, 1), a1=(0, 1), b1=(0, 1), c1=(0, 0.18))),
    ("c1 .82 + b2 1", shape(c1=(0.18, 1), b2=(0, 1))),
]

def main():
    log_n = int(sys.argv[1])
    shapes = [a.split("=", 1) for a in sys.argv[2:]] or BUILTIN
    lib = _lib.load()
    zk, wit, _ = bench.make_fixture(g, log_n)
    w = torch.from_numpy(np.ascontiguousarray(wit).view(np.int64).copy()).to("cuda")
    mask = g.Mask(bench.MASK_R, bench.MASK_S)
    steps, d = 16, 2
    for name, sh in shapes:
        os.environ["G16_SHARD_SHAPE"] = sh
        base = g.ProverContext(zk, 1, 2, trusted=True)
        slots = [base, base.clone()]
        parts = [torch.zeros(400, dtype=torch.uint8, device="cuda") for _ in slots]
        def run(n):
            for i in range(n):
                c = slots[i % d]
                if i >= d:
                    _lib.check(lib.g16_prove_partials_wait(c._h, None))
                c.set_mask(mask)
                _lib.check(lib.g16_prove_partials_submit(c._h, w.data_ptr(), 1, 1, parts[i % d].data_ptr()))
            for i in range(min(d, n)):
                _lib.check(lib.g16_prove_partials_wait(slots[(n - min(d, n) + i) % d]._h, None))
        run(4)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(steps)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / steps * 1e3)
        print("shape %-32s %-44s %.3f ms" % (name, sh, best), flush=True)
        for c in slots:
            c.close()

if __name__ == "__main__":
    main()
