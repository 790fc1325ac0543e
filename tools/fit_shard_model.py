"""Least-squares fit of the shard cost model (prover.cu: rank_cost) to a tools/shape_probe.py log.
    python tools/fit_shard_model.py gpurun_out/r2_shape_probe.log [LOG_N]"""
import sys, re
import numpy as np
from scipy.optimize import lsq_linear

def num_windows(c): return (255 + c - 1) // c
def pick_window(n):
    best, bc = 1e300, 4
    for c in range(4, 23):
        W = num_windows(c); nb = 1 << (c - 1)
        cost = 10.0 * W * n + 28.0 * nb + 0.6 * W * n + 4000.0
        if W * n >= 2147483648.0: continue
        if cost < best: best, bc = cost, c
    return bc

NAMES = ["g1_pair", "g2_pair", "sort_pair", "g1_bucket", "g2_bucket", "rank_fixed", "extra_group", "g1_and_g2", "quotient", "early", "g2_fixed"]

def features(fr, n, nvars):
    a1, b1, c1, b2, h = [(fr[2 * i], fr[2 * i + 1]) for i in range(5)]
    groups = {}
    for nm, r, g2 in (("a1", a1, 0), ("b1", b1, 0), ("c1", c1, 0), ("b2", b2, 1)):
        if r[1] > r[0]:
            groups.setdefault((r, nvars), []).append(g2)
    if h[1] > h[0]:
        groups.setdefault((h, n, "h"), []).append(0)
    x = dict.fromkeys(NAMES, 0.0)
    ng1 = 0
    for key, sets in groups.items():
        r, N = key[0], key[1]
        npts = int(round((r[1] - r[0]) * N))
        c = pick_window(npts); W = num_windows(c); nb = 1 << (c - 1)
        pairs = npts * W
        x["sort_pair"] += pairs
        for g2 in sets:
            x["g2_pair" if g2 else "g1_pair"] += pairs
            x["g2_bucket" if g2 else "g1_bucket"] += nb
        if 0 in sets: ng1 += 1
    has_g2 = b2[1] > b2[0]
    x["rank_fixed"] = 1.0
    x["extra_group"] = max(0, len(groups) - 1)
    x["g1_and_g2"] = 1.0 if (has_g2 and ng1) else 0.0
    x["quotient"] = 1.0 if h[1] > h[0] else 0.0
    x["early"] = 1.0 if (a1[1] > a1[0] or b1[1] > b1[0]) else 0.0
    x["g2_fixed"] = 1.0 if has_g2 else 0.0
    return [x[k] for k in NAMES]

def main():
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n = nvars = 1 << log_n
    rows, y, names = [], [], []
    for line in open(sys.argv[1]):
        m = re.match(r"shape (.+?)\s+([-0-9.,e]+)\s+([0-9.]+) ms", line)
        if not m: continue
        fr = [float(v) for v in m.group(2).split(",")]
        rows.append(features(fr, n, nvars)); y.append(float(m.group(3))); names.append(m.group(1).strip())
    A = np.array(rows); y = np.array(y)
    scale = np.maximum(A.max(axis=0), 1e-30)
    lo = np.zeros(len(NAMES)); hi = np.full(len(NAMES), np.inf)
    lo[NAMES.index("extra_group")] = -np.inf
    res = lsq_linear(A / scale, y, bounds=(lo, hi))
    coef = res.x / scale
    UNIT = 6.76e7      # modmuls per ms at the measured peak
    for k, v in zip(NAMES, coef):
        print("%-12s %12.4g ms   = %10.4g units" % (k, v, v * UNIT))
    pred = A @ coef
    for nm, a, b in zip(names, y, pred):
        print("%-32s meas %6.3f  model %6.3f  %+5.1f%%" % (nm, a, b, (b - a) / a * 100))
    print("rms %.3f ms" % np.sqrt(np.mean((pred - y) ** 2)))

if __name__ == "__main__":
    main()
