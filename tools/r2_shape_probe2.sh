#!/bin/bash
# second round of shard shapes for the cost-model fit: fused pairs, two- and three-group ranks, offset ranges
mkdir -p gpurun_out
S() { python - "$@" <<'PY'
import sys
names = {"a1": 0, "b1": 1, "c1": 2, "b2": 3, "h": 4}
out = []
for spec in sys.argv[1:]:
    fr = [0.0] * 10
    for part in spec.split("+"):
        nm, rng = part.strip().split("[")
        lo, hi = rng.rstrip("]").split(",")
        for a in (["a1", "b1"] if nm == "ab" else ["a1", "b1", "c1"] if nm == "abc" else [nm]):
            fr[2 * names[a]] = float(lo); fr[2 * names[a] + 1] = float(hi)
    out.append("%s=%s" % (spec.replace(" ", ""), ",".join("%g" % v for v in fr)))
print("\n".join(out))
PY
}
mapfile -t SHAPES < <(S "ab[0,.25]" "ab[0,.5]" "ab[0,.75]" "ab[.19,1]" "ab[.5,1]" "ab[.75,1]" \
  "abc[0,.29]" "abc[0,.5]" "abc[.29,1]" "abc[0,1]+b2[0,.13]" \
  "a1[0,1]+b1[0,.25]" "b1[.25,1]+c1[0,.47]" "c1[.47,1]+b2[0,.15]" \
  "ab[.19,1]+c1[0,.35]" "c1[0,.35]" "ab[0,1]+c1[0,.25]" "ab[0,1]+c1[0,.5]" "ab[.5,1]+c1[0,.5]" "ab[.5,1]+c1[0,1]" \
  "b1[0,1]+c1[0,1]+b2[0,.13]" "b1[0,1]+c1[0,1]" \
  "h[0,1]+ab[0,.19]" "h[0,1]+ab[0,.4]" "h[0,1]+c1[0,.5]" "h[0,.5]+c1[0,.25]" "h[0,.5]+ab[0,.25]" \
  "c1[.35,1]+b2[0,.33]" "c1[0,1]+b2[0,.5]" "c1[.75,1]+b2[0,.75]" "ab[.5,1]+b2[0,.25]" "ab[0,1]+b2[0,.25]" \
  "b2[.33,1]" "b2[.5,1]" "b1[0,1]" "c1[.5,1]" "a1[.5,1]+b1[0,.5]" "a1[0,.5]+c1[.5,1]" "a1[0,1]+c1[0,1]+b2[0,1]" "abc[0,1]+b2[0,1]" "h[0,1]+abc[0,1]")
timeout 500 python tools/shape_probe.py 20 "${SHAPES[@]}" 2>&1 | grep "^shape" | tee gpurun_out/r2_shape_probe2.log
