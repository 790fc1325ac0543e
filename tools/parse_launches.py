"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_run.py.
Usage: python tools/parse_launches.py gpurun_out/launches.csv"""
import collections, csv, re, sys
txt = open(sys.argv[1]).read()
rd = csv.DictReader(txt[txt.index('"ID"'):].splitlines())
rows = []
for r in rd:
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
        v = v / 1e6 if u.startswith('n') else v / 1e3 if u.startswith('u') else v
        rows.append((int(r['ID']), r['Kernel Name'], v, r.get('Grid Size'), r.get('Block Size')))
short = lambda n: re.sub(r'\(.*', '', n).replace('void ', '').replace('g16::', '')[:64]
idx = [i for i, r in enumerate(rows) if 'k_build_abc' in r[1]]
def seg_report(title, seg):
    tot = sum(r[2] for r in seg)
    agg = collections.OrderedDict()
    for r in seg:
        k = short(r[1]); agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += r[2]
    print("== %s: %d launches, serialized kernel time %.3f ms" % (title, len(seg), tot))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("  %8.3f ms %5.1f%% x%-3d %s" % (t, 100 * t / tot, c, k))
i0 = idx[-1]
j = i0
while 'k_assemble_final' not in rows[j][1]: j += 1
# include the mask-term kernel launched just before the witness copy
k0 = i0
while k0 > 0 and 'k_mask_terms' not in rows[k0][1]: k0 -= 1
seg_report("second full proof", rows[k0:j + 1])
rest = rows[j + 1:]
# standalone sections are separated by k_msm_digits launches
starts = [i for i, r in enumerate(rest) if 'k_msm_digits' in r[1]] + [len(rest)]
for a, b in zip(starts[:-1], starts[1:]):
    seg = [r for r in rest[a:b] if 'at::' not in r[1]]
    kind = 'G2' if any('Fp2' in r[1] for r in seg) else 'G1'
    lay = 'table' if any('k_final_sum' in r[1] for r in seg) else 'plain'
    if any('k_build_table' in r[1] for r in seg):
        seg = [r for r in seg if 'k_build_table' not in r[1] and 'k_ntt' not in r[1]]
    seg = [r for r in seg if 'k_ntt' not in r[1]]
    seg_report("standalone MSM %s %s" % (kind, lay), seg)
ntt = [r for r in rest if 'k_ntt_pass' in r[1]]
print("== ntt passes:", ["%.3f" % r[2] for r in ntt])
bt = [r for r in rows if 'k_build_table' in r[1]]
print("== table builds:", ["%s %.1f ms" % ('G2' if 'Fp2' in r[1] else 'G1', r[2]) for r in bt])
