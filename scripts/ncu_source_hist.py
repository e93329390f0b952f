"""Opcode / stall histogram of one kernel from `ncu -i rep --page source --csv` (SASS view).
    python scripts/ncu_source_hist.py gpurun_out/fused_src.csv [--top 40] [--lines 60]"""
import csv, collections, re, sys

path = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break                       # next launch of the same kernel: keep the first section only
    if len(r) >= len(hdr) - 2:
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(float(r[ix[k]] or 0))
tot_inst = sum(I(r, "Instructions Executed") for r in data)
tot_samp = sum(I(r, "# Samples") for r in data)
print("total warp-inst", tot_inst, "samples", tot_samp, "sass lines", len(data))
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += I(r, "Instructions Executed")
    ops[o] += I(r, "# Samples")
for o, c in op.most_common(top):
    print(f"{o:12s} inst {c / tot_inst * 100:6.2f}%  samples {ops[o] / tot_samp * 100:6.2f}%")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(I(r, h) for r in data) for h in stalls}
print({k: round(v / tot_samp * 100, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]})
if nlines:
    print("--- hottest SASS lines by samples")
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:nlines]:
        st = sorted(((I(r, h), h) for h in stalls), reverse=True)[:2]
        print(f"{I(r,'# Samples'):7d} {I(r,'Instructions Executed'):10d}  {r[ix['Source']].strip()[:90]:90s} {st}")
