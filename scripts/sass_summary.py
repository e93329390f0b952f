"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP / UTMAPF (TMA), HMMA (legacy mma.sync -- must be absent), plus registers from the ELF.
    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "osteosarcoma_diffusionmodel_b200/libosteo_ddpm.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTMAPF", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "MUFU", "FFMA"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {so}: SASS instruction counts per kernel (static), sm_100a")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG (tensor loads) / UBLKCP (bulk copy global -> shared) / UBLKPF (bulk L2 prefetch), cp.async -> LDGSTS; HMMA (mma.sync) must be 0")
print("kernel," + ",".join(keys) + ",total_sass")
for (mangled, c), name in zip(counts.items(), demangle):
    if c["_total"] == 0:
        continue
    short = re.sub(r"\(.*\)$", "", name.replace("osteo::", "").replace("void ", ""))
    print(f"\"{short[:110]}\"," + ",".join(str(c[k]) for k in keys) + f",{c['_total']}")
