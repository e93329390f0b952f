"""Turn an ncu launch list (csv) and/or a full-set report (.ncu-rep) into the small summaries kept under profiles/.

    python scripts/summarize_ncu.py --launches gpurun_out/launches_r1.csv --out profiles/r1_launches.txt
    python scripts/summarize_ncu.py --report gpurun_out/prof_r1_step.ncu-rep --out profiles/r1_step_full.csv [--traffic profiles/traffic.json]
"""
import argparse, collections, csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "launch__grid_size", "launch__block_size"]
EPI = {"0": "EPI_LINEAR", "1": "EPI_GN_SILU", "2": "EPI_DDPM", "3": "EPI_MSE", "4": "EPI_RBF", "5": "EPI_GN_BWD", "6": "EPI_WGRAD"}


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES) from {path}\n")
        f.write(f"# total {tot/1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches\n")
        f.write("share_pct,launches,avg_us,kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{sum(v)/tot*100:.2f},{len(v)},{sum(v)/len(v)/1e3:.1f},\"{k[:160]}\"\n")
    print(open(out).read())


def report(path, out, traffic):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    tr = {}
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{k} [{units[idx[k]]}]" for k in keys])
        for r in data:
            w.writerow([r[idx["Kernel Name"]]] + [r[idx[k]] for k in keys])
            def val(k):
                v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tr.setdefault(r[idx["Kernel Name"]], []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    print(open(out).read())
    if traffic:
        names = {"gemm_tc_kernel<2,": "output_proj+reverse_update", "gemm_tc_kernel<0,": "input_proj+emb_add",
                 "ddpm_fused_kernel": "output_proj+reverse_update+next_input_proj"}
        try:
            outj = json.load(open(traffic))      # keep the entries of kernels this report does not contain
        except Exception:
            outj = {}
        for k, v in tr.items():
            for pat, nice in names.items():
                if pat in k:
                    outj[nice] = sum(v) / len(v)
        outj["_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch (100k patients, one row chunk) from {path}"
        json.dump(outj, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches"); ap.add_argument("--report"); ap.add_argument("--out", required=True); ap.add_argument("--traffic")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.out)
    if a.report:
        report(a.report, a.out, a.traffic)
