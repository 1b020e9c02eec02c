#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep and the launch-list CSV into the committed summaries under profiles/ (run in the build
container: ncu can read reports without a GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__shared_mem_per_block_dynamic"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarise(rep, md_path, title, round_tag):
    hdr, units, rows = raw_rows(rep)
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    lines = [f"# {title}", "", f"Source: `{os.path.basename(rep)}` (ncu --set full --clock-control none, {round_tag}); one block per captured launch.",
             "ncu replays each kernel ~40x with cold caches: durations here are NOT bench numbers, traffic and ratios are.", ""]
    recs = []
    for r in rows:
        lines.append(f"## {r[ik]}  grid {r[ig]} block {r[ib]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        rec = {"kernel": r[ik], "grid": r[ig]}
        for w, i in cols:
            lines.append(f"| {w} | {r[i]} | {units[i]} |")
            rec[w] = (r[i], units[i])
        lines.append("")
        recs.append(rec)
    open(md_path, "w").write("\n".join(lines))
    return recs


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def launches(csv_path, md_path, round_tag):
    lines_in = [l for l in open(csv_path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines_in))))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = [f"# Launch list of `bench.py --steps 20 --warmup 3 --no-cpu-baseline` under ncu ({round_tag})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and serialised: compare SHARES.",
           "The run contains: the timed region and its warm-up (`k_step_multi<float, L1, LINEAR>`: ONE launch = all K iterations, so 2 launches),",
           "the kernel-only reference of the roofline block (24 back-to-back `k_step` launches, no exchange), the e2e solve and its warm-up",
           "(`k_step_multi<float, L1, SQDIST>`: SquaredDistance gradient fused into the step, + one `k_ew<OP_SUB>` for the final f value),",
           "the counter-based fills of the synthetic inputs (`k_fill_counter`) and the stand-alone exchanges of the final scalar reads (`k_xchg`).",
           "", "| launches | total us | share | kernel |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | `{k[:110]}` |")
    open(md_path, "w").write("\n".join(out) + "\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    g = os.path.join(ROOT, "gpurun_out")
    p = os.path.join(ROOT, "profiles")
    os.makedirs(p, exist_ok=True)
    traffic = {}
    rep = os.path.join(g, f"{tag}_prof_step_multi.ncu-rep")
    if os.path.exists(rep):
        # the headline kernel since round 2: ONE persistent launch covers the K iterations of `bench.py --steps K`
        K = int(os.environ.get("NCU_STEPS", "20"))
        recs = summarise(rep, os.path.join(p, f"{tag}_ncu_k_step_multi.md"),
                         f"ncu: persistent multi-iteration FISTA step kernel k_step_multi at n = 1e8 fp32 (`bench.py --steps {K} --warmup 3`: one launch = {K} iterations)", tag)
        r = recs[-1]
        rd = to_bytes(*r["dram__bytes_read.sum"]) / K
        wr = to_bytes(*r["dram__bytes_write.sum"]) / K
        traffic["k_step_ffb_l1_f32"] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch_at_n": rd + wr, "n": 100000000,
                                         "algorithmic_bytes": 2e9, "iterations_in_captured_launch": K, "kernel": r["kernel"],
                                         "note": "per ITERATION: the captured launch runs K iterations, its dram__bytes are divided by K",
                                         "source": f"profiles/{tag}_ncu_k_step_multi.md"}
    for name, title in ((f"{tag}_prof_persist_small", "ncu: persistent on-device solver k_persist_solve, lasso_small ForwardBackward (tools/one_persist.py small fb), one CTA"),
                        (f"{tag}_prof_lsq", "ncu: least-squares kernels (tools/tune_lsq.py)")):
        rep = os.path.join(g, name + ".ncu-rep")
        if os.path.exists(rep):
            summarise(rep, os.path.join(p, name.replace("_prof_", "_ncu_") + ".md"), title, tag)
    if os.path.exists(os.path.join(g, f"{tag}_launches.csv")):
        launches(os.path.join(g, f"{tag}_launches.csv"), os.path.join(p, f"{tag}_launches.md"), tag)
    if traffic:
        json.dump(traffic, open(os.path.join(p, "traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(p)))


if __name__ == "__main__":
    main()
