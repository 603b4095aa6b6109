#!/usr/bin/env python
"""Turns the raw ncu artefacts a gpurun call leaves in gpurun_out/ into the tracked summaries
under profiles/ (run here, in the container; ncu reads the .ncu-rep without a GPU):

  python profiles/summarise.py TAG [gpurun_out/prof_dev.ncu-rep] [gpurun_out/launches_dev.csv]

  profiles/TAG_launches.csv          the `--metrics gpu__time_duration.sum` launch list, verbatim
  profiles/TAG_launch_shares.txt     mean duration per kernel and its share of one evaluation
  profiles/TAG_eval_fast_details.txt `ncu --page details` of the `--set full` capture
  profiles/TAG_eval_fast_hotspots.txt instruction mix / stall reasons per loop level (source page)
  profiles/traffic.json              dram bytes per launch of the dominant kernel (read by bench.py)
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [a for a in sys.argv if a != "--no-traffic"] + (["--no-traffic"] if "--no-traffic" in sys.argv else [])
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else os.path.join(ROOT, "gpurun_out", "prof_dev.ncu-rep")
launches = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else os.path.join(ROOT, "gpurun_out", "launches_dev.csv")
out = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


# ---- launch list -------------------------------------------------------------------------
if os.path.exists(launches):
    shutil.copy(launches, out("launches.csv"))
    rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    per_eval = {k: sum(v) / len(v) for k, v in agg.items()
                if not k.startswith("void at::") and "fp64_peak" not in k and "modes_kernel" not in k}
    tot = sum(per_eval.values())
    with open(out("launch_shares.txt"), "w") as f:
        f.write("# mean gpu__time_duration per launch (ns, ncu --clock-control none: cold-cache, serialised)\n")
        f.write("# and share of one fused evaluation (ytab + prep_cluster + eval_fast + eval + score_reduce + finalize)\n")
        for k, v in agg.items():
            m = sum(v) / len(v)
            share = f"{100 * m / tot:6.2f}%" if k in per_eval else "   (not part of an evaluation)"
            f.write(f"{m:12.0f} ns  n={len(v):3d}  {share}  {k}\n")
        f.write(f"{tot:12.0f} ns  one evaluation\n")

# ---- full capture --------------------------------------------------------------------------
open(out("eval_fast_details.txt"), "w").write(ncu("--page", "details"))
raw = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
h, r = raw[0], dict(zip(raw[0], raw[2]))
units = dict(zip(raw[0], raw[1]))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second"]
src = list(csv.reader(ncu("--page", "source", "--csv", "--print-source", "sass").splitlines()))
shdr = src[1]
data = [dict(zip(shdr, x)) for x in src[2:] if len(x) == len(shdr)]
tot_i = sum(int(d["Instructions Executed"]) for d in data)
tot_s = sum(int(d["# Samples"]) for d in data)


def opname(d):
    s = re.sub(r"^@!?U?P\w+\s+", "", d["Source"].strip())
    o = s.split()[0].rstrip(";")
    b = o.split(".")[0]
    return "IMAD.MOV" if (b == "IMAD" and ".MOV" in o) else b


with open(out("eval_fast_hotspots.txt"), "w") as f:
    f.write(f"# {src[0][1]}\n# report: {os.path.basename(rep)}\n")
    for k in keys:
        if k in r:
            f.write(f"{k:85s} {r[k]} {units.get(k, '')}\n")
    f.write(f"\n# instructions executed {tot_i}, stall samples {tot_s}\n")
    f.write("# loop level = executions of the SASS instruction over the whole launch; per level: number of\n"
            "# distinct instructions, share of executed instructions, share of stall samples, opcode mix\n")
    c, cs = collections.Counter(), collections.Counter()
    for d in data:
        n = int(d["Instructions Executed"])
        c[n] += 1
        cs[n] += int(d["# Samples"])
    for k, v in sorted(c.items(), key=lambda x: -x[0] * x[1])[:8]:
        ops = collections.Counter(opname(d) for d in data if int(d["Instructions Executed"]) == k)
        f.write(f"{k:9d} x {v:4d} instr  {100 * k * v / tot_i:5.1f}% of instructions  {100 * cs[k] / tot_s:5.1f}% of samples  "
                f"{dict(ops.most_common(14))}\n")
    stalls = [x for x in shdr if x.startswith("stall_") and "Not Issued" not in x]
    st = {x: sum(int(d[x]) for d in data) for x in stalls}
    tt = sum(st.values())
    f.write("\n# warp stall reasons (all samples)\n")
    for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
        f.write(f"{k:28s} {100 * v / tt:5.1f}%\n")


def to_bytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * m[unit]


m = re.search(r"--clusters (\d+)", open(os.path.join(ROOT, "scripts", "ncu_full_capture.sh")).read())
clusters = int(m.group(1)) if m else None
traffic = {"kernels": [{
    "kernel": src[0][1].replace("void ", "").replace("(int)", "").replace("(bool)", "").split("(")[0],
    "source": f"profiles/{tag}_eval_fast_details.txt (ncu --set full --clock-control none)",
    "clusters": clusters,
    "dram_bytes_per_launch": to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) +
    to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]),
    "duration_us_under_ncu": float(r["gpu__time_duration.sum"].replace(",", "")),
    "fp64_pipe_busy_pct": float(r["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
    "issue_slots_busy_pct": float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
    "shared_mem_pipe_pct": float(r["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
    "registers_per_thread": int(float(r["launch__registers_per_thread"])),
}]}
if "--no-traffic" not in sys.argv:          # traffic.json describes the headline kernel only
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(open(out("eval_fast_hotspots.txt")).read())
print(open(out("launch_shares.txt")).read() if os.path.exists(out("launch_shares.txt")) else "")
print(json.dumps(traffic))
