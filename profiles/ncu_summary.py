"""Summarise an .ncu-rep (run where ncu is installed): key metrics of the first kernel + opcode histogram."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-72s %s [%s]" % (w, " | ".join(r[i] for r in rows[2:4]), units[i]))
stall = [(h, rows[2][i]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
for h, v in sorted(stall, key=lambda kv: -float(kv[1] or 0))[:8]:
    print("  stall %-64s %s" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__warp_issue_stalled_", ""), v))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]
ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
first = []
for r in rows[2:]:
    if len(r) < 10:
        break
    first.append(r)
tot = sum(int(r[ie]) for r in first)
print("warp-instructions: %d total, %.1f per env-step" % (tot, tot / n_env))
c, s = Counter(), Counter()
for r in first:
    parts = r[src].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    c[op] += int(r[ie]); s[op] += int(r[smp])
for op, v in c.most_common(70):
    print("  %-22s %8.1f /env-step   samples %d" % (op, v / n_env, s[op]))

# per source line (correlated view)
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(cs.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0]:
        lines.append(r)
ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print("top source lines by warp-instructions (per env-step):")
def _n(r):
    try:
        return int(r[ie] or 0)
    except ValueError:
        return 0


for r in sorted(lines, key=lambda r: -_n(r))[:45]:
    print("  L%-4s %8.1f  smp %-4s %s" % (r[0], _n(r) / n_env, r[smp], r[1].strip()[:110]))

# per function (by source line ranges found from markers in the .cu file)
import re
src_path = hdr and rows[0][1] if rows and rows[0] and rows[0][0] == "File Path" else None
marks = []
try:
    text = open(src_path).read().splitlines()
    for i, l in enumerate(text, 1):
        m = re.match(r"(?:template <[^>]*>\s*)?__device__ __forceinline__ \w[\w<>]* (\w+)\(|__global__ void .*? (\w+)\(", l)
        if m:
            marks.append((i, m.group(1) or m.group(2)))
except Exception:
    marks = []
if marks:
    marks.append((10 ** 9, "end"))
    agg = Counter()
    for r in lines:
        try:
            ln = int(r[0])
        except ValueError:
            continue
        name = "pre"
        for (a, nm), (b, _) in zip(marks, marks[1:]):
            if a <= ln < b:
                name = nm
        agg[name] += _n(r)
    print("warp-instructions per env-step by function:")
    for k, v in agg.most_common():
        print("  %-18s %8.1f" % (k, v / n_env))

print("top source lines by stall samples:")
tot_s = sum(int(r[smp] or 0) for r in lines if (r[smp] or "0").isdigit())
for r in sorted(lines, key=lambda r: -(int(r[smp]) if (r[smp] or "0").isdigit() else 0))[:25]:
    print("  L%-4s smp %-5s (%4.1f%%) instr/env %7.1f  %s" % (r[0], r[smp], 100.0 * int(r[smp]) / max(tot_s, 1), _n(r) / n_env, r[1].strip()[:100]))
if marks:
    aggs = Counter()
    for r in lines:
        try:
            ln = int(r[0]); sm_ = int(r[smp] or 0)
        except ValueError:
            continue
        name = "pre"
        for (a, nm), (b, _) in zip(marks, marks[1:]):
            if a <= ln < b:
                name = nm
        aggs[name] += sm_
    print("stall samples by function (share of warp residency time):")
    for k, v in aggs.most_common():
        if v:
            print("  %-18s %6d  %5.1f%%" % (k, v, 100.0 * v / max(tot_s, 1)))
