"""Summarise an .ncu-rep (run here, no GPU): key raw metrics + top stalled SASS lines.
Usage: python tools/ncu_summary.py <rep> [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "derived__lts__lts2xbar_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]
for vals in rows[2:]:
    print("---- launch")
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print("%-75s %-10s %s" % (k, u[k], d[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
h = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if h is not None:
            break
        h = r
        continue
    if h is not None and len(r) == len(h):
        data.append(r)
if h:
    idx = {x: i for i, x in enumerate(h)}
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print("---- source: %d SASS lines, %d samples" % (len(data), tot))
    agg = {x: sum(int(r[idx[x]]) for r in data) for x in stalls}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:ntop]
    for i in sorted(top):
        r = data[i]
        s = {x[6:]: int(r[idx[x]]) for x in stalls if int(r[idx[x]]) > 0.1 * int(r[idx["# Samples"]])}
        print("%5d %7s %10s  %-70s %s" % (i, r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:70], s))
