"""Summarise an .ncu-rep (first kernel): key raw metrics + top stall lines.  usage: ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active",
        "lts__t_sector_hit_rate.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if any(h.endswith(w) or h == w for w in want):
        print(f"{h:95s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ci = {h: i for i, h in enumerate(h2)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ci[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:18]:
    st = {k: f(r, k) for k in h2 if k.startswith("stall_") and "Not Issued" not in k and f(r, k) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{f(r, '# Samples'):7.0f} {r[ci['Source']][:70]:70s} {top}")
