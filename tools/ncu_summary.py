"""Summarise an .ncu-rep (read here, no GPU needed): headline counters + hottest SASS lines by stall samples.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"   {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("total samples", tot, "instructions", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
print("stalls:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:nlines]
for i in order:
    r = data[i]
    top = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    prev = data[i - 1][ix["Source"]][:60] if i > 0 else ""
    print(f"{r[ix['# Samples']]:>6s} x{r[ix['Instructions Executed']]:>7s}  {r[ix['Source']][:70]:70s} {top[6:]:10s} | prev: {prev}")
