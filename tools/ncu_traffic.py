"""profiles/r2_traffic.json from .ncu-rep captures of k_pass (read here, no GPU needed):
    python tools/ncu_traffic.py C2x1=gpurun_out/prof_C2.ncu-rep C2x10=gpurun_out/prof_x10.ncu-rep
bench.py copies dram read+write of the matching workload into roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, units))

    def to_bytes(name):
        v, unit = float(d[name]), u[name].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]

    def to_us(name):
        v, unit = float(d[name]), u[name].lower()
        return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)

    out[key] = {"kernel": d.get("Kernel Name", "?")[:80], "grid": d.get("Grid Size"), "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
                "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "gpu_time_us_under_ncu": to_us("gpu__time_duration.sum"),
                "sm_throughput_pct": float(d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", "nan")),
                "issue_active_pct": float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan")),
                "alu_pipe_pct": float(d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "nan")),
                "lsu_pipe_pct": float(d.get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "nan")),
                "dram_throughput_pct": float(d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "nan")),
                "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '.txt')} (ncu --set full --clock-control none, one launch)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
