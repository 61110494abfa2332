"""profiles/r02_kernels.json from an ncu report: per kernel (first launch of each name) the metrics bench.py attaches to
its live per-stage timings.  Usage: python tools/ncu_kernels_json.py report.ncu-rep [more.ncu-rep ...] > out.json"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "time_ms_in_capture",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_of_32",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
}
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("unnamed>::", "").replace("void ", "").strip()
        if name in out:
            continue
        out[name] = {v: float(r[idx[k]]) for k, v in KEEP.items() if k in idx and r[idx[k]] not in ("", "n/a")}
        out[name]["capture"] = rep.split("/")[-1]
json.dump(out, sys.stdout, indent=1)
