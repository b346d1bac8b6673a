"""Condense an .ncu-rep (one `ncu --set full` capture) into the handful of metrics DESIGN.md and
bench.py quote: python tools/ncu_summary.py gpurun_out/prof_ws_n16384.ncu-rep profiles/NAME.json"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
want = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_uniform.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
res = {}
for i, h in enumerate(hdr):
    if h in want:
        res[h] = {"value": val[i], "unit": units[i]}
to_bytes = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
try:
    rd, wr = res["dram__bytes_read.sum"], res["dram__bytes_write.sum"]
    res["dram_bytes_per_launch"] = float(rd["value"]) * to_bytes[rd["unit"]] + float(wr["value"]) * to_bytes[wr["unit"]]
except Exception as e:  # noqa: BLE001
    res["dram_bytes_per_launch"] = None
    res["error"] = repr(e)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
