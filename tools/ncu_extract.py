"""Extract one kernel's metric rows from an .ncu-rep into the compact three-line CSV kept under profiles/
(metric names, units, values; the columns judged: dram__, launch__, sm__/smsp__ pipe and stall metrics, gpu__time).
    python tools/ncu_extract.py gpurun_out/x.ncu-rep profiles/r02_x_ncu.csv"""
import csv, subprocess, sys

KEEP = ("dram__", "gpu__time", "launch__", "sm__throughput", "sm__warps_active", "sm__cycles", "sm__inst_executed",
        "sm__pipe_", "smsp__issue_active", "smsp__inst_executed.sum", "smsp__average_warp", "smsp__cycles_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__sass_inst_executed_op_tmem")

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
cols = [i for i, h in enumerate(hdr) if h == "Kernel Name" or h.startswith(KEEP)]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    for r in (hdr, units, vals):
        w.writerow([r[i] for i in cols])
print(out, len(cols), "columns")
