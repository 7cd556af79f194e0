"""Summarise an `ncu --set full` report of the three MLP kernels into profiles/<name>_summary.csv:
one row per metric of interest, one column per kernel (first captured launch of each).
Usage: python tools/ncu_summary.py gpurun_out/prof_r02_mlp.ncu-rep profiles/prof_r02_mlp_summary.csv"""
import csv, io, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
WANT = re.compile(r"^(dram__bytes_(read|write)\.sum(\.pct.*|\.per_second)?|gpu__time_duration\.sum|"
                  r"sm__pipe_tensor.*cycles_active.*|sm__inst_executed_pipe_(xu|fma|alu|lsu|uniform)\S*pct\S*|"
                  r"sm__throughput\.avg\.pct.*|sm__issue_active\.avg\.pct.*|sm__inst_issued\.avg\.pct.*|sm__cycles_active\.avg|"
                  r"smsp__inst_executed\.sum|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|cluster.*)|"
                  r"lts__t_bytes\.sum.*|lts__throughput.*|l1tex__throughput.*|lts__t_sector_hit_rate\.pct|"
                  r"l1tex__data_bank_conflicts_pipe_lsu.*sum|smsp__pcsamp_warps_issue_stalled_\w+|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__clock.*|dram__throughput.*|"
                  r"smsp__cycles_active\.avg|sm__sass_inst_executed_op_local.*)$")
kernels = {}
for r in data:
    short = re.sub(r".*::", "", r[name_col].split("(")[0]).split("<")[0]
    kernels.setdefault(short, r)
cols = [k for k in ("mlp_fwd_kernel", "mlp_bwd_kernel", "wgrad_kernel") if k in kernels] or sorted(kernels)
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + cols)
    for i, m in enumerate(hdr):
        if WANT.match(m):
            w.writerow([m, units[i]] + [kernels[k][i] for k in cols])
print("wrote", out, "kernels:", cols)
