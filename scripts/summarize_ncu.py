"""Summarise an ncu report (--set full) into a small text file for profiles/.
usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.txt"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum"]
with open(out, "w") as fh:
    fh.write(f"# ncu --set full --clock-control none summary of {rep}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        fh.write(f"\nkernel: {d.get('Kernel Name', '')}\n")
        for k in KEYS:
            if k in d:
                fh.write(f"  {k} = {d[k]} {u.get(k, '')}\n")
        for k in hdr:
            if "warp_issue_stalled" in k and k.endswith("_per_warp_active.pct") and d.get(k):
                try:
                    if float(d[k]) >= 2.0:
                        fh.write(f"  {k} = {d[k]}\n")
                except ValueError:
                    pass
print(open(out).read())
