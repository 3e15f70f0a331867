"""Summarise an ncu report (--set full) into a small text file for profiles/.
usage: python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.txt [--traffic]
--traffic: also write profiles/scan_traffic.json = DRAM bytes (read + write) of one kernel R launch + one kernel W
launch (one batch's scan), stamped with the digest of the kernel sources (bench.py refuses a stale stamp)."""
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

if "--traffic" in sys.argv:
    import json, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = "W" if "tc_scan_wide" in d.get("Kernel Name", "") else ("R" if "tc_scan_kernel" in d.get("Kernel Name", "") else None)
        if name is None or name in per:
            continue
        per[name] = sum(float(d[k].replace(",", "")) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out_j = {"dram_bytes_per_launch": sum(per.values()), "per_kernel": per, "source_digest": bench.scan_source_digest(),
             "from": os.path.basename(rep), "note": "one kernel R launch + one kernel W launch = the scan of one 1024-query batch"}
    with open(os.path.join(os.path.dirname(os.path.abspath(out)), "scan_traffic.json"), "w") as fh:
        json.dump(out_j, fh, indent=1)
    print(out_j)
