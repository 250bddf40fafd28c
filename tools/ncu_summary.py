"""Print the interesting metrics of an .ncu-rep (raw page) per kernel.  usage: python tools/ncu_summary.py rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(
    r"gpu__time_duration.sum|sm__throughput.avg.pct|pipe_tensor.*(pct|sum)|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|"
    r"registers_per_thread|smsp__inst_executed.sum$|issue_active.avg.pct|warps_active.avg.pct|lts__t_sector_hit_rate.pct|"
    r"sm__cycles_elapsed.max|l1tex__data_bank_conflicts_pipe_lsu.sum$|sm__inst_executed_pipe_(xu|lsu|alu|fma|uniform).sum$|"
    r"smsp__average_warps?_issue_stalled.*_per_issue_active|launch__occupancy_limit|shared_mem_per_block|lts__t_bytes.sum$|l1tex__t_bytes.sum$")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print(f"   {h:95s} {v} {u}")
