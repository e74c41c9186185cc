"""Key metrics of every kernel in an .ncu-rep (raw page) as a compact table.
usage: python tools/ncu_summary.py rep.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.max", "cycles"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_insts")]
for r in rows[2:]:
    print("-" * 100)
    for key, label in want:
        if key in ci:
            v = r[ci[key]]
            if key == "Kernel Name":
                v = v[:140]
            print(f"{label:24s} {v} {units[ci[key]]}")
