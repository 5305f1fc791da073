"""Key metrics + top stall sites of the first kernel in an .ncu-rep (development helper). usage: ncu_kernel.py rep [ntop]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for k in range(2, len(rows)):
    print("==", rows[k][hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in keys or ('stalled' in h and 'per_issue_active' in h and 'not_issued' not in h and float(rows[k][i] or 0) > 0.05):
            print("  %-95s %s %s" % (h, rows[k][i], rows[1][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
base = int(data[0][ia], 16)
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[isamp]))[:ntop]:
    top = sorted(((int(r[i] or 0), h) for i, h in stalls), reverse=True)[:2]
    print("%5x %6s (%4.1f%%) ex=%8s %-34s %s" % (int(r[ia], 16) - base, r[isamp], 100 * int(r[isamp]) / tot, r[iex], " ".join("%s=%d" % (h[6:], v) for v, h in top if v), r[isrc].strip()[:60]))
