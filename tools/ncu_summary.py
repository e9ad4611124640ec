"""Summarise an ncu report's raw+source CSV pages: key metrics and dynamic instruction mix per cell-update."""
import csv, collections, re, sys
raw, src, cells = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum [', 'dram__bytes_write.sum [', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread [', 'launch__occupancy_limit', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'smsp__inst_executed.sum [', 'launch__waves', 'sm__cycles_elapsed.avg [', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum.per_second', 'sm__cycles_elapsed.avg.per_second']
for h, u, v in zip(hdr, units, vals):
    key = f"{h} [{u}]"
    if any(w in key for w in want):
        try:
            if 'stalled' in h and float(v) < 0.05: continue
        except ValueError: pass
        print(f"{key} = {v}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
c, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    n = int(r[iE] or 0)
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', r[iS].strip())
    full = (m.group(1) + m.group(2)) if m else r[iS]
    c[full] += n; tot += n; samp[full] += int(r[iSamp] or 0)
wc = cells / 32
print(f"total warp-instructions {tot}  = {tot / wc:.1f} per cell-update")
for k, v in c.most_common(32):
    print(f"  {k:26s} {v / wc:7.2f}/cell   stall-samples {samp[k]}")
