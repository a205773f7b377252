"""Summarise an exported ncu report: python tools/ncu_summary.py raw.csv source.csv"""
import csv, sys
raw, srcf = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
h, u, v = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active']
print("| metric | unit | value |\n|---|---|---|")
for k in keys:
    if k in h:
        i = h.index(k); print(f"| {k} | {u[i]} | {v[i]} |")
rows = list(csv.reader(open(srcf)))
h, data = rows[1], rows[2:]
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
si, src, ii = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
recs = [r for r in data if r[si].isdigit()]
tot = sum(int(r[si]) for r in recs)
print(f"\nPC-sample segments (split at BAR/EXIT), total samples {tot}:\n")
start = 0
for i, r in enumerate(recs):
    if 'BAR.' in r[src] or 'EXIT' in r[src] or i == len(recs) - 1:
        seg = recs[start:i + 1]; n = sum(int(x[si]) for x in seg); ins = sum(int(x[ii]) for x in seg)
        if n > tot * 0.01:
            st = {h[c]: sum(int(x[c]) for x in seg if x[c].isdigit()) for c in stall_cols}
            top = {k[6:]: round(100 * val / n, 1) for k, val in sorted(st.items(), key=lambda x: -x[1]) if val > n * 0.015}
            print(f"* SASS {start}-{i}: {100*n/tot:.1f}% of samples, {ins/1e6:.0f}M warp-insts, stalls % {top}")
        start = i + 1
