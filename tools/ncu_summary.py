"""Summarise an .ncu-rep (read here, no GPU): per kernel the metrics the roofline discussion needs."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print('   ', w, r[i], units[i])
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
            try:
                st.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1.0
    print('    warp-state samples (% of all):', ', '.join(f'{n} {v / tot * 100:.1f}' for v, n in sorted(st, reverse=True)[:9]))
