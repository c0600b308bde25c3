#!/usr/bin/env python
"""summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json]"""
import csv, io, json, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for i, h in enumerate(hdr):
            base = h.split('.', 2)[-1] if h.count('.') >= 2 and h.split('.')[1] in ('TriageCompute',) else h
            for k in KEYS:
                if h == k or base == k:
                    d[k] = (r[i], units[i])
        # stall breakdown
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls[h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(r[i])
                except ValueError:
                    pass
        d['top_stalls'] = sorted(stalls.items(), key=lambda kv: -kv[1])[:6]
        out.append(d)
    for d in out:
        print('==', d['kernel'])
        for k, v in d.items():
            if k not in ('kernel',):
                print('   ', k, v)
    if '--json' in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)


if __name__ == '__main__':
    main()
