#!/usr/bin/env python
"""Summarise an Nsight Compute report (``ncu --set full``) as a small CSV + markdown table.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_em_full

Writes <out>.csv (one row per profiled launch, the metrics the roofline entry of bench.py and
DESIGN.md quote) and <out>.md.  Reads the report with ``ncu -i ... --page raw --csv``.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('Kernel Name', 'kernel'), ('Block Size', 'block'), ('Grid Size', 'grid'),
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
    ('smsp__inst_executed.sum', 'warp_inst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pipe_pct'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem_pct'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall_short_sb'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_sb'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall_wait'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall_math'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall_mio'),
    ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'stall_membar'),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), n, units[hdr.index(m)]) for m, n in METRICS if m in hdr]
    table = [[f'{n} [{u}]' if u else n for _, n, u in cols]]
    for r in rows[2:]:
        table.append([r[i] for i, _, _ in cols])
    with open(out + '.csv', 'w', newline='') as fh:
        csv.writer(fh).writerows(table)
    with open(out + '.md', 'w') as fh:
        fh.write(f'ncu --set full summary of `{rep.split("/")[-1]}` (one row per profiled launch)\n\n')
        fh.write('| ' + ' | '.join(table[0]) + ' |\n|' + '---|' * len(table[0]) + '\n')
        for r in table[1:]:
            r = list(r)
            r[0] = r[0].replace('|', '/')[:70]
            fh.write('| ' + ' | '.join(r) + ' |\n')
    print(f'{len(table) - 1} launches -> {out}.csv/.md')


if __name__ == '__main__':
    main()
