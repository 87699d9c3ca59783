#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per launch, the counters DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
    ('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'L1 data-pipe wavefronts %peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'L1 wavefronts shared'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'shared LSU wavefronts %peak'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'tensor-core operand wavefronts %peak'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared bank conflicts'),
    ('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active %'),
    ('sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor operand fetch active %'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'global ld requests'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'global ld sectors'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2 read sectors from L1'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %peak'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__grid_size', 'grid'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__cycles_elapsed.avg', 'SM cycles elapsed'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio_throttle'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe'),
    ('smsp__average_warp_latency_per_inst_issued.ratio', 'warps per issue'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx['Kernel Name']] for r in rows[2:]]
    print('%-34s %-10s' % ('metric', 'unit'), *['%-22s' % n.split('<')[0][-22:] + n[n.find('<'):n.find('>') + 1][:12] for n in names], sep=' | ')
    for key, label in WANT:
        if key not in idx:
            continue
        vals = []
        for r in rows[2:]:
            v = r[idx[key]].replace(',', '')
            try:
                v = '%.4g' % float(v)
            except ValueError:
                pass
            vals.append('%-34s' % v)
        print('%-34s %-10s' % (label, units[idx[key]][:10]), *vals, sep=' | ')


if __name__ == '__main__':
    main(sys.argv[1])
