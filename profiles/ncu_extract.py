"""Summarise an .ncu-rep (read here, no GPU needed): python profiles/ncu_extract.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    for r in rows[2:]:
        print('----', r[idx['Kernel Name']][:70])
        for w in WANT:
            if w in idx:
                print(f'  {w} = {r[idx[w]]} {units[idx[w]]}')
        vals = []
        for h in stall:
            try:
                vals.append((float(r[idx[h]].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
        print('  stalls:', ', '.join(f'{n}={v:.2f}' for v, n in sorted(vals, reverse=True)[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
