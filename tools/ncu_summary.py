"""Extract the judged subset of an ncu --set full report into a small text file.
usage: python tools/ncu_summary.py report.ncu-rep out.txt "header comment" """
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'memory_l1_wavefronts_shared', 'memory_l1_wavefronts_shared_ideal',
        'sass__inst_executed_shared_loads', 'smsp__inst_executed.sum', 'sm__inst_issued.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']
STALL = 'smsp__average_warps_issue_stalled_'


def main():
    rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# %s\n# source: %s (ncu --set full --clock-control none); raw-page extract by tools/ncu_summary.py\n' % (comment, rep))
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write('## kernel: %s\n' % name)
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or (h.startswith(STALL) and h.endswith('_per_issue_active.ratio')) or 'tensor' in h and 'pct_of_peak_sustained_active' in h and h.endswith('.avg.pct_of_peak_sustained_active'):
                    f.write('%s [%s] = %s\n' % (h, u, v))
    print(open(out).read())


if __name__ == '__main__':
    main()
