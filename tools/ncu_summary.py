"""Print the metrics that matter for a bandwidth-bound kernel from an `ncu --page raw --csv` export."""
import csv
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum',
    'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors.sum', 'l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum',
    'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.max.pct_of_peak_sustained_elapsed',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
    'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print(f"== {name[:100]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {vals[i]:>16s} {units[i]}")


if __name__ == '__main__':
    main(sys.argv[1])
