"""Prints the metrics the profiles/ summaries quote from an .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'gpc__cycles_elapsed.max.per_second',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('==', path)
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'| `{w}` | {vals[i][:100]} {units[i]} |')
