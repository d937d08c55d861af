"""Condense an .ncu-rep (ncu -i ... --page raw --csv) into the per-kernel lines DESIGN.md / bench.py quote.
    python scripts/ncu_summarize.py gpurun_out/r02_prof_net_a.ncu-rep > profiles/r02_net_a_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_i = hdr.index('Kernel Name')
print('# %s  (ncu --set full --clock-control none; one block per captured launch)' % rep)
for r in data:
    print('Kernel Name'.ljust(72), r[name_i][:150])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(k.ljust(72), r[i], units[i])
    rd, wr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')

    def to_bytes(v, u):
        f = float(v.replace(',', ''))
        return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    print('dram bytes read + written per launch'.ljust(72), '%.1f MB' % ((to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])) / 1e6))
    print()
