"""Summarise ncu outputs brought back from the GPU box into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/rN_launches_summary.txt
    python tools/ncu_summary.py full     gpurun_out/x_full.ncu-rep  > profiles/rN_x_full.txt

`launches` aggregates the `--metrics gpu__time_duration.sum` launch list per kernel (cold-cache, serialised: compare
SHARES, not absolutes).  `full` prints the roofline-relevant raw metrics of every profiled launch of a `--set full`
report (read here with `ncu -i ... --page raw --csv`)."""
import collections
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex.sum', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
]


def short(name):
    return name.split('(')[0].replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')[:60]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r['Kernel Name'])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value']) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {len(rows)} launches, {tot:.3f} ms summed gpu__time_duration (ncu, cold cache, serialised)')
    print(f'{"kernel":60s} {"n":>5s} {"ms":>10s} {"share":>7s}')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:60s} {v[0]:5d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%')


def full(path):
    if path.endswith('.csv'):        # a `ncu -i x.ncu-rep --page raw --csv` dump made on the GPU box
        out = open(path).read()
    else:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in out.splitlines() if l.startswith('"')]))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    for vals in rows[2:]:
        print(f'## {short(vals[kn])}  (id {vals[0]})')
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f'  {h:70s} {vals[i]:>18s} {units[i]}')
        try:
            rd = float(vals[hdr.index('dram__bytes_read.sum')].replace(',', ''))
            wr = float(vals[hdr.index('dram__bytes_write.sum')].replace(',', ''))
            ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
            print(f'  traffic = dram read {rd} {ur} + write {wr} {uw}')
        except (ValueError, IndexError):
            pass


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
