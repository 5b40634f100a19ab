"""Turn ncu outputs into the small text summaries that are committed under profiles/.

    python scripts/summarise_ncu.py launches <launches.csv> [--steps N]     # per-kernel totals + share of the run
    python scripts/summarise_ncu.py full <report.ncu-rep>                   # roofline-relevant counters per captured kernel

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`;
`full` shells out to `ncu -i <rep> --page raw --csv` (ncu must be on PATH; no GPU needed).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'smsp__cycles_active.avg',
    'sm__cycles_elapsed.max',
]


def short(name, n=96):
    name = name.replace('hopk::', '').replace('void ', '')
    return name[:n]


def launches(path, steps=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    tot, cnt = collections.Counter(), collections.Counter()
    ours_names = set()
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (KeyError, ValueError):
            continue
        unit = row.get('Metric Unit', 'ns')
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
        k = short(row['Kernel Name'])
        if 'hopk::' in row['Kernel Name']:
            ours_names.add(k)
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    n = sum(cnt.values())
    print(f'# {path}: {n} launches, {total / 1e3:.3f} ms of kernel time (ncu: serialised, cold caches)')
    if steps:
        print(f'# per step ({steps} steps captured): {n / steps:.0f} launches, {total / steps / 1e3:.3f} ms')
    ours = sum(v for k, v in tot.items() if k in ours_names)
    nours = sum(c for k, c in cnt.items() if k in ours_names)
    print(f'# hand-written kernels (namespace hopk, libhopk.so): {nours} launches, {ours / 1e3:.3f} ms = {100 * ours / max(total, 1e-9):.1f} % of the kernel time')
    print(f'{"us total":>10} {"share":>6} {"count":>6} {"us avg":>8}  kernel   (* = hand-written, namespace hopk)')
    for k, v in tot.most_common():
        print(f'{v:10.1f} {100 * v / total:5.1f}% {cnt[k]:6d} {v / cnt[k]:8.1f}  {"*" if k in ours_names else " "} {k}')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(head)}
    print(f'# {path}: {len(rows) - 2} captured launches (ncu --set full --clock-control none)')
    for r in rows[2:]:
        print(f'\n## {short(r[ix["Kernel Name"]], 140)}')
        for m in FULL_METRICS:
            if m in ix:
                print(f'{m:80s} {r[ix[m]]:>16s} {units[ix[m]]}')
        stalls = [(float(r[i] or 0), n) for n, i in ix.items()
                  if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio') and 'not_issued' not in n]
        stalls.sort(reverse=True)
        for v, n in stalls[:6]:
            print(f'{n:80s} {v:16.2f}')


if __name__ == '__main__':
    if len(sys.argv) < 3 or sys.argv[1] not in ('launches', 'full'):
        sys.exit(__doc__)
    if sys.argv[1] == 'launches':
        st = int(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else None
        launches(sys.argv[2], st)
    else:
        full(sys.argv[2])
