"""profiles/r2_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per call of each hand-written kernel
group, from ncu metric passes (`--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv`):

  traffic_{gwnet,xattn}_{fwd,fwdbwd}.csv   around scripts/traffic_probe.py (N calls each; backward = (forward+backward) - forward)
  traffic_step_gemm.csv                    around `bench.py --profile-step` restricted to gemm_tma_kernel: the average over
                                           every dense-GEMM launch of one training step (the roofline's dominant kernel)

    python scripts/traffic_from_ncu.py <dir with the csv files> <N> [source note]
"""
import csv
import json
import os
import sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
OURS = ('hopk', 'gemm_tc', 'gemm_tma', 'fz_', 'xattn', 'gram_', 'node_mix', 'bn_', 'adp_', 'nchw', 'fill_identity', 'sums_to_float')


def rows(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        yield row['Kernel Name'], row.get('ID', ''), v * UNIT.get(row.get('Metric Unit', 'byte'), 1)


def total(path):
    return sum(v for name, _, v in rows(path) if any(s in name for s in OURS))


def main():
    d, n = sys.argv[1], int(sys.argv[2])
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    out = {}
    for g in ('gwnet', 'xattn'):
        pf, pfb = os.path.join(d, f'traffic_{g}_fwd.csv'), os.path.join(d, f'traffic_{g}_fwdbwd.csv')
        if os.path.exists(pf) and os.path.exists(pfb):
            f, fb = total(pf) / n, total(pfb) / n
            out[g + '_fwd'], out[g + '_bwd'] = f, fb - f
    pg = os.path.join(d, 'traffic_step_gemm.csv')
    if os.path.exists(pg):
        launches, tot = set(), 0.0
        for name, kid, v in rows(pg):
            if 'gemm_tma_kernel' in name:
                launches.add(kid)
                tot += v
        if launches:
            out['gemm_tma'] = tot / len(launches)
            out['gemm_tma_launches'] = len(launches)
    out['source'] = ('ncu dram__bytes_read.sum + dram__bytes_write.sum per call, TED B = 128 (scripts/traffic_probe.py; gemm_tma: mean over '
                     'the gemm_tma_kernel launches of one profiled step)' + (' -- ' + note if note else ''))
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r2_traffic.json')
    with open(dst, 'w') as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == '__main__':
    main()
