"""profiles/r1_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per call of each hand-written kernel
group, from ncu metric passes around scripts/traffic_probe.py (N calls each; backward = (forward+backward) - forward).

    python scripts/traffic_from_ncu.py <dir with traffic_{gwnet,xattn}_{fwd,fwdbwd}.csv> <N>
"""
import csv, json, os, sys

def total(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    t = 0.0
    for row in csv.DictReader(lines):
        if not any(s in row['Kernel Name'] for s in ('hopk', 'gemm_tc', 'fz_', 'xattn', 'gram_', 'node_mix', 'bn_', 'adp_', 'nchw', 'fill_identity')):
            continue
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        t += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(row.get('Metric Unit', 'byte'), 1)
    return t

d, n = sys.argv[1], int(sys.argv[2])
out = {}
for g in ('gwnet', 'xattn'):
    f = total(os.path.join(d, f'traffic_{g}_fwd.csv')) / n
    fb = total(os.path.join(d, f'traffic_{g}_fwdbwd.csv')) / n
    out[g + '_fwd'] = f
    out[g + '_bwd'] = fb - f
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r1_traffic.json')
json.dump(out, open(dst, 'w'), indent=1)
print(out)
