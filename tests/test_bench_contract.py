"""CPU: host logic of bench.py's roofline object -- the dominant group is chosen by time per step, achieved / frac are
algorithmic work over CUDA-event time, the committed traffic file is attached, and the contract keys are present."""
import json
import types

import bench


def test_roofline_object_from_synthetic_spans():
    a = types.SimpleNamespace(batch=128, precision='bf16')
    # (calls, total ms) over 5 profiled steps: 490 GEMM launches taking 20 ms in total, 5 Graph-WaveNet forwards etc.
    spans = {'gemm_tma': (490, 20.0), 'gwnet_fwd': (5, 1.6), 'gwnet_bwd': (5, 3.2), 'xattn_fwd': (10, 1.0), 'xattn_bwd': (5, 1.7),
             'bert_fwd': (10, 11.0)}
    work = {'gemm_tma': 7.0e12}
    r = bench.roofline(spans, a, 'TED', 5, work)
    assert r['kernel'] == 'gemm_tma' and r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s'
    assert abs(r['achieved'] - 7.0e12 / 20.0e-3 / 1e12) < 1e-9
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12 and 0 < r['frac'] < 1
    assert abs(r['ms_per_step'] - 4.0) < 1e-12 and r['launches_timed'] == 490
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert key in r
    g = r['groups']
    assert set(g) == {'gemm_tma', 'gwnet_fwd', 'gwnet_bwd', 'xattn_fwd', 'xattn_bwd'}
    assert g['gwnet_fwd']['bound'] == 'hbm' and g['gwnet_fwd']['unit'] == 'GB/s'
    assert abs(g['gwnet_fwd']['algorithmic_bytes'] - 87902208.0) < 1 and g['gwnet_bwd']['algorithmic_bytes'] == 2 * g['gwnet_fwd']['algorithmic_bytes']
    assert g['xattn_bwd']['executed_flops'] == 1.4 * g['xattn_bwd']['algorithmic_flops']
    # the committed ncu traffic (profiles/r2_traffic.json) rides along for the TED B = 128 configuration
    with open(bench.TRAFFIC_FILE) as f:
        t = json.load(f)
    assert r['traffic'] == t['gemm_tma'] and g['gwnet_bwd']['traffic'] == t['gwnet_bwd']
    # without the per-launch GEMM pass the dominant entry falls back to the largest remaining group
    r2 = bench.roofline({k: v for k, v in spans.items() if k != 'gemm_tma'}, a, 'TED', 5, None)
    assert r2['kernel'] == 'gwnet_bwd'
