"""Micro-benchmarks of the hand-written kernels at BASELINE sizes (CUDA events, L2 flushed between iterations).

    python scripts/bench_ops.py xattn|gwnet [--precision bf16|fp32] [--B 128] [--V 9] [--iters 20]
Prints one JSON line per op with avg ms, achieved TFLOP/s or GB/s and the fraction of the measured peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return p['hbm_gbs'], p['bf16_tflops'], 'measured'
    except (OSError, KeyError, ValueError):
        return 6650.0, 1590.0, 'fallback'


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()                                   # > L2 (126 MB): evict the working set
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(3_000_000)                    # ~1.5 ms of GPU spin: the host runs ahead, so host-side launch
        e0.record(); fn(); e1.record()                  # preparation is not inside the timed region
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['xattn', 'gwnet'])
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--B', type=int, default=128)
    ap.add_argument('--V', type=int, default=9)
    ap.add_argument('--C', type=int, default=64)
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    dev = torch.device('cuda')
    hbm, tf, src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    if a.what == 'xattn':
        from hop_b200.HOP import _XattnFn
        B, L, H, E, S = a.B, 34, 8, 128, 1500
        q = torch.randn(B, L, H, E, device=dev, requires_grad=True)
        k = torch.randn(S, H, E, device=dev, requires_grad=True)
        v = torch.randn(S, H, E, device=dev, requires_grad=True)
        do = torch.randn(B, L, H, E, device=dev)
        tc = a.precision == 'bf16'
        with torch.no_grad():
            t_f = timeit(lambda: _XattnFn.apply(q, k, v, 0.1, 5, tc), a.iters, flush)
        o = _XattnFn.apply(q, k, v, 0.1, 5, tc)
        t_b = timeit(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True), a.iters, flush)
        unit = 4.0 * B * L * S * H * E
        for name, t, fl in (('xattn_fwd', t_f, unit), ('xattn_bwd', t_b, 3.5 * unit)):
            print(json.dumps({'op': name, 'precision': a.precision, 'B': B, 'ms': t, 'flops': fl, 'tflops': fl / t / 1e9,
                              'frac_of_bf16_peak': fl / t / 1e9 / tf, 'peak': tf, 'peak_source': src}))
    else:
        from hop_b200 import gwnet as G
        m = G.gwnet(dev, a.V, dropout=0, in_dim=173, out_dim=173, residual_channels=a.C, dilation_channels=a.C,
                    skip_channels=256, end_channels=512).to(dev).set_precision(a.precision)
        x = torch.randn(a.B, 16, a.V, 173, device=dev).permute(0, 3, 2, 1).requires_grad_(True)
        dy = torch.randn(a.B, 173, a.V, 4, device=dev)
        with torch.no_grad():
            t_f = timeit(lambda: m(x), a.iters, flush)

        def fb():
            m.zero_grad(set_to_none=True)
            x.grad = None
            y = m(x)
            y.backward(dy)
        t_fb = timeit(fb, a.iters, flush)
        s = 4
        floor = s * a.B * a.V * (173 * 16 + a.C * 16 + a.C * (88 + 76) + 2 * 8 * a.C * 4 + 173 * 4)
        for name, t, by in (('gwnet_fwd', t_f, floor), ('gwnet_fwd+bwd', t_fb, 3 * floor)):
            print(json.dumps({'op': name, 'precision': a.precision, 'B': a.B, 'V': a.V, 'C': a.C, 'ms': t,
                              'algorithmic_bytes': by, 'gbs': by / t / 1e6, 'frac_of_hbm_peak': by / t / 1e6 / hbm,
                              'peak': hbm, 'peak_source': src}))


if __name__ == '__main__':
    main()
