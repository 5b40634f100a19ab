"""Timing of the hand-written GRU decoder (hop_b200.gru) next to cuDNN's (torch.nn.GRU) at the HOP decoder size.

    python scripts/bench_gru.py [--B 128] [--layers 4] [--iters 20] [--once]      # --once: one forward + backward (for ncu)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hop_b200 import gru as hgru  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=128)
    ap.add_argument('--T', type=int, default=34)
    ap.add_argument('--I', type=int, default=992)
    ap.add_argument('--layers', type=int, default=4)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--once', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    gru = torch.nn.GRU(a.I, hidden_size=350, num_layers=a.layers, batch_first=True, bidirectional=True).to(dev)
    x = torch.randn(a.B, a.T, a.I, device=dev, requires_grad=True)
    dout = torch.randn(a.B, a.T, 700, device=dev)

    def ours_fwd():
        with torch.no_grad():
            hgru.run(gru, x)

    def ours_fb():
        y = hgru.run(gru, x)
        y.backward(dout)

    if a.once:
        ours_fb()
        torch.cuda.synchronize()
        return

    def cudnn_fwd():
        with torch.no_grad():
            gru(x)

    def cudnn_fb():
        y, _ = gru(x)
        y.backward(dout)
    res = {'ours_fwd_ms': timeit(ours_fwd, a.iters), 'ours_fwd_bwd_ms': timeit(ours_fb, a.iters)}
    torch.backends.cudnn.allow_tf32 = True
    res['cudnn_tf32_fwd_ms'] = timeit(cudnn_fwd, a.iters)
    res['cudnn_tf32_fwd_bwd_ms'] = timeit(cudnn_fb, a.iters)
    print({k: round(v, 3) for k, v in res.items()})


if __name__ == '__main__':
    main()
