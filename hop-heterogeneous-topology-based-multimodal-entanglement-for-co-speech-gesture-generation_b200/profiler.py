"""CUDA-event timing of the kernel entry points, live inside a training step.

``enable()`` makes every autograd node of this package bracket its C-ABI call with two events on the
launching stream; ``summary()`` synchronises and returns {op: (calls, total_ms)}.  Used by bench.py for the
roofline object; off by default (two event records per call are cheap but not free).
"""
import contextlib

import torch

_on = False
_spans = []


def enable(flag=True):
    global _on
    _on = flag
    _spans.clear()


@contextlib.contextmanager
def span(name):
    if not _on:
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    _spans.append((name, e0, e1))


def summary():
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _spans:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + e0.elapsed_time(e1))
    return out
