"""CUDA-event timing of the kernel entry points, live inside a training step.

``enable()`` makes every autograd node of this package bracket its C-ABI call with two events on the
launching stream; ``summary()`` synchronises and returns {op: (calls, total_ms)}.  Used by bench.py for the
roofline object; off by default (two event records per call are cheap but not free).
"""
import contextlib

import torch

_on = False
_fine = False
_spans = []


def enable(flag=True, fine=False):
    """``fine``: also time the spans marked ``fine`` (single launches nested inside a group, e.g. every dense GEMM); kept
    out of the default pass because their event records widen the host gaps inside the enclosing group."""
    global _on, _fine
    _on, _fine = flag, bool(flag and fine)
    _spans.clear()


def enabled():
    return _on


@contextlib.contextmanager
def span(name, work=0.0, fine=False):
    """``work``: algorithmic FLOPs (or bytes) of the bracketed call, summed per name by ``work_summary``."""
    if not _on or (fine and not _fine):
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    _spans.append((name, e0, e1, float(work)))


def summary():
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, _ in _spans:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + e0.elapsed_time(e1))
    return out


def work_summary():
    out = {}
    for name, _, _, w in _spans:
        out[name] = out.get(name, 0.0) + w
    return out
