"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# Tolerances of BASELINE.json's north_star: fp32 mode 1e-5, bf16 mode 2e-2, both measured as
# max|a - ref| / max|ref| per tensor ("relative to the tensor's scale").
TOL_FP32 = 1e-5
TOL_BF16 = 2e-2


def relerr(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30))


def l2err(a, ref):
    """Relative Frobenius error ||a - ref|| / ||ref||: the bf16-mode metric.  A single ReLU-mask flip (an activation
    within rounding distance of zero) changes one gradient element completely; max-norm over a tensor built from a
    handful of rows is dominated by such flips for ANY reduced-precision implementation, the 2-norm is not."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return float(np.linalg.norm(a - ref) / (np.linalg.norm(ref) + 1e-30))


def hash_name(name):
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def sample_idx(name, n, count=512):
    """Same seeded subsample as tests/golden/make_golden.py::sample_idx."""
    rs = np.random.RandomState(abs(hash_name(name)) % (2 ** 31))
    return rs.randint(0, n, size=count)


def golden_compare(fix, key, value, zero_scale=1.0):
    """Compare ``value`` with the packed golden entry ``key`` (full tensor or sample + norm). Returns relerr."""
    v = np.asarray(value, np.float64).ravel()
    ref_any = fix['full:' + key] if 'full:' + key in fix else fix['samp:' + key]
    if np.abs(ref_any).max() < 1e-9:
        # analytically-zero gradient (a bias in front of a train-mode BatchNorm, the key bias under a
        # softmax): the reference holds rounding noise, so compare absolutely (zero_scale sets the unit)
        return float(np.abs(v).max()) / zero_scale
    if 'full:' + key in fix:
        return relerr(v, fix['full:' + key])
    ref = fix['samp:' + key]
    scale = float(fix['norm:' + key]) / np.sqrt(v.size) + 1e-30      # rms of the reference tensor
    e_s = float(np.abs(v[sample_idx(key, v.size)] - ref).max() / max(np.abs(ref).max(), scale))
    e_n = abs(float(np.linalg.norm(v)) - float(fix['norm:' + key])) / (float(fix['norm:' + key]) + 1e-30)
    return max(e_s, e_n)


class Report:
    """Collects per-tensor errors, dumps them to gpurun_out/ for post-mortem, asserts once at the end."""

    def __init__(self, name, tol):
        self.name, self.tol, self.rows = name, tol, []

    def add(self, what, err, tol=None):
        self.rows.append((what, float(err), float(self.tol if tol is None else tol)))

    def finish(self):
        out_dir = os.path.join(ROOT, 'gpurun_out')
        try:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, f'parity_{self.name}.json'), 'w') as f:
                json.dump([dict(tensor=w, err=e, tol=t, ok=bool(e <= t)) for w, e, t in self.rows], f, indent=1)
        except OSError:
            pass
        bad = [(w, e, t) for w, e, t in self.rows if not (e <= t)]
        worst = max(self.rows, key=lambda r: r[1] / r[2]) if self.rows else None
        assert not bad, f'{self.name}: {len(bad)}/{len(self.rows)} tensors out of tolerance; worst {worst}; first {bad[:8]}'
        return worst
