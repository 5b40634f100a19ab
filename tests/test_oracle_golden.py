"""CPU: the numpy oracle must reproduce the reference-generated golden fixtures (tests/golden/*.npz).

The fixtures were produced by tests/golden/make_golden.py, which executes the reference's
model/gwnet.py and model/HOP.py::ReprogrammingLayer in float64; this is what pins the oracle.
"""
import os

import numpy as np
import pytest

from oracle import gwnet_np, reprog_np
from tests.golden.make_golden import GW_CASES, RP_CASES, gw_inputs, rp_inputs
from tests.util import GOLDEN, golden_compare, relerr


@pytest.mark.parametrize('name', list(GW_CASES))
def test_gwnet_oracle_matches_reference(name):
    seed, B, V, T, cfg = GW_CASES[name]
    training = not name.endswith('_eval')
    fix = np.load(os.path.join(GOLDEN, name + '.npz'))
    P, x, dout = gw_inputs(seed, B, V, T, cfg)
    out, bufs, cache = gwnet_np.forward(P, x, training=training, keep=True)
    dx, G = gwnet_np.backward(P, cache, dout)
    assert relerr(out, fix['out']) < 1e-9
    assert golden_compare(fix, 'dx', dx) < 1e-8
    none_grads = set(fix['none_grads'].tolist())
    for k in P:
        if k in none_grads or 'running_' in k or 'num_batches' in k:
            assert k not in G
            continue
        if k.startswith('residual_convs'):
            continue
        assert golden_compare(fix, k, G[k]) < 1e-7, k
    if training:
        for k, v in bufs.items():
            assert relerr(v, fix['buf:' + k]) < 1e-12, k


@pytest.mark.parametrize('name', list(RP_CASES))
def test_reprog_oracle_matches_reference(name):
    seed, B, L, S, cfg = RP_CASES[name]
    fix = np.load(os.path.join(GOLDEN, name + '.npz'))
    P, x, src, dY = rp_inputs(seed, B, L, S, cfg)
    y, cache = reprog_np.forward(P, x, src, src, cfg['n_heads'], keep=True)
    dx, ds, dv, G = reprog_np.backward(P, cache, dY, cfg['n_heads'])
    assert relerr(y, fix['out']) < 1e-10
    assert golden_compare(fix, 'dx', dx) < 1e-9
    assert golden_compare(fix, 'dsource', ds + dv) < 1e-9
    for k in G:
        assert golden_compare(fix, k, G[k]) < 1e-9, k


def test_dropout_mask_statistics():
    idx = np.arange(1 << 20, dtype=np.uint64)
    for p in (0.1, 0.5):
        keep = reprog_np.dropout_keep(1234567, idx, p)
        assert abs(keep.mean() - (1 - p)) < 3e-3
    a = reprog_np.dropout_keep(1, idx, 0.1)
    b = reprog_np.dropout_keep(2, idx, 0.1)
    assert (a != b).mean() > 0.1          # different seeds decorrelate
    hi = reprog_np.dropout_keep(1, idx + np.uint64(1 << 40), 0.1)
    assert (a != hi).mean() > 0.1         # the high index word matters
