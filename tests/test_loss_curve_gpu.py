"""GPU: 200 fixed-seed training steps of hop_b200 (Model + train_llm on the CUDA kernels) against the loss curve of
the REFERENCE train_llm (tests/golden/loss_curve_ted.npz, produced by tests/golden/make_golden.py curve on the CPU).
Same initial weights (same seed + same initialiser order), same batches, dropout p = 0, shared reparameterize noise
and speaker permutation."""
import os

import numpy as np
import pytest
import torch

from tests.golden.make_golden import (CURVE_B, CURVE_STEPS, MODEL_SEED, DummySpk, DummyTok, NoiseSource, PlainAccelerator,
                                      build_bert, curve_args, curve_batch, model_configs)
from tests.test_model_oracle import weights_match_golden
from tests.util import GOLDEN

pytestmark = pytest.mark.gpu


def run_curve(cuda, monkeypatch, precision, steps):
    from hop_b200 import HOP, train_llm as step_mod
    from hop_b200.discriminator import ConvDiscriminator
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(MODEL_SEED)
    bert = build_bert()
    m = HOP.Model(model_configs('TED'), bert, DummyTok(), DummySpk()).float()
    fix = np.load(os.path.join(GOLDEN, 'hop_model_ted.npz'))
    same_init, _ = weights_match_golden(m.state_dict(), fix)
    m.reprogramming_layer.dropout.p = 0.0
    disc = ConvDiscriminator(27)
    m, disc = m.to(cuda).set_precision(precision), disc.to(cuda)
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=4e-4, betas=(0.5, 0.999))
    dopt = torch.optim.Adam(disc.parameters(), lr=4e-4, betas=(0.5, 0.999))
    src = NoiseSource()
    monkeypatch.setattr(HOP, 'reparameterize', lambda mu, logvar: mu + src.noise().to(cuda) * torch.exp(0.5 * logvar))
    monkeypatch.setattr(step_mod.torch, 'randperm', lambda n, **kw: src.perm().to(cuda))
    rows = []
    for step in range(steps):
        b = {k: torch.from_numpy(v).to(cuda) for k, v in curve_batch(step).items()}
        ret = step_mod.train_llm(curve_args(), 1, b['in_audio'], b['melspec'], b['text'], b['target'], b['vid'], m, disc,
                                 opt, dopt, PlainAccelerator())
        rows.append([ret.get('loss', 0.0), ret.get('KLD', 0.0), ret.get('DIV_REG', 0.0)])
    return np.array(rows), same_init


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_loss_curve_tracks_reference(precision, cuda, monkeypatch):
    path = os.path.join(GOLDEN, 'loss_curve_ted.npz')
    ref = np.load(path)['curve']
    assert ref.shape == (CURVE_STEPS, 3)
    got, same_init = run_curve(cuda, monkeypatch, precision, CURVE_STEPS)
    out = os.path.join(os.path.dirname(GOLDEN), '..', 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    np.savetxt(os.path.join(out, f'loss_curve_{precision}.txt'), np.concatenate([ref, got], 1), fmt='%.6f',
               header='ref_loss ref_kld ref_div ours_loss ours_kld ours_div')
    if not same_init:
        pytest.skip('this CPU draws a different torch RNG stream for the initialisers than the build container')
    rel = np.abs(got[:, 0] - ref[:, 0]) / ref[:, 0]
    # The regression loss (600 x huber) is the curve the reference logs.  Training at the reference's settings
    # (Adam 4e-4, betas (0.5, 0.999), weight 600) is chaotic: two fp32 trajectories started 1e-7 apart separate
    # exponentially (measured: 1e-6 at step 10, 1e-4 at step 20, 1e-3 at step 50) and after a loss spike near step 100
    # they sit in different basins.  So: step-for-step agreement while the trajectories are numerically comparable,
    # and the same loss level afterwards.  (profiles/r1_loss_curve_*.txt hold the measured curves.)
    # bf16 mode: the same code gives 2.8 %, 3.4 % and 3.5 % over the first ten steps in three consecutive runs (the order of the
    # split-K reductions differs from run to run and steps 8-9 already amplify it): 5 % bounds the first ten, 8 % the first sixty.
    first10, first60 = (2e-4, 5e-3) if precision == 'fp32' else (5e-2, 8e-2)
    assert rel[:10].max() < first10, rel[:10]
    assert rel[:60].max() < first60, (rel[:60].max(), int(rel[:60].argmax()))
    tail_ref, tail_got = ref[100:, 0].mean(), got[100:, 0].mean()
    assert abs(tail_got - tail_ref) / tail_ref < (0.05 if precision == 'fp32' else 0.10), (tail_ref, tail_got)
    assert np.isfinite(got).all() and got[-20:, 0].mean() < got[:5, 0].mean()      # it trains
