"""GPU: hop_b200.HOP.Model (kernels for gwnet + reprogramming, stock torch elsewhere) against the golden
outputs of the reference HOP.Model and against the torch oracle run on the CPU in float64."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import hop_torch
from tests.golden.make_golden import GRAD_PARAMS, model_inputs
from tests.test_model_oracle import build_model, weights_match_golden
from tests.util import GOLDEN, Report, golden_compare, relerr

pytestmark = pytest.mark.gpu
TOL_MODEL = 3e-4      # whole generator: fp32 cuBLAS/cuDNN (BERT, GRU) + our fp32 kernels vs an fp64 oracle


@pytest.mark.parametrize('datasets', ['TED', 'TED_expressive'])
def test_model_forward_backward(datasets, cuda, monkeypatch):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    fix = np.load(os.path.join(GOLDEN, 'hop_model_' + ('ted' if datasets == 'TED' else 'expr') + '.npz'))
    m, bert = build_model(datasets)
    same_init, _ = weights_match_golden(m.state_dict(), fix)
    inp = model_inputs(datasets)
    # fp64 CPU oracle on a deep copy of the very same weights
    m64 = copy.deepcopy(m).double()
    sd64 = dict(m64.state_dict())
    p64 = dict(m64.named_parameters())
    sd64.update(p64)
    t64 = lambda k: torch.from_numpy(inp[k]).double() if inp[k].dtype.kind == 'f' else torch.from_numpy(inp[k])
    o_out, o_z, o_mu, o_lv = hop_torch.model_forward(sd64, m64.llm_model, t64('in_audio'), t64('x_enc'), t64('text'),
                                                     t64('pre_seq'), t64('vid'), t64('noise'))
    ((o_out * t64('d_out')).sum() + (o_mu * t64('d_mu')).sum() + (o_lv * t64('d_lv')).sum()).backward()

    from hop_b200 import HOP
    m = m.to(cuda)
    m.reprogramming_layer.dropout.p = 0.0
    tg = lambda k: torch.from_numpy(inp[k]).to(cuda)
    noise = tg('noise')
    monkeypatch.setattr(HOP, 'reparameterize', lambda mu, logvar: mu + noise * torch.exp(0.5 * logvar))
    out, z, z_mu, z_lv = m(tg('in_audio'), tg('x_enc'), tg('text'), tg('pre_seq'), tg('vid'))
    ((out * tg('d_out')).sum() + (z_mu * tg('d_mu')).sum() + (z_lv * tg('d_lv')).sum()).backward()
    rep = Report('model_' + datasets, TOL_MODEL)
    rep.add('out', relerr(out.detach().cpu().numpy(), o_out.detach().numpy()))
    rep.add('z', relerr(z.detach().cpu().numpy(), o_z.detach().numpy()))
    if same_init:
        rep.add('out(golden)', relerr(out.detach().cpu().numpy(), fix['out']))
    params = dict(m.named_parameters())
    gscale = max(float(p.grad.abs().max()) for p in p64.values() if p.grad is not None)
    for k, p_ in params.items():
        if not p_.requires_grad:
            continue
        if p64[k].grad is None:
            assert p_.grad is None, f'{k}: reference gives no gradient'
            continue
        assert p_.grad is not None, k
        ref = p64[k].grad.numpy()
        if np.abs(ref).max() < 1e-9 * gscale:        # analytically zero (bias before train-mode BN, key bias)
            rep.add('grad0:' + k, float(p_.grad.abs().max()) / gscale, tol=1e-5)
            continue
        rep.add('grad:' + k, relerr(p_.grad.cpu().numpy(), ref), tol=1e-3 if ref.size < 64 else TOL_MODEL)
        if same_init and k in GRAD_PARAMS:
            rep.add('grad(golden):' + k, golden_compare(fix, k, p_.grad.cpu().numpy()), tol=1e-3)
    rep.finish()
