"""CPU: (1) hop_b200.HOP.Model keeps the reference's 314 state_dict keys and, under the same seed, the
reference's initial weights; (2) the functional torch oracle (oracle/hop_torch.py) reproduces the
reference HOP.Model's outputs and gradients stored in tests/golden/hop_model_*.npz."""
import os

import numpy as np
import pytest
import torch

from oracle import hop_torch
from tests.golden.make_golden import (CHECK_PARAMS, GRAD_PARAMS, MODEL_SEED, DummySpk, DummyTok, build_bert, model_configs,
                                      model_inputs)
from tests.util import GOLDEN, golden_compare, relerr


def build_model(datasets):
    from hop_b200.HOP import Model
    torch.manual_seed(MODEL_SEED)
    bert = build_bert()
    return Model(model_configs(datasets), bert, DummyTok(), DummySpk()).float(), bert


def weights_match_golden(sd, fix):
    for k in CHECK_PARAMS:
        got = np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum())])
        if not np.allclose(got, fix['wsum:' + k], rtol=1e-9, atol=1e-9):
            return False, k
    return True, None


@pytest.mark.parametrize('datasets', ['TED', 'TED_expressive'])
def test_model_keys_init_and_torch_oracle(datasets):
    fix = np.load(os.path.join(GOLDEN, 'hop_model_' + ('ted' if datasets == 'TED' else 'expr') + '.npz'))
    m, bert = build_model(datasets)
    sd = m.state_dict()
    assert len(sd) == 314 == int(fix['n_keys'])
    assert sorted(sd.keys()) == fix['keys'].tolist()              # key-for-key the reference's state_dict
    ok, bad = weights_match_golden(sd, fix)
    assert ok, f'same-seed initial weights differ from the reference at {bad}'
    # oracle forward/backward in fp32 on the CPU, sharing the reparameterize noise with the golden run
    inp = model_inputs(datasets)
    t = lambda k: torch.from_numpy(inp[k])
    params = dict(m.named_parameters())
    osd = {k: (params[k] if k in params else v) for k, v in sd.items()}
    out, z, z_mu, z_lv = hop_torch.model_forward(osd, bert, t('in_audio'), t('x_enc'), t('text'), t('pre_seq'), t('vid'),
                                                 t('noise'))
    loss = (out * t('d_out')).sum() + (z_mu * t('d_mu')).sum() + (z_lv * t('d_lv')).sum()
    loss.backward()
    tol = 2e-4                                                    # fp32 on both sides, different op order
    assert relerr(out.detach().numpy(), fix['out']) < tol
    assert relerr(z_mu.detach().numpy(), fix['z_mu']) < tol
    assert relerr(z.detach().numpy(), fix['z']) < tol
    for k in GRAD_PARAMS:
        assert params[k].grad is not None, k
        assert golden_compare(fix, k, params[k].grad.numpy()) < 5e-4, k
    none_grads = sorted(k for k, p_ in params.items() if p_.requires_grad and p_.grad is None)
    assert none_grads == fix['none_grads'].tolist()
