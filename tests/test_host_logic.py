"""CPU: host-side logic that needs no kernels.

* the BatchNorm "one more update from the same batch statistics" recurrence used when the forwards of a training step
  share the Graph-WaveNet output (hop_b200/HOP.py::_repeat_bn_update) equals what running nn.BatchNorm2d a second and a
  third time on the same input does to the running buffers;
* finish_losses drops falsy KLD / DIV_REG entries exactly like the reference's `if kld:` (train_eval/train_llm.py:88-98)."""
import copy
import types

import torch
import torch.nn as nn


def test_repeat_bn_update_matches_repeated_forward():
    from hop_b200.HOP import Model
    torch.manual_seed(0)
    bns = nn.ModuleList([nn.BatchNorm2d(8) for _ in range(3)])
    for bn in bns:                                   # non-trivial starting buffers
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    ref = copy.deepcopy(bns)
    xs = [torch.randn(4, 8, 5, 7) * (i + 1) + i for i in range(3)]
    fake = types.SimpleNamespace(gwnet=types.SimpleNamespace(bn=bns), training=True)
    fake._gwnet_bn_buffers = lambda: Model._gwnet_bn_buffers(fake)
    shared = {'bn_prev': [b.clone() for b in fake._gwnet_bn_buffers()]}
    for bn, x in zip(bns, xs):                       # the one real forward of the step
        bn.train()(x)
    for _ in range(2):                               # second and third forward: buffers only
        Model._repeat_bn_update(fake, shared)
    for bn, x in zip(ref, xs):
        for _ in range(3):
            bn.train()(x)
    for a, b in zip(bns, ref):
        assert torch.allclose(a.running_mean, b.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(a.running_var, b.running_var, rtol=1e-5, atol=1e-6)
        assert int(a.num_batches_tracked) == int(b.num_batches_tracked) == 3


def test_finish_losses_drops_falsy_regularisers():
    from hop_b200.train_llm import finish_losses
    assert finish_losses(['loss', 'KLD', 'DIV_REG'], [1.5, 0.0, -2.0]) == {'loss': 1.5, 'DIV_REG': -2.0}
    assert finish_losses(['loss', 'gen', 'dis'], [1.0, 0.0, 0.0]) == {'loss': 1.0, 'gen': 0.0, 'dis': 0.0}
