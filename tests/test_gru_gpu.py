"""GPU parity of the hand-written GRU decoder (csrc/gru.cu through hop_b200.gru.run) against torch.nn.GRU evaluated in
float64 on the same device (the checker).  dtype-1 arithmetic (bf16 tensor-core operands, fp32 accumulation and state):
2e-2 of each tensor's scale (north_star's bf16 tolerance), forward and every gradient."""
import numpy as np
import pytest
import torch

from tests.util import TOL_BF16, Report, l2err, relerr

pytestmark = pytest.mark.gpu


def _case(dev, B, T, I, H, L, seed):
    torch.manual_seed(seed)
    gru = torch.nn.GRU(I, hidden_size=H, num_layers=L, batch_first=True, bidirectional=True, dropout=0).to(dev)
    x = torch.randn(B, T, I, device=dev)
    dout = torch.randn(B, T, 2 * H, device=dev)
    return gru, x, dout


@pytest.mark.parametrize('B,T,I,H,L', [(128, 34, 992, 350, 4),      # TED decoder (model/HOP.py:146-167)
                                       (16, 34, 1751, 350, 4),      # Expressive decoder input width (odd: padded to 1752)
                                       (5, 7, 33, 350, 2),          # ragged batch (one partly filled 16-sample slice), short sequence
                                       (37, 3, 64, 96, 1)])         # smaller hidden size, several slices, single layer
def test_gru_vs_float64(B, T, I, H, L, cuda):
    from hop_b200 import gru as hgru
    gru, x, dout = _case(cuda, B, T, I, H, L, seed=B + T + I)
    ref = torch.nn.GRU(I, hidden_size=H, num_layers=L, batch_first=True, bidirectional=True, dropout=0).to(cuda).double()
    ref.load_state_dict({k: v.double() for k, v in gru.state_dict().items()})
    xr = x.double().requires_grad_(True)
    yr, _ = ref(xr)
    yr.backward(dout.double())
    xo = x.clone().requires_grad_(True)
    yo = hgru.run(gru, xo)
    yo.backward(dout)
    torch.cuda.synchronize()
    npy = lambda t: t.detach().double().cpu().numpy()
    rep = Report(f'gru_{B}_{T}_{I}_{H}_{L}', TOL_BF16)
    rep.add('out', relerr(npy(yo), npy(yr)))
    rep.add('out(l2)', l2err(npy(yo), npy(yr)))
    rep.add('dx', relerr(npy(xo.grad), npy(xr.grad)))
    rep.add('dx(l2)', l2err(npy(xo.grad), npy(xr.grad)))
    for (k, p), (_, q) in zip(gru.named_parameters(), ref.named_parameters()):
        rep.add('grad:' + k, relerr(npy(p.grad), npy(q.grad)))
    rep.finish()


def test_gru_no_grad_pass_matches(cuda):
    """The inference / no-grad path (nothing saved for backward) gives the same outputs as the training path."""
    from hop_b200 import gru as hgru
    gru, x, _ = _case(cuda, 20, 9, 40, 350, 2, seed=3)
    with torch.no_grad():
        a = hgru.run(gru, x)
    b = hgru.run(gru, x.clone().requires_grad_(True))
    assert torch.equal(a, b.detach())


def test_gru_rejects_cpu_and_unsupported(cuda):
    from hop_b200 import gru as hgru
    g = torch.nn.GRU(8, 16, batch_first=True, bidirectional=True)
    with pytest.raises(RuntimeError):
        hgru.run(g, torch.randn(2, 3, 8))
    with pytest.raises(NotImplementedError):
        hgru.run(torch.nn.GRU(8, 16, batch_first=True, bidirectional=False).to(cuda), torch.randn(2, 3, 8, device=cuda))
