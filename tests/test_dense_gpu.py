"""GPU parity of hop_b200.dense (nn.Linear / text prototypes / beat rows on the TMA + tcgen05 GEMM) against float64 torch.
Operands that are bf16-representable isolate the kernels' own arithmetic (fp32 accumulation: 1e-5); general fp32 inputs
add the operand rounding of the mode (bf16: 2^-9 per operand), checked at the mode's 2e-2."""
import numpy as np
import pytest
import torch

from tests.util import TOL_BF16, Report, relerr

pytestmark = pytest.mark.gpu
npy = lambda t: t.detach().double().cpu().numpy()


@pytest.mark.parametrize('M,K,N,relu_in', [(4352, 128, 1024, False), (1500, 768, 1024, False), (4352, 1024, 768, True), (70, 40, 24, True),
                                           (4352, 1536, 768, False)])
def test_dense_linear(M, K, N, relu_in, cuda):
    from hop_b200 import dense
    g = torch.Generator(device='cpu').manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).bfloat16().float().to(cuda).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().float().to(cuda).requires_grad_(True)
    b = torch.randn(N, generator=g).to(cuda).requires_grad_(True)
    dy = torch.randn(M, N, generator=g).bfloat16().float().to(cuda)
    y = dense.linear(x, w, b, relu_in)
    y.backward(dy)
    xr, wr, br = [t.detach().double().requires_grad_(True) for t in (x, w, b)]
    yr = torch.nn.functional.linear(torch.relu(xr) if relu_in else xr, wr, br)
    yr.backward(dy.double())
    rep = Report(f'dense_linear_{M}_{K}_{N}_{int(relu_in)}', 2e-5)
    rep.add('y', relerr(npy(y), npy(yr)))
    rep.add('dx', relerr(npy(x.grad), npy(xr.grad)))
    rep.add('dw', relerr(npy(w.grad), npy(wr.grad)))
    rep.add('db', relerr(npy(b.grad), npy(br.grad)))
    rep.finish()


def test_dense_source(cuda):
    """Text prototypes W_map @ WE + b (HOP.py:200) at the real width of the mapping layer (30522 -> 1500 x 768)."""
    from hop_b200 import dense
    g = torch.Generator(device='cpu').manual_seed(1)
    S, V, D = 1500, 30522, 768
    w = (torch.randn(S, V, generator=g) / V ** 0.5).bfloat16().float().to(cuda).requires_grad_(True)
    b = torch.randn(S, generator=g).to(cuda).requires_grad_(True)
    we = torch.randn(V, D, generator=g).bfloat16().to(cuda)
    dsrc = torch.randn(S, D, generator=g).bfloat16().float().to(cuda)
    src = dense.source(w, b, we)
    src.backward(dsrc)
    wr, br = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    ref = wr @ we.double() + br[:, None]
    ref.backward(dsrc.double())
    rep = Report('dense_source', 1e-4)
    rep.add('source', relerr(npy(src), npy(ref)))
    rep.add('dW_map', relerr(npy(w.grad), npy(wr.grad)))
    rep.add('db_map', relerr(npy(b.grad), npy(br.grad)))
    rep.finish()


@pytest.mark.parametrize('B,J', [(128, 9), (6, 42)])
def test_dense_beat_rows(B, J, cuda):
    """unfold + beat MLP + the reference's repeat / view index map + concat with the seed bones (HOP.py:210-217), forward
    and the gradients of the two Linear layers, against the reference formulation in float64 (J-fold repeat included).
    The gradients are checked against the reference pinned to the kernel's own LeakyReLU gate pattern: bf16 operand rounding
    flips ~0.2 % of the gates (pre-activations within rounding distance of zero) and one flipped gate moves a row of dW_1 by
    ~1/sqrt(rows) of its scale -- the same effect as the head ReLUs of Graph-WaveNet (tests/test_baseline_parity_gpu.py)."""
    from hop_b200 import dense
    dense.KEEP = True
    torch.manual_seed(B + J)
    beat = torch.nn.Sequential(torch.nn.Linear(3400, 1700), torch.nn.LeakyReLU(0.2), torch.nn.Linear(1700, 170)).to(cuda)
    audio = 0.1 * torch.randn(B, 36267, device=cuda)
    seed = torch.randn(B, 16, 3 * J, device=cuda)
    drows = torch.randn(B, 16, J, 173, device=cuda)
    rows = dense.beat_rows(audio, seed, beat, J)
    rows.backward(drows)
    ref = torch.nn.Sequential(torch.nn.Linear(3400, 1700), torch.nn.LeakyReLU(0.2), torch.nn.Linear(1700, 170)).to(cuda).double()
    ref.load_state_dict({k: v.double() for k, v in beat.state_dict().items()})
    win = audio.double().unfold(1, 3400, 2191).unsqueeze(1).repeat(1, J, 1, 1)          # HOP.py:210
    feat = ref(win).view(B, 16, J, 170)                                                   # HOP.py:211-212 (reinterpretation)
    rref = torch.cat([seed.double().view(B, 16, J, 3), feat], dim=3)                      # HOP.py:214
    rep = Report(f'dense_beat_rows_{B}_{J}', TOL_BF16)
    rep.add('rows', relerr(npy(rows), npy(rref)))
    # pinned gates: LeakyReLU(pre) == pre * (1 or 0.2) with the factor taken from the kernel's own activation signs
    gate = torch.where(dense.LAST['beat_h1'].double() > 0, 1.0, 0.2).view(B, 1, 16, 1700)           # same for all J copies
    pre = ref[0](win)
    rep.add('flipped LeakyReLU gates (fraction)', float(((pre > 0) != (gate > 0.5)).double().mean()), tol=1e-2)
    feat = ref[2](pre * gate).view(B, 16, J, 170)
    rref = torch.cat([seed.double().view(B, 16, J, 3), feat], dim=3)
    rref.backward(drows.double())
    dense.KEEP = False
    for (k, p), (_, q) in zip(beat.named_parameters(), ref.named_parameters()):
        rep.add('grad:' + k, relerr(npy(p.grad), npy(q.grad)))
    rep.finish()
