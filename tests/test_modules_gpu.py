"""GPU: the small stand-alone modules of model/gwnet.py (nconv, linear, gcn) on their kernels,
checked against the same ops in float64 torch on the CPU (what the reference would compute)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import TOL_FP32, Report, relerr

pytestmark = pytest.mark.gpu


def test_nconv_linear_gcn(cuda):
    from hop_b200 import gwnet as G
    torch.manual_seed(1)
    N, C, V, T = 3, 8, 7, 5
    x = torch.randn(N, C, V, T, dtype=torch.float64)
    A = torch.softmax(torch.randn(V, V, dtype=torch.float64), 1)
    rep = Report('modules', TOL_FP32)
    # nconv
    xr, Ar = x.clone().requires_grad_(True), A.clone().requires_grad_(True)
    ref = torch.einsum('ncvl,vw->ncwl', xr, Ar).contiguous()
    d = torch.randn_like(ref)
    ref.backward(d)
    xg, Ag = x.float().to(cuda).requires_grad_(True), A.float().to(cuda).requires_grad_(True)
    out = G.nconv()(xg, Ag)
    out.backward(d.float().to(cuda))
    rep.add('nconv.out', relerr(out.detach().cpu().numpy(), ref.detach().numpy()))
    rep.add('nconv.dx', relerr(xg.grad.cpu().numpy(), xr.grad.numpy()))
    rep.add('nconv.dA', relerr(Ag.grad.cpu().numpy(), Ar.grad.numpy()))
    # linear
    lin = G.linear(C, 11).to(cuda)
    w64, b64 = lin.mlp.weight.detach().double().cpu().requires_grad_(True), lin.mlp.bias.detach().double().cpu().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv2d(xr, w64, b64)
    d = torch.randn_like(ref)
    ref.backward(d)
    xg = x.float().to(cuda).requires_grad_(True)
    out = lin(xg)
    out.backward(d.float().to(cuda))
    rep.add('linear.out', relerr(out.detach().cpu().numpy(), ref.detach().numpy()))
    rep.add('linear.dx', relerr(xg.grad.cpu().numpy(), xr.grad.numpy()))
    rep.add('linear.dw', relerr(lin.mlp.weight.grad.cpu().numpy(), w64.grad.numpy()))
    rep.add('linear.db', relerr(lin.mlp.bias.grad.cpu().numpy(), b64.grad.numpy()))
    # gcn (order 2, one support, dropout 0)
    g = G.gcn(C, 6, 0.0, support_len=1).to(cuda)
    w64 = g.mlp.mlp.weight.detach().double().cpu(); b64 = g.mlp.mlp.bias.detach().double().cpu()
    x1 = torch.einsum('ncvl,vw->ncwl', x, A); x2 = torch.einsum('ncvl,vw->ncwl', x1, A)
    ref = F.conv2d(torch.cat([x, x1, x2], 1), w64, b64)
    out = g(x.float().to(cuda), [A.float().to(cuda)])
    rep.add('gcn.out', relerr(out.detach().cpu().numpy(), ref.numpy()))
    rep.finish()


def test_linear_fn_flags(cuda):
    from hop_b200.HOP import _LinearFn
    torch.manual_seed(2)
    M, K, N = 70, 37, 45
    x = torch.randn(M, K, dtype=torch.float64); w = torch.randn(N, K, dtype=torch.float64) * 0.2
    b = torch.randn(N, dtype=torch.float64)
    rep = Report('linear_flags', TOL_FP32)
    for flags in (0, 1, 2, 3):
        xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
        xin = torch.relu(xr) if flags & 1 else xr
        ref = xin @ wr.T + br
        if flags & 2:
            ref = torch.relu(ref)
        d = torch.randn_like(ref)
        ref.backward(d)
        xg, wg, bg = [t.float().to(cuda).requires_grad_(True) for t in (x, w, b)]
        y = _LinearFn.apply(xg, wg, bg, flags)
        y.backward(d.float().to(cuda))
        rep.add(f'y[{flags}]', relerr(y.detach().cpu().numpy(), ref.detach().numpy()))
        rep.add(f'dx[{flags}]', relerr(xg.grad.cpu().numpy(), xr.grad.numpy()))
        rep.add(f'dw[{flags}]', relerr(wg.grad.cpu().numpy(), wr.grad.numpy()))
        rep.add(f'db[{flags}]', relerr(bg.grad.cpu().numpy(), br.grad.numpy()))
    rep.finish()


@pytest.mark.parametrize('M,K,N', [(128, 64, 128), (300, 128, 1024), (1500, 768, 1024), (70, 37, 45), (4352, 1024, 768)])
def test_linear_tcgen05_forward(M, K, N, cuda):
    """bf16 tensor-core path (tcgen05, flag 0x100): with inputs already representable in bf16 the only
    error left is fp32 accumulation order, so the UMMA plumbing is checked at 1e-5."""
    from hop_b200.HOP import _LinearFn
    torch.manual_seed(M + K)
    x = torch.randn(M, K).bfloat16().double()
    w = (torch.randn(N, K) * 0.1).bfloat16().double()
    b = torch.randn(N, dtype=torch.float64)
    rep = Report(f'linear_tc_{M}_{K}_{N}', TOL_FP32)
    for flags in (0, 1, 2):
        xin = torch.relu(x) if flags & 1 else x
        ref = xin @ w.T + b
        if flags & 2:
            ref = torch.relu(ref)
        with torch.no_grad():
            y = _LinearFn.apply(x.float().to(cuda), w.float().to(cuda), b.float().to(cuda), flags | 0x100)
        rep.add(f'y[{flags}]', relerr(y.cpu().numpy(), ref.numpy()))
    rep.finish()


@pytest.mark.parametrize('M,K,N', [(128, 64, 128), (68, 128, 1024), (1500, 768, 1024), (70, 37, 45), (4352, 1024, 768),
                                   (68, 1024, 768), (20000, 64, 64)])
def test_linear_tcgen05_backward_exact(M, K, N, cuda):
    """dgrad / wgrad / bias-grad on tcgen05 with bf16-representable x, w, dy: only fp32 accumulation error remains."""
    from hop_b200.HOP import _LinearFn
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K).bfloat16().double()
    w = (torch.randn(N, K) * 0.1).bfloat16().double()
    b = torch.randn(N, dtype=torch.float64)
    dy = torch.randn(M, N).bfloat16().double()
    rep = Report(f'linear_tc_bwd_{M}_{K}_{N}', 2e-5)
    for flags in (0, 1):
        xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
        ((torch.relu(xr) if flags & 1 else xr) @ wr.T + br).backward(dy)
        xg, wg, bg = [t.float().to(cuda).requires_grad_(True) for t in (x, w, b)]
        _LinearFn.apply(xg, wg, bg, flags | 0x100).backward(dy.float().to(cuda))
        rep.add(f'dx[{flags}]', relerr(xg.grad.cpu().numpy(), xr.grad.numpy()))
        rep.add(f'dw[{flags}]', relerr(wg.grad.cpu().numpy(), wr.grad.numpy()))
        rep.add(f'db[{flags}]', relerr(bg.grad.cpu().numpy(), br.grad.numpy()))
    rep.finish()
