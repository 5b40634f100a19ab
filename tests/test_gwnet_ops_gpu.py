"""GPU parity of the per-operator C-ABI entry points of the Graph-WaveNet block (adp_softmax, gated_tcn,
gcn_diffuse_mlp_res_bnstat, bn_finalize) through ctypes, against float64 torch restatements of reference model/gwnet.py.
fp32 mode: 1e-5; bf16 mode (bf16-representable operands, so that the kernels' own arithmetic is what is measured): 2e-2."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import Report, relerr

pytestmark = pytest.mark.gpu
npy = lambda t: t.detach().double().cpu().numpy()


def rows(t):
    """NCHW (B, C, V, T) -> rows layout (B, T, V, C) contiguous."""
    return t.permute(0, 3, 2, 1).contiguous()


def unrows(t):
    return t.permute(0, 3, 2, 1)


def test_adp_softmax_fwd_bwd(cuda):
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(0)
    for V in (9, 42):
        e1 = torch.randn(V, 10, generator=g).to(cuda); e2 = torch.randn(10, V, generator=g).to(cuda); dA = torch.randn(V, V, generator=g).to(cuda)
        A5 = torch.empty(5, V, V, device=cuda); de1 = torch.empty_like(e1); de2 = torch.empty_like(e2)
        _lib.check(L.hopk_adp_softmax_fwd(_lib.ptr(e1), _lib.ptr(e2), V, 10, _lib.ptr(A5), _lib.stream_ptr()))
        _lib.check(L.hopk_adp_softmax_bwd(_lib.ptr(e1), _lib.ptr(e2), _lib.ptr(A5), _lib.ptr(dA), V, 10, _lib.ptr(de1), _lib.ptr(de2), _lib.stream_ptr()))
        a, b = e1.double().requires_grad_(True), e2.double().requires_grad_(True)
        A = torch.softmax(torch.relu(a @ b), dim=1)                 # gwnet.py:163
        A.backward(dA.double())
        rep = Report(f'op_adp_V{V}', 1e-5)
        rep.add('A', relerr(npy(A5[0]), npy(A))); rep.add('A^2', relerr(npy(A5[1]), npy(A @ A))); rep.add('A^T', relerr(npy(A5[2]), npy(A.t())))
        rep.add('dE1', relerr(npy(de1), npy(a.grad))); rep.add('dE2', relerr(npy(de2), npy(b.grad)))
        rep.finish()


@pytest.mark.parametrize('dtype', [0, 1])
@pytest.mark.parametrize('B,V,Ti,d,Cc', [(128, 9, 16, 1, 64), (8, 43, 13, 2, 128), (3, 10, 7, 2, 32)])
def test_gated_tcn_fwd_bwd(B, V, Ti, d, Cc, dtype, cuda):
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(B + V + Cc)
    q = (lambda t: t.bfloat16().float()) if dtype else (lambda t: t)
    x = q(torch.randn(B, Cc, V, Ti, generator=g)).to(cuda)
    wf = q(torch.randn(Cc, Cc, 1, 2, generator=g) / (2 * Cc) ** 0.5).to(cuda); wg = q(torch.randn(Cc, Cc, 1, 2, generator=g) / (2 * Cc) ** 0.5).to(cuda)
    bf = torch.randn(Cc, generator=g).to(cuda) * 0.1; bg = torch.randn(Cc, generator=g).to(cuda) * 0.1
    To = Ti - d
    dy = torch.randn(B, Cc, V, To, generator=g).to(cuda)
    ss = torch.cat([torch.ones(Cc), torch.zeros(Cc)]).to(cuda)
    xr = rows(x)
    tf = torch.empty(B, To, V, Cc, device=cuda); sg = torch.empty_like(tf); y = torch.empty_like(tf)
    _lib.check(L.hopk_gated_tcn_fwd(_lib.ptr(xr), _lib.ptr(ss), _lib.ptr(wf), _lib.ptr(bf), _lib.ptr(wg), _lib.ptr(bg), B, V, Ti, d, Cc, dtype,
                                    _lib.ptr(tf), _lib.ptr(sg), _lib.ptr(y), _lib.stream_ptr()))
    scratch = torch.empty(L.hopk_gated_tcn_bwd_scratch_bytes(B, V, Ti, d, Cc), device=cuda, dtype=torch.uint8)
    dx = torch.empty(B, Ti, V, Cc, device=cuda); dwf = torch.empty_like(wf); dwg = torch.empty_like(wg); dbf = torch.empty_like(bf); dbg = torch.empty_like(bg)
    _lib.check(L.hopk_gated_tcn_bwd(_lib.ptr(xr), _lib.ptr(ss), _lib.ptr(wf), _lib.ptr(wg), _lib.ptr(tf), _lib.ptr(sg), _lib.ptr(rows(dy)), B, V, Ti,
                                    d, Cc, dtype, _lib.ptr(scratch), _lib.ptr(dx), _lib.ptr(dwf), _lib.ptr(dbf), _lib.ptr(dwg), _lib.ptr(dbg),
                                    _lib.stream_ptr()))
    X, WF, WG, BF, BG = [t.double().requires_grad_(True) for t in (x, wf, wg, bf, bg)]
    f_ = torch.tanh(F.conv2d(X, WF, BF, dilation=(1, d))); g_ = torch.sigmoid(F.conv2d(X, WG, BG, dilation=(1, d)))      # gwnet.py:186-200
    (f_ * g_).backward(dy.double())
    tol = 2e-2 if dtype else 1e-5
    rep = Report(f'op_gated_tcn_{B}_{V}_{Ti}_{d}_{Cc}_{dtype}', tol)
    rep.add('tanh f', relerr(npy(unrows(tf)), npy(f_))); rep.add('sigmoid g', relerr(npy(unrows(sg)), npy(g_))); rep.add('y', relerr(npy(unrows(y)), npy(f_ * g_)))
    rep.add('dx', relerr(npy(unrows(dx)), npy(X.grad)))
    for n, a, b in (('dWf', dwf, WF), ('dWg', dwg, WG), ('dbf', dbf, BF), ('dbg', dbg, BG)):
        rep.add(n, relerr(npy(a), npy(b.grad)))
    rep.finish()


@pytest.mark.parametrize('dtype', [0, 1])
@pytest.mark.parametrize('B,V,Ti,d,Cc', [(128, 9, 16, 1, 64), (8, 43, 13, 2, 128), (3, 10, 7, 2, 32)])
def test_gcn_diffuse_mlp_res_bnstat_fwd_bwd(B, V, Ti, d, Cc, dtype, cuda):
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(B + 2 * V + Cc)
    q = (lambda t: t.bfloat16().float()) if dtype else (lambda t: t)
    To = Ti - d
    e1 = torch.randn(V, 10, generator=g).to(cuda); e2 = torch.randn(10, V, generator=g).to(cuda)
    A5 = torch.empty(5, V, V, device=cuda)
    _lib.check(L.hopk_adp_softmax_fwd(_lib.ptr(e1), _lib.ptr(e2), V, 10, _lib.ptr(A5), _lib.stream_ptr()))
    yv = q(torch.randn(B, Cc, V, To, generator=g)).to(cuda); xres = torch.randn(B, Cc, V, Ti, generator=g).to(cuda)
    wm = q(torch.randn(Cc, 3 * Cc, 1, 1, generator=g) / (3 * Cc) ** 0.5).to(cuda); bm = torch.randn(Cc, generator=g).to(cuda) * 0.1
    du = q(torch.randn(B, Cc, V, To, generator=g)).to(cuda)
    ss = torch.cat([torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1]).to(cuda)
    yr = rows(yv)
    x1 = torch.empty_like(yr); x2 = torch.empty_like(yr); u = torch.empty_like(yr); stats = torch.empty(2 * Cc, device=cuda, dtype=torch.float64)
    _lib.check(L.hopk_gcn_diffuse_mlp_res_bnstat_fwd(_lib.ptr(yr), _lib.ptr(A5), _lib.ptr(rows(xres)), _lib.ptr(ss), _lib.ptr(wm), _lib.ptr(bm), B, V,
                                                      Ti, d, Cc, dtype, _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(u), _lib.ptr(stats), _lib.stream_ptr()))
    scratch = torch.empty(L.hopk_gcn_scratch_bytes(B, V, To, Cc), device=cuda, dtype=torch.uint8)
    dy = torch.empty_like(yr); dwm = torch.empty_like(wm); dbm = torch.empty_like(bm); dA = torch.empty(V, V, device=cuda)
    _lib.check(L.hopk_gcn_diffuse_mlp_res_bnstat_bwd(_lib.ptr(rows(du)), _lib.ptr(yr), _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(A5), _lib.ptr(wm), B, V, To,
                                                      Cc, dtype, _lib.ptr(scratch), _lib.ptr(dy), _lib.ptr(dwm), _lib.ptr(dbm), _lib.ptr(dA),
                                                      _lib.stream_ptr()))
    A = A5[0].double().requires_grad_(True)
    Y, WM, BM = yv.double().requires_grad_(True), wm.double().requires_grad_(True), bm.double().requires_grad_(True)
    X1 = torch.einsum('ncvl,vw->ncwl', Y, A); X2 = torch.einsum('ncvl,vw->ncwl', X1, A)                              # gwnet.py:12-14, 35-41
    res = xres.double()[..., d:] * ss[:Cc].double()[None, :, None, None] + ss[Cc:].double()[None, :, None, None]
    U = F.conv2d(torch.cat([Y, X1, X2], 1), WM, BM) + res                                                             # gwnet.py:43-45, 233
    U.backward(du.double())
    tol = 2e-2 if dtype else 1e-5
    rep = Report(f'op_gcn_{B}_{V}_{Ti}_{d}_{Cc}_{dtype}', tol)
    rep.add('x1', relerr(npy(unrows(x1)), npy(X1)), tol=1e-5); rep.add('x2', relerr(npy(unrows(x2)), npy(X2)), tol=1e-5)
    rep.add('u', relerr(npy(unrows(u)), npy(U)))
    rep.add('sum u', relerr(npy(stats[:Cc]), npy(U.sum((0, 2, 3)))), tol=max(tol, 1e-4)); rep.add('sum u^2', relerr(npy(stats[Cc:]), npy((U * U).sum((0, 2, 3)))))
    rep.add('dy', relerr(npy(unrows(dy)), npy(Y.grad))); rep.add('dWm', relerr(npy(dwm), npy(WM.grad))); rep.add('dbm', relerr(npy(dbm), npy(BM.grad)))
    rep.add('dA', relerr(npy(dA), npy(A.grad)))
    rep.finish()


def test_bn_finalize(cuda):
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(2)
    Cc, n = 64, 1000
    u = torch.randn(n, Cc, generator=g, dtype=torch.float64) * 2 + 0.3
    stats = torch.cat([u.sum(0), (u * u).sum(0)]).to(cuda)
    gam = (torch.rand(Cc, generator=g) + 0.5).to(cuda); bet = torch.randn(Cc, generator=g).to(cuda)
    rm = torch.randn(Cc, generator=g).to(cuda); rv = (torch.rand(Cc, generator=g) + 0.5).to(cuda); nbt = torch.tensor(5, device=cuda)
    rm0, rv0 = rm.clone(), rv.clone()
    mr = torch.empty(2 * Cc, device=cuda); ssn = torch.empty(2 * Cc, device=cuda)
    _lib.check(L.hopk_bn_finalize(_lib.ptr(stats), float(n), _lib.ptr(gam), _lib.ptr(bet), _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(nbt), _lib.ptr(mr),
                                  _lib.ptr(ssn), Cc, 1, 0.1, 1e-5, _lib.stream_ptr()))
    mean, var = u.mean(0), u.var(0, unbiased=False)
    rstd = 1 / torch.sqrt(var + 1e-5)
    rep = Report('op_bn_finalize', 1e-5)
    rep.add('mean', relerr(npy(mr[:Cc]), npy(mean))); rep.add('rstd', relerr(npy(mr[Cc:]), npy(rstd)))
    rep.add('scale', relerr(npy(ssn[:Cc]), npy(gam.cpu().double() * rstd))); rep.add('shift', relerr(npy(ssn[Cc:]), npy(bet.cpu().double() - mean * gam.cpu().double() * rstd)))
    rep.add('running_mean', relerr(npy(rm), npy(0.9 * rm0.cpu().double() + 0.1 * mean)))
    rep.add('running_var', relerr(npy(rv), npy(0.9 * rv0.cpu().double() + 0.1 * u.var(0, unbiased=True))))
    assert int(nbt) == 6
    rep.finish()
