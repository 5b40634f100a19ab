"""GPU parity of the hand-written frozen-BERT path (hop_b200.bert: TMA GEMMs + csrc/bert.cu) against the Hugging Face
BertModel evaluated in float64 on the same device.  dtype-1 arithmetic: 2e-2 of each tensor's scale, forward and dX."""
import numpy as np
import pytest
import torch

from tests.util import TOL_BF16, Report, l2err, relerr

pytestmark = pytest.mark.gpu
npy = lambda t: t.detach().double().cpu().numpy()


def _bert(dev, layers, seed=0):
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    m = BertModel(BertConfig(num_hidden_layers=layers)).eval().to(dev)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


@pytest.mark.parametrize('B,S,layers', [(128, 34, 6), (3, 34, 2), (5, 17, 1)])
def test_bert_vs_float64(B, S, layers, cuda):
    from hop_b200 import bert as hbert
    m = _bert(cuda, layers)
    assert hbert.supported(m, S)
    g = torch.Generator(device='cpu').manual_seed(B + S)
    x = torch.randn(B, S, 768, generator=g).to(cuda)
    dy = torch.randn(B, S, 768, generator=g).to(cuda)
    xo = x.clone().requires_grad_(True)
    yo = hbert.run(m, xo)
    yo.backward(dy)
    import copy
    md = copy.deepcopy(m).double()
    xr = x.double().requires_grad_(True)
    yr = md(inputs_embeds=xr).last_hidden_state
    yr.backward(dy.double())
    rep = Report(f'bert_{B}_{S}_{layers}', TOL_BF16)
    rep.add('last_hidden_state', relerr(npy(yo), npy(yr)))
    rep.add('last_hidden_state(l2)', l2err(npy(yo), npy(yr)))
    rep.add('d inputs_embeds', relerr(npy(xo.grad), npy(xr.grad)))
    rep.add('d inputs_embeds(l2)', l2err(npy(xo.grad), npy(xr.grad)))
    with torch.no_grad():
        yn = hbert.run(m, x)                                   # inference path (GELU fused in the GEMM epilogue)
    rep.add('no-grad path', relerr(npy(yn), npy(yr)))
    rep.finish()


def test_bert_layer_norm_and_attention_ops(cuda):
    """The two small kernels alone, against float64 torch (fp32 statistics: 1e-5; bf16 outputs: 2^-8)."""
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(4)
    M, C = 300, 768
    x = torch.randn(M, C, generator=g).to(cuda); add = torch.randn(34, C, generator=g).to(cuda)
    gam = torch.rand(C, generator=g).to(cuda) + 0.5; bet = torch.randn(C, generator=g).to(cuda); dy = torch.randn(M, C, generator=g).to(cuda)
    y32 = torch.empty(M, C, device=cuda); y16 = torch.empty(M, C, device=cuda, dtype=torch.bfloat16); st = torch.empty(M, 2, device=cuda)
    _lib.check(L.hopk_ln_fwd(_lib.ptr(x), _lib.ptr(add), 34, _lib.ptr(gam), _lib.ptr(bet), 1e-12, _lib.ptr(y32), _lib.ptr(y16), _lib.ptr(st), M, C, _lib.stream_ptr()))
    d32 = torch.empty(M, C, device=cuda)
    _lib.check(L.hopk_ln_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(add), 34, _lib.ptr(gam), _lib.ptr(st), _lib.ptr(d32), None, M, C, _lib.stream_ptr()))
    xr = (x.double() + add.double()[torch.arange(M, device=cuda) % 34]).requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (C,), gam.double(), bet.double(), 1e-12)
    yr.backward(dy.double())
    rep = Report('bert_ops', 1e-5)
    rep.add('ln y32', relerr(npy(y32), npy(yr)))
    rep.add('ln y16', relerr(npy(y16), npy(yr)), tol=5e-3)
    rep.add('ln dx', relerr(npy(d32), npy(xr.grad)))
    rep.finish()


@pytest.mark.parametrize('B,S,H', [(7, 34, 12), (3, 50, 3), (1, 64, 1), (5, 17, 12), (128, 34, 12)])
def test_bert_attention_op(B, S, H, cuda):
    """hopk_bert_attn_fwd / _bwd (two (sample, head) pairs per CTA, tcgen05) against float64 torch attention on the same bf16
    inputs: probabilities 1e-5 (fp32 softmax of exact bf16 products), context / dqkv within the bf16 operand rounding."""
    from hop_b200 import _lib
    L = _lib.lib()
    D = 64
    g = torch.Generator(device='cpu').manual_seed(B * 100 + S)
    qkv = torch.randn(B * S, 3 * H * D, generator=g).bfloat16().to(cuda); dctx = torch.randn(B * S, H * D, generator=g).bfloat16().to(cuda)
    ctx = torch.full((B * S, H * D), float('nan'), device=cuda, dtype=torch.bfloat16); P = torch.empty(B * H, S, S, device=cuda)
    dqkv = torch.full_like(qkv, float('nan'))
    _lib.check(L.hopk_bert_attn_fwd(_lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(P), B, S, H, D, _lib.stream_ptr()))
    _lib.check(L.hopk_bert_attn_bwd(_lib.ptr(qkv), _lib.ptr(dctx), _lib.ptr(dqkv), B, S, H, D, _lib.stream_ptr()))
    qr = qkv.double().requires_grad_(True)
    q, k, v = [t.view(B, S, H, D).transpose(1, 2) for t in qr.split(H * D, dim=1)]
    pr = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    cr = (pr @ v).transpose(1, 2).reshape(B * S, H * D)
    cr.backward(dctx.double())
    rep = Report(f'bert_attn_{B}_{S}_{H}', 1e-2)
    rep.add('attn ctx', relerr(npy(ctx), npy(cr)))
    rep.add('attn P', relerr(npy(P), npy(pr.reshape(B * H, S, S))), tol=1e-5)
    rep.add('attn dqkv', relerr(npy(dqkv), npy(qr.grad)))
    rep.add('attn dqkv(l2)', l2err(npy(dqkv), npy(qr.grad)))
    rep.finish()
