"""The frozen BERT encoder of HOP.Model on hand-written kernels.

``run(bert, inputs_embeds)`` computes ``bert(inputs_embeds=inputs_embeds).last_hidden_state`` (reference model/HOP.py:204) for
the Hugging Face ``BertModel`` the reference loads at run_ted.py:176-196: embeddings (+ position + token-type-0, LayerNorm),
N x [self-attention, output projection + residual + LayerNorm, GELU feed-forward + residual + LayerNorm].  The encoder is
frozen (HOP.py:90-91), so backward produces only the gradient with respect to ``inputs_embeds``.

dtype-1 arithmetic: every projection is a bf16 tensor-core GEMM on csrc/gemm_tma.cu (bias / residual / GELU-derivative fused in
the epilogue, dX products read the same weight matrix as an MN-major operand), layer norm / softmax / the residual stream in
fp32 (csrc/bert.cu).  Weights are converted to bf16 once and cached (they are frozen).  Dropout inside the encoder must be
inactive (``bert.eval()``, what ``from_pretrained`` returns); otherwise ``supported`` is False and the caller keeps the stock
module.  No fallback inside: CPU tensors raise.
"""
import torch

from . import _lib, profiler
from ._lib import check, f32c, lib, ptr, stream_ptr
from .dense import gemm


class _Weights:
    """bf16 / fp32 device copies of the frozen encoder weights in the layouts the kernels read."""

    def __init__(self, bert, device):
        cfg = bert.config
        self.hidden, self.heads, self.layers = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
        self.inter, self.eps = cfg.intermediate_size, float(cfg.layer_norm_eps)
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        h = lambda t: t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
        emb = bert.embeddings
        # position + token-type-0 embedding rows: what BertEmbeddings adds to inputs_embeds before its LayerNorm
        self.table = f(emb.position_embeddings.weight + emb.token_type_embeddings.weight[0][None, :])
        self.g0, self.b0 = f(emb.LayerNorm.weight), f(emb.LayerNorm.bias)
        self.L = []
        for lyr in bert.encoder.layer:
            att, so = lyr.attention.self, lyr.attention.output
            self.L.append(dict(
                wqkv=h(torch.cat([att.query.weight, att.key.weight, att.value.weight], 0)),
                bqkv=f(torch.cat([att.query.bias, att.key.bias, att.value.bias], 0)),
                wo=h(so.dense.weight), bo=f(so.dense.bias), g1=f(so.LayerNorm.weight), b1=f(so.LayerNorm.bias),
                w1=h(lyr.intermediate.dense.weight), c1=f(lyr.intermediate.dense.bias),
                w2=h(lyr.output.dense.weight), c2=f(lyr.output.dense.bias), g2=f(lyr.output.LayerNorm.weight), b2=f(lyr.output.LayerNorm.bias)))


_CACHE = {}


def weights(bert, device):
    key = (id(bert), str(device))
    w = _CACHE.get(key)
    if w is None:
        w = _CACHE[key] = _Weights(bert, device)
    return w


def invalidate(bert=None):
    """Forget cached weight copies (after ``load_state_dict`` of the frozen encoder)."""
    if bert is None:
        _CACHE.clear()
    else:
        for k in [k for k in _CACHE if k[0] == id(bert)]:
            del _CACHE[k]


def supported(bert, seq_len=34):
    cfg = getattr(bert, 'config', None)
    if cfg is None or type(bert).__name__ != 'BertModel':
        return False
    drop = bert.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)
    return (not drop and cfg.hidden_act == 'gelu' and cfg.hidden_size % 128 == 0 and cfg.hidden_size <= 1024 and
            cfg.hidden_size // cfg.num_attention_heads == 64 and seq_len <= 64 and not getattr(cfg, 'is_decoder', False) and
            getattr(cfg, 'position_embedding_type', 'absolute') in (None, 'absolute') and
            all(p.requires_grad is False for p in bert.parameters()))


def _ln_fwd(x, add, period, g, b, eps, M, C, want32=True):
    dev = x.device
    y32 = torch.empty((M, C), device=dev, dtype=torch.float32) if want32 else None
    y16 = torch.empty((M, C), device=dev, dtype=torch.bfloat16)
    stat = torch.empty((M, 2), device=dev, dtype=torch.float32)
    check(lib().hopk_ln_fwd(ptr(x), ptr(add), period, ptr(g), ptr(b), eps, ptr(y32), ptr(y16), ptr(stat), M, C, stream_ptr()))
    return y32, y16, stat


def _ln_bwd(dy, x, add, period, g, stat, M, C, want16=True):
    dev = dy.device
    d32 = torch.empty((M, C), device=dev, dtype=torch.float32)
    d16 = torch.empty((M, C), device=dev, dtype=torch.bfloat16) if want16 else None
    check(lib().hopk_ln_bwd(ptr(dy), ptr(x), ptr(add), period, ptr(g), ptr(stat), ptr(d32), ptr(d16), M, C, stream_ptr()))
    return d32, d16


class _BertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        B, S, C = x.shape
        M, H, I = B * S, W.heads, W.inter
        l = lib()
        dev = x.device
        save = bool(ctx.needs_input_grad[0])
        x2 = f32c(x).reshape(M, C)
        keep = []
        with profiler.span('bert_fwd'):
            h32, h16, st0 = _ln_fwd(x2, W.table, S, W.g0, W.b0, W.eps, M, C)
            for p in W.L:
                qkv = gemm(h16, p['wqkv'], M, 3 * C, C, bias=p['bqkv'], out_bf16=True)
                cx = torch.empty((M, C), device=dev, dtype=torch.bfloat16)
                check(l.hopk_bert_attn_fwd(ptr(qkv), ptr(cx), None, B, S, H, C // H, stream_ptr()))
                a = gemm(cx, p['wo'], M, C, C, bias=p['bo'], addend=h32)                       # + residual
                y32, y16, st1 = _ln_fwd(a, None, 0, p['g1'], p['b1'], W.eps, M, C)
                if save:
                    pre = gemm(y16, p['w1'], M, I, C, bias=p['c1'], out_bf16=True)
                    hh = torch.empty_like(pre)
                    check(l.hopk_gelu_bf16(ptr(pre), ptr(hh), pre.numel(), stream_ptr()))
                else:
                    pre, hh = None, gemm(y16, p['w1'], M, I, C, bias=p['c1'], out_bf16=True, act=_lib.GEMM_GELU)
                f = gemm(hh, p['w2'], M, C, I, bias=p['c2'], addend=y32)                       # + residual
                h32, h16, st2 = _ln_fwd(f, None, 0, p['g2'], p['b2'], W.eps, M, C)
                if save:
                    keep += [qkv, a, st1, pre, f, st2]
        if save:
            ctx.save_for_backward(x2, st0, *keep)
        ctx.W, ctx.dims = W, (B, S, C)
        return h32.view(B, S, C)

    @staticmethod
    def backward(ctx, dout):
        W = ctx.W
        B, S, C = ctx.dims
        M, H, I = B * S, W.heads, W.inter
        l = lib()
        x2, st0, *keep = ctx.saved_tensors
        d = f32c(dout).reshape(M, C)
        with profiler.span('bert_bwd'):
            for li in range(len(W.L) - 1, -1, -1):
                p = W.L[li]
                qkv, a, st1, pre, f, st2 = keep[6 * li:6 * li + 6]
                d2_32, d2_16 = _ln_bwd(d, f, None, 0, p['g2'], st2, M, C)
                # dY @ W2 through GELU': the saved pre-activation rides the epilogue
                dpre = gemm(d2_16, p['w2'], M, I, C, b_mn=True, mask=pre, mask_gelu=True, out_bf16=True)
                dy1 = gemm(dpre, p['w1'], M, C, I, b_mn=True, addend=d2_32)
                d1_32, d1_16 = _ln_bwd(dy1, a, None, 0, p['g1'], st1, M, C)
                dcx = gemm(d1_16, p['wo'], M, C, C, b_mn=True, out_bf16=True)
                dqkv = torch.empty_like(qkv)
                check(l.hopk_bert_attn_bwd(ptr(qkv), ptr(dcx), ptr(dqkv), B, S, H, C // H, stream_ptr()))
                d = gemm(dqkv, p['wqkv'], M, C, 3 * C, b_mn=True, addend=d1_32)
            dx, _ = _ln_bwd(d, x2, W.table, S, W.g0, st0, M, C, want16=False)
        return dx.view(B, S, C), None


def run(bert, inputs_embeds):
    """``bert(inputs_embeds=inputs_embeds).last_hidden_state`` for a frozen, dropout-free BertModel."""
    if not inputs_embeds.is_cuda:
        raise RuntimeError('hop_b200.bert needs CUDA tensors (no CPU fallback)')
    if not supported(bert, inputs_embeds.shape[1]):
        raise NotImplementedError('hop_b200.bert: needs a frozen BertModel without active dropout, GELU, head dim 64, sequence <= 64')
    return _BertFn.apply(inputs_embeds, weights(bert, inputs_embeds.device))
