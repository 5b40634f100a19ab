"""Dense layers of the HOP step on the TMA + tcgen05 bf16 GEMM (csrc/gemm_tma.cu, hopk_gemm_bf16) -- dtype-1 arithmetic:
bf16 operands, fp32 accumulation, fp32 results.

  linear(x, w, b, relu_in)   nn.Linear (+ the ReLU the reference applies in front of out_projection, HOP.py:284-285):
                             the reprogramming Q/K/V/O projections (HOP.py:276-285) and align_layer (HOP.py:202-203)
  source(w_map, b_map, we)   text prototypes  W_map @ WE + b  == mapping_layer(WE^T)^T  (HOP.py:200)
  beat_rows(audio, seed, beat_mlp, J)
                             in_audio.unfold(1, 3400, 2191) -> Linear -> LeakyReLU(0.2) -> Linear (HOP.py:130-134, 210-212) run ONCE
                             on the 16 windows (the reference repeats them J times), gathered with idx = (t*J+j) % 16
                             and concatenated with the seed bones straight into Graph-WaveNet's (B, 16, J, 173) rows buffer
                             (HOP.py:214-217)

Forward: x @ W^T is the GEMM's default "K-major" form.  Backward: dX = dY @ W reads W as an MN-major B operand, dW = dY^T @ X
reads both operands MN-major (split-K, vector atomics), so no transposed copies exist; bias gradients are fp32 column sums.
No fallback: CPU tensors raise in ``_lib.ptr``.
"""
import torch

from . import _lib, profiler
from ._lib import check, f32c, lib, ptr, stream_ptr

_SMS = 148
KEEP = False          # tests / tools: keep references to intermediate activations in LAST (gate patterns for pinned oracles)
LAST = {}


def _pad8(n):
    return (n + 7) // 8 * 8


def cast_bf16(x2, relu=False):
    """fp32 (rows, cols) -> bf16 (rows, pad8(cols)), zero padded."""
    rows, cols = x2.shape
    out = torch.empty((rows, _pad8(cols)), device=x2.device, dtype=torch.bfloat16)
    check(lib().hopk_cast_bf16(ptr(x2), ptr(out), rows, cols, x2.stride(0), out.shape[1], out.stride(0), 1 if relu else 0, stream_ptr()))
    return out


def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, bias=None, addend=None, mask=None, mask_bf16=False, mask_gelu=False, out=None,
         out_bf16=False, act=0, slope=0.0, splits=1, bias_row=False, accumulate=False, ldc=None):
    """C[M, N] = A . B^T-style contraction over K (see include/hopk.h: hopk_gemm_bf16)."""
    ldc = ldc or (out.stride(0) if out is not None else (_pad8(N) if out_bf16 else N))
    if out is None:
        out = torch.empty((M, ldc), device=A.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    flags = ((_lib.GEMM_A_MN if a_mn else 0) | (_lib.GEMM_B_MN if b_mn else 0) | (_lib.GEMM_OUT_BF16 if out_bf16 else 0) | act |
             (_lib.GEMM_BIAS_ROW if bias_row else 0) | (_lib.GEMM_MASK_BF16 if mask_bf16 else 0) | (_lib.GEMM_MASK_GELU if mask_gelu else 0) |
             (_lib.GEMM_ACCUMULATE if accumulate else 0))
    with profiler.span('gemm_tma', 2.0 * M * N * K, fine=True):              # the step's dominant kernel: timed per launch for the roofline
        check(lib().hopk_gemm_bf16(ptr(A), ptr(B), ptr(out), ptr(bias), ptr(addend), ptr(mask), M, N, K, A.stride(0), B.stride(0), ldc,
                                   flags, float(slope), int(splits), stream_ptr()))
    return out


def _wgrad_splits(M, N, K):
    """split-K factor of a weight-gradient GEMM (small output, long contraction): about two tiles per SM."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    return max(1, min((2 * _SMS + tiles - 1) // tiles, K // 256 or 1, 16))


def colsum(x2):
    out = torch.empty(x2.shape[1], device=x2.device, dtype=torch.float32)
    check(lib().hopk_colsum(ptr(x2), ptr(out), x2.shape[0], x2.shape[1], x2.stride(0), 1 if x2.dtype == torch.bfloat16 else 0, stream_ptr()))
    return out


# ------------------------------------------------------------------------------------ nn.Linear
class _LinearTmaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu_in):
        shp = x.shape
        x2 = f32c(x).reshape(-1, shp[-1])
        w = f32c(w)
        M, K = x2.shape
        N = w.shape[0]
        with profiler.span('linear_fwd'):
            xb = cast_bf16(x2, relu=relu_in)
            wb = cast_bf16(w)
            y = gemm(xb, wb, M, N, K, bias=b)
        ctx.save_for_backward(xb, wb, x2 if relu_in else None)
        ctx.shp, ctx.relu_in, ctx.has_bias, ctx.dims = shp, relu_in, b is not None, (M, N, K)
        return y.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xb, wb, x2 = ctx.saved_tensors
        M, N, K = ctx.dims
        dy2 = f32c(dy).reshape(M, N)
        dx = dw = db = None
        with profiler.span('linear_bwd'):
            dyb = cast_bf16(dy2)
            if ctx.needs_input_grad[0]:                   # dX = dY @ W  (W is [N][K] = [contraction][out]: MN-major B)
                dx = gemm(dyb, wb, M, K, N, b_mn=True, mask=x2 if ctx.relu_in else None).view(ctx.shp)
            if ctx.needs_input_grad[1]:                   # dW = dY^T @ X  (both MN-major), split-K
                dw = gemm(dyb, xb, N, K, M, a_mn=True, b_mn=True, splits=_wgrad_splits(N, K, M))
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = colsum(dy2)
        return dx, dw, db, None


def linear(x, w, b, relu_in=False):
    return _LinearTmaFn.apply(x, w, b, relu_in)


# ------------------------------------------------------------------------------------ text prototypes
class _SourceTmaFn(torch.autograd.Function):
    """source = W_map @ WE + b[:, None] (HOP.py:200).  ``we_b`` is the cached bf16 copy of the frozen word embeddings.
    With data parallelism the small upstream gradient is all-reduced through ``reducer`` before dW_map is formed locally
    (see hop_b200.dp): ``reducer(t)`` reduces in stream order; a reducer with a ``defer(t, finish)`` method
    takes the tensor, reduces it asynchronously and calls ``finish`` itself once backward has been issued."""

    @staticmethod
    def forward(ctx, w_map, b_map, we_b, reducer):
        S, Vc = w_map.shape
        D = we_b.shape[1]
        with profiler.span('mapping_fwd'):
            wb = cast_bf16(f32c(w_map))
            src = gemm(wb, we_b, S, D, Vc, b_mn=True, bias=f32c(b_map), bias_row=True, splits=8)
        ctx.save_for_backward(we_b)
        ctx.reducer, ctx.dims = reducer, (S, Vc, D)
        return src

    @staticmethod
    def backward(ctx, dsrc):
        (we_b,) = ctx.saved_tensors
        S, Vc, D = ctx.dims
        dsrc = f32c(dsrc)

        def finish(d):                                   # d: the (rank-averaged) upstream gradient
            with profiler.span('mapping_bwd'):
                db16 = cast_bf16(d)
                dw = gemm(db16, we_b, S, Vc, D)          # dW_map[s][v] = sum_d dsrc[s][d] WE[v][d]: both K-major
            return dw, d.sum(1)

        red = ctx.reducer
        if red is not None and hasattr(red, 'defer'):
            # data parallelism on CUDA: the all-reduce of dSource travels on the communication stream under the rest of
            # backward; DataParallel.backward() finishes the two parameter gradients after it (hop_b200.dp)
            red.defer(dsrc.clone(), finish)
            return None, None, None, None
        if red is not None:
            dsrc = red(dsrc.clone())
        dw, db = finish(dsrc)
        return dw, db, None, None


def source(w_map, b_map, we_b, reducer=None):
    return _SourceTmaFn.apply(w_map, b_map, we_b, reducer)


# ------------------------------------------------------------------------------------ beat features -> gwnet rows
WIN, HOP_, NWIN = 3400, 2191, 16


class _BeatRowsFn(torch.autograd.Function):
    """(audio (B, 36267), seed bones (B, 16, 3J), beat MLP weights) -> Graph-WaveNet input rows (B, 16, J, 173)."""

    @staticmethod
    def forward(ctx, audio, seed, w1, b1, w2, b2, J):
        audio, seed = f32c(audio), f32c(seed)
        B = audio.shape[0]
        F1, F2 = w1.shape[0], w2.shape[0]                # 1700, 170
        M = B * NWIN
        dev = audio.device
        with profiler.span('beat_fwd'):
            win = torch.empty((M, _pad8(WIN)), device=dev, dtype=torch.bfloat16)
            check(lib().hopk_unfold_bf16(ptr(audio), ptr(win), B, NWIN, WIN, HOP_, audio.stride(0), win.stride(0), stream_ptr()))
            w1b, w2b = cast_bf16(f32c(w1)), cast_bf16(f32c(w2))
            h1 = gemm(win, w1b, M, F1, WIN, bias=f32c(b1), act=_lib.GEMM_LEAKY, slope=0.2, out_bf16=True)       # (M, 1704) bf16
            feat = gemm(h1, w2b, M, F2, F1, bias=f32c(b2))                                                       # (M, 170) fp32
            rows = torch.empty((B, NWIN, J, 3 + F2), device=dev, dtype=torch.float32)
            check(lib().hopk_beat_rows_fwd(ptr(feat), ptr(seed), ptr(rows), B, J, F2, stream_ptr()))
        ctx.save_for_backward(win, w1b, w2b, h1)
        if KEEP:
            LAST['beat_h1'] = h1[:, :F1]
        ctx.dims = (B, J, F1, F2)
        return rows

    @staticmethod
    def backward(ctx, drows):
        win, w1b, w2b, h1 = ctx.saved_tensors
        B, J, F1, F2 = ctx.dims
        M = B * NWIN
        drows = f32c(drows)
        with profiler.span('beat_bwd'):
            dfeat = torch.empty((M, _pad8(F2)), device=drows.device, dtype=torch.bfloat16)
            db2 = torch.empty(F2, device=drows.device, dtype=torch.float32)
            check(lib().hopk_beat_rows_bwd(ptr(drows), ptr(dfeat), ptr(db2), B, J, F2, dfeat.stride(0), stream_ptr()))
            dw2 = gemm(dfeat, h1, F2, F1, M, a_mn=True, b_mn=True, splits=_wgrad_splits(F2, F1, M))
            # dh1 = (dfeat @ W2) * LeakyReLU'(h1): the mask is the saved bf16 activation (sign-preserving)
            dh1 = gemm(dfeat, w2b, M, F1, F2, b_mn=True, mask=h1, mask_bf16=True, slope=0.2, out_bf16=True, ldc=h1.stride(0))
            dw1 = gemm(dh1, win, F1, WIN, M, a_mn=True, b_mn=True, splits=_wgrad_splits(F1, WIN, M))
            db1 = colsum(dh1[:, :F1])
        return None, None, dw1, db1, dw2, db2, None


def beat_rows(audio, seed, beat_mlp, J):
    """``beat_mlp`` = the reference's nn.Sequential(Linear(3400, 1700), LeakyReLU(0.2), Linear(1700, 170))."""
    l1, l2 = beat_mlp[0], beat_mlp[2]
    return _BeatRowsFn.apply(audio, seed, l1.weight, l1.bias, l2.weight, l2.bias, J)
