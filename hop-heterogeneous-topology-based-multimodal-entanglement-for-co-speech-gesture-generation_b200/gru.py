"""The decoder GRU of HOP.Model on hand-written kernels (csrc/gru.cu through the C ABI of include/hopk.h).

``run(gru_module, x)`` computes what ``gru_module(x)[0]`` computes for the ``nn.GRU`` the reference builds at
model/HOP.py:166-167 (4 layers, bidirectional, batch_first, hidden 350, zero initial state) -- the parameters stay the
module's own (``state_dict`` keys ``gru.weight_ih_l0`` ... unchanged).  Arithmetic is the library's dtype-1 mode: bf16
tensor-core operands, fp32 accumulation, fp32 recurrent state; input projections / weight gradients on the TMA + tcgen05
GEMM, the recurrence as a cluster-resident persistent kernel.  No fallback: CPU tensors raise.
"""
import torch

from . import _lib, profiler
from ._lib import GruGrads, GruParams, GruShape, check, f32c, lib, ptr, stream_ptr


def _names(L):
    out = []
    for l in range(L):
        for suf in ('', '_reverse'):
            out += [f'weight_ih_l{l}{suf}', f'weight_hh_l{l}{suf}', f'bias_ih_l{l}{suf}', f'bias_hh_l{l}{suf}']
    return out


def _fill(struct, tensors, L):
    """tensors: flat list in the order of :func:`_names`."""
    i = 0
    for l in range(L):
        for d in range(2):
            for field in ('w_ih', 'w_hh', 'b_ih', 'b_hh'):
                t = tensors[i]
                getattr(struct, field)[l][d] = t.data_ptr() if t is not None else 0
                i += 1
    return struct


class _GruFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, L, *weights):
        x = f32c(x)
        B, T, I = x.shape
        ws_ = [f32c(w) for w in weights]
        shape = GruShape(B, T, I, H, L, 1 if any(ctx.needs_input_grad) else 0)       # 0: no-grad pass, nothing kept for backward
        l = lib()
        out = torch.empty((B, T, 2 * H), device=x.device, dtype=torch.float32)
        work = torch.empty(l.hopk_gru_workspace_bytes(shape), device=x.device, dtype=torch.uint8)
        with profiler.span('gru_fwd'):
            check(l.hopk_gru_forward(shape, _fill(GruParams(), ws_, L), ptr(x), ptr(out), ptr(work), stream_ptr()))
        ctx.save_for_backward(work, *ws_)
        ctx.shape = shape
        return out

    @staticmethod
    def backward(ctx, dout):
        work, *ws_ = ctx.saved_tensors
        shape = ctx.shape
        if not shape.save:
            raise RuntimeError('hop_b200.gru: backward through a forward that ran without gradient bookkeeping')
        l = lib()
        dout = f32c(dout)
        dev = dout.device
        grads = [torch.empty_like(w) for w in ws_]
        dx = torch.empty((shape.B, shape.T, shape.I), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        scratch = torch.empty(l.hopk_gru_scratch_bytes(shape), device=dev, dtype=torch.uint8)
        with profiler.span('gru_bwd'):
            check(l.hopk_gru_backward(shape, _fill(GruParams(), ws_, shape.L), ptr(dout), ptr(work), ptr(scratch),
                                      _fill(GruGrads(), grads, shape.L), ptr(dx), stream_ptr()))
        return (dx, None, None, *grads)


def supported(gru):
    return (isinstance(gru, torch.nn.GRU) and gru.bidirectional and gru.batch_first and gru.bias and gru.dropout == 0
            and gru.hidden_size <= 352 and gru.num_layers <= _lib.GRU_MAX_LAYERS and getattr(gru, 'proj_size', 0) == 0)


def run(gru, x):
    """Output sequence (B, T, 2*hidden) of ``gru`` on ``x`` (B, T, input) from a zero initial state."""
    if not supported(gru):
        raise NotImplementedError('hop_b200.gru: needs a bidirectional, batch_first nn.GRU with biases, dropout 0, hidden <= 352')
    if not x.is_cuda:
        raise RuntimeError('hop_b200.gru needs CUDA tensors (no CPU fallback)')
    weights = [getattr(gru, n) for n in _names(gru.num_layers)]
    return _GruFn.apply(x, gru.hidden_size, gru.num_layers, *weights)
