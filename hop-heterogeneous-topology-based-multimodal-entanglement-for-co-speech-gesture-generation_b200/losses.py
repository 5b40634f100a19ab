"""The generator's loss terms of the training step (reference train_eval/train_llm.py:46-79) as one kernel forward and one
backward (csrc/glue.cu: hopk_step_losses_fwd / _bwd) instead of ~65 elementwise / reduction launches:

    huber   = F.smooth_l1_loss(out / 0.1, target / 0.1) * 0.1
    pose_l1 = (F.smooth_l1_loss(out / 0.05, out_rand.detach() / 0.05, reduction='none') * 0.05).sum(1).sum(1)
    z_l1    = F.l1_loss(z_context.detach(), z_rand.detach(), reduction='none').mean(1)
    div_reg = torch.clamp(-(pose_l1 / (z_l1 + 1e-5)), min=-1000).mean()
    kld     = -0.5 * torch.mean(1 + logvar - mu ** 2 - logvar.exp())
    loss    = w_reg * huber + w_div * div_reg + w_kld * kld

``step_losses`` returns ``(loss, vals)`` with ``vals = [loss, huber, div_reg, kld]`` detached (for reporting); gradients flow
to ``out``, ``mu`` and ``logvar`` exactly as in the expressions above.  fp32 arithmetic; CUDA only (no fallback here: the
caller keeps the torch expressions on CPU).
"""
import torch

from ._lib import check, f32c, lib, ptr, stream_ptr

_TICKETS = {}


def _ticket(device):
    key = torch.device(device).index
    if key not in _TICKETS:
        _TICKETS[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return _TICKETS[key]


class _StepLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, tgt, rnd, zc, zr, mu, logvar, w_reg, w_div, w_kld):
        B = out.shape[0]
        out2, tgt2 = f32c(out).reshape(B, -1), f32c(tgt).reshape(B, -1)
        TP = out2.shape[1]
        rnd2 = f32c(rnd).reshape(B, -1) if rnd is not None else None
        zc2 = f32c(zc).reshape(B, -1) if zc is not None else None
        zr2 = f32c(zr).reshape(B, -1) if zr is not None else None
        mu2 = f32c(mu).reshape(B, -1) if mu is not None else None
        lv2 = f32c(logvar).reshape(B, -1) if logvar is not None else None
        Z = (zc2 if zc2 is not None else mu2).shape[1] if (zc2 is not None or mu2 is not None) else 0
        per = torch.empty((B, 4), device=out.device, dtype=torch.float32)
        vals = torch.empty(4, device=out.device, dtype=torch.float32)
        check(lib().hopk_step_losses_fwd(ptr(out2), ptr(tgt2), ptr(rnd2), ptr(zc2), ptr(zr2), ptr(mu2), ptr(lv2), B, TP, Z,
                                         float(w_reg), float(w_div), float(w_kld), ptr(per), ptr(vals), ptr(_ticket(out.device)),
                                         stream_ptr()))
        ctx.save_for_backward(out2, tgt2, rnd2, mu2, lv2, per)
        ctx.meta = (B, TP, Z, float(w_reg), float(w_div), float(w_kld), out.shape, None if mu is None else mu.shape)
        ctx.mark_non_differentiable(vals)
        return vals[0].clone(), vals

    @staticmethod
    def backward(ctx, gl, _):
        out2, tgt2, rnd2, mu2, lv2, per = ctx.saved_tensors
        B, TP, Z, w_reg, w_div, w_kld, oshape, mshape = ctx.meta
        gl = f32c(gl).reshape(1)
        dout = torch.empty_like(out2)
        dmu = torch.empty_like(mu2) if mu2 is not None else None
        dlv = torch.empty_like(lv2) if lv2 is not None else None
        check(lib().hopk_step_losses_bwd(ptr(out2), ptr(tgt2), ptr(rnd2), ptr(mu2), ptr(lv2), ptr(per), ptr(gl), B, TP, Z, w_reg, w_div,
                                         w_kld, ptr(dout), ptr(dmu), ptr(dlv), stream_ptr()))
        return (dout.view(oshape), None, None, None, None, None if dmu is None else dmu.view(mshape),
                None if dlv is None else dlv.view(mshape), None, None, None)


def step_losses(out, tgt, rnd=None, zc=None, zr=None, mu=None, logvar=None, w_reg=1.0, w_div=0.0, w_kld=0.0):
    if not out.is_cuda:
        raise RuntimeError('hop_b200.losses needs CUDA tensors (no CPU fallback)')
    return _StepLossFn.apply(out, tgt, rnd, zc, zr, mu, logvar, w_reg, w_div, w_kld)
