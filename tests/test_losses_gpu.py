"""GPU: the fused generator-loss kernels (hop_b200/losses.py, csrc/glue.cu) against the reference's own expressions
(train_eval/train_llm.py:46-79) evaluated by torch in float64: values and gradients within 1e-5."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _reference(out, tgt, rnd, zc, zr, mu, logvar, w_reg, w_div, w_kld):
    huber = F.smooth_l1_loss(out / 0.1, tgt / 0.1) * 0.1
    loss = huber * w_reg
    div = kld = None
    if rnd is not None:
        beta = 0.05
        pose = F.smooth_l1_loss(out / beta, rnd.detach() / beta, reduction='none') * beta
        pose = pose.sum(dim=1).sum(dim=1)
        pose = pose.view(pose.shape[0], -1).mean(1)
        zl1 = F.l1_loss(zc.detach(), zr.detach(), reduction='none')
        zl1 = zl1.view(zl1.shape[0], -1).mean(1)
        div = torch.clamp(-(pose / (zl1 + 1.0e-5)), min=-1000).mean()
        loss = loss + div * w_div
    if mu is not None:
        kld = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
        loss = loss + kld * w_kld
    return loss, huber, div, kld


@pytest.mark.parametrize('mode', ['speaker', 'random', 'none'])
@pytest.mark.parametrize('B,T,P,Z', [(128, 34, 27, 32), (5, 34, 126, 16)])
def test_step_losses_match_reference_expressions(mode, B, T, P, Z, cuda):
    from hop_b200.losses import step_losses
    g = torch.Generator(device='cpu').manual_seed(B + P)
    out = (torch.randn(B, T, P, generator=g) * 0.2).to(cuda).requires_grad_(True)      # both smooth-l1 branches are exercised
    tgt = (torch.randn(B, T, P, generator=g) * 0.2).to(cuda)
    rnd = (out.detach() + torch.randn(B, T, P, generator=g).to(cuda) * 0.05) if mode != 'none' else None
    zc = torch.randn(B, Z, generator=g).to(cuda) if mode != 'none' else None
    zr = torch.randn(B, Z, generator=g).to(cuda) if mode != 'none' else None
    if zr is not None:
        zr[0] = zc[0]                                            # z_l1 = 0: the clamp at -1000 is active for this sample
    mu = torch.randn(B, Z, generator=g).to(cuda).requires_grad_(True) if mode == 'speaker' else None
    lv = (torch.randn(B, Z, generator=g) * 0.5).to(cuda).requires_grad_(True) if mode == 'speaker' else None
    w = (5.0, 0.05 if mode != 'none' else 0.0, 0.1 if mode == 'speaker' else 0.0)
    loss, vals = step_losses(out, tgt, rnd, zc, zr, mu, lv, *w)
    up = torch.tensor(0.7, device=cuda)
    (loss * up).backward()
    d = lambda t: None if t is None else t.detach().double().requires_grad_(t.requires_grad)
    o2, m2, l2 = d(out), d(mu), d(lv)
    rl, rh, rd, rk = _reference(o2, d(tgt), d(rnd), d(zc), d(zr), m2, l2, *w)
    (rl * 0.7).backward()
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(loss, rl) < 1e-5 and rel(vals[0], rl) < 1e-5 and rel(vals[1], rh) < 1e-5
    if rd is not None:
        assert rel(vals[2], rd) < 1e-5
    if rk is not None:
        assert rel(vals[3], rk) < 1e-5
    assert rel(out.grad, o2.grad) < 1e-5
    if mu is not None:
        assert rel(mu.grad, m2.grad) < 1e-5 and rel(lv.grad, l2.grad) < 1e-5
    # the ticket is left zero: a second call gives the same numbers
    loss2, _ = step_losses(out.detach(), tgt, rnd, zc, zr, None if mu is None else mu.detach(), None if lv is None else lv.detach(), *w)
    assert torch.equal(loss2, loss.detach())
