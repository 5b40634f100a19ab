"""Run exactly N calls of one hand-written kernel group (TED shapes, B = 128) so that an ncu metrics pass around this
process can attribute DRAM traffic to it:  python scripts/traffic_probe.py gwnet_fwd|gwnet_fwdbwd|xattn_fwd|xattn_fwdbwd [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
what, N = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda'); torch.manual_seed(0)
B = 128
if what.startswith('gwnet'):
    from hop_b200 import gwnet as G
    m = G.gwnet(dev, 9, dropout=0, in_dim=173, out_dim=173, residual_channels=64, dilation_channels=64, skip_channels=256,
                end_channels=512).to(dev).set_precision('bf16')
    x = torch.randn(B, 16, 9, 173, device=dev).permute(0, 3, 2, 1).requires_grad_(True)
    dy = torch.randn(B, 173, 9, 4, device=dev)
    for _ in range(N):
        if what == 'gwnet_fwd':
            with torch.no_grad():
                m(x)
        else:
            m.zero_grad(set_to_none=True); x.grad = None
            m(x).backward(dy)
else:
    from hop_b200.HOP import _XattnFn
    L, H, E, S = 34, 8, 128, 1500
    q = torch.randn(B, L, H, E, device=dev, requires_grad=True)
    k = torch.randn(S, H, E, device=dev, requires_grad=True)
    v = torch.randn(S, H, E, device=dev, requires_grad=True)
    do = torch.randn(B, L, H, E, device=dev)
    for _ in range(N):
        if what == 'xattn_fwd':
            with torch.no_grad():
                _XattnFn.apply(q, k, v, 0.1, 5, True)
        else:
            torch.autograd.grad(_XattnFn.apply(q, k, v, 0.1, 5, True), (q, k, v), do)
torch.cuda.synchronize()
