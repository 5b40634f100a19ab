"""Timing experiment: the fused gwnet layer kernel cut short after stage HOPK_FZ_STOP (results are garbage for stop > 0)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hop_b200 import gwnet as G
dev = torch.device('cuda')
B, V = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 9
m = G.gwnet(dev, V, dropout=0, in_dim=173, out_dim=173, residual_channels=64, dilation_channels=64, skip_channels=256, end_channels=512).to(dev).set_precision('bf16')
x = torch.randn(B, 16, V, 173, device=dev).permute(0, 3, 2, 1)
with torch.no_grad():
    for _ in range(5): m(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(3_000_000); e0.record(); m(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps({'stop': os.environ.get('HOPK_FZ_STOP', '0'), 'B': B, 'fwd_ms_median': ts[len(ts) // 2], 'min': ts[0]}))
