"""Probe: how fast does cuDNN run the 4-layer bi-GRU decoder (hidden 350, input 992, batch 128, 34 steps) per dtype / TF32 flag?"""
import torch, json
dev = torch.device('cuda')
def run(H, I, dtype, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    g = torch.nn.GRU(I, hidden_size=H, num_layers=4, batch_first=True, bidirectional=True).to(dev)
    x = torch.randn(128, 34, I, device=dev, requires_grad=True)
    def step():
        with torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            y, _ = g(x)
        y.float().sum().backward()
    def fwd():
        with torch.no_grad(), torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            g(x)
    out = {}
    for name, fn in (('fwd_bwd_ms', step), ('fwd_nograd_ms', fwd)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 10
    return out
for dt, tf32 in ((None, False), (None, True), (torch.bfloat16, False), (torch.float16, False)):
    print(json.dumps({'H': 350, 'I': 992, 'dtype': str(dt), 'tf32': tf32} | run(350, 992, dt, tf32)))
