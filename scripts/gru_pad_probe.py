"""Probe: does cuDNN run the 4-layer bi-GRU decoder faster when the hidden size is padded 350 -> 352 / 384?"""
import torch, time, json
dev = torch.device('cuda')
def run(H, I, dtype):
    g = torch.nn.GRU(I, hidden_size=H, num_layers=4, batch_first=True, bidirectional=True).to(dev)
    x = torch.randn(128, 34, I, device=dev, requires_grad=True)
    def step():
        with torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            y, _ = g(x)
        y.float().sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for H in (350, 352, 384):
    for dt in (torch.bfloat16, torch.float16, None):
        print(json.dumps({'H': H, 'I': 992, 'dtype': str(dt), 'fwd_bwd_ms': run(H, 992, dt)}))
