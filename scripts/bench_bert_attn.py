import torch, sys
sys.path.insert(0, '.')
from hop_b200 import _lib
L = _lib.lib()
B, S, H, D = 128, 34, 12, 64
qkv = torch.randn(B * S, 3 * H * D, device='cuda').bfloat16(); dctx = torch.randn(B * S, H * D, device='cuda').bfloat16()
ctx = torch.empty_like(dctx); dqkv = torch.empty_like(qkv)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print('attn fwd us', t(lambda: L.hopk_bert_attn_fwd(_lib.ptr(qkv), _lib.ptr(ctx), None, B, S, H, D, _lib.stream_ptr())))
print('attn bwd us', t(lambda: L.hopk_bert_attn_bwd(_lib.ptr(qkv), _lib.ptr(dctx), _lib.ptr(dqkv), B, S, H, D, _lib.stream_ptr())))
