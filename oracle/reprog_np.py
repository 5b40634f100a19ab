"""CPU oracle (TEST INFRASTRUCTURE ONLY) for HOP's audio->text reprogramming cross-attention.

numpy float64 restatement (forward + hand-derived backward) of
``ReprogrammingLayer`` in the reference's ``model/HOP.py``:
  __init__ parameter shapes   model/HOP.py:256-268
  forward                     model/HOP.py:271-285
  reprogramming               model/HOP.py:289-299
Written from SURVEY.md Appendix A, not from the reference's code.  Product code
must never import this; tests / smoke / bench's CPU legs use it as the checker.

Parity status: no golden vectors exist in the reference; pinned against the
reference module executed here (tests/golden/make_golden.py -> reprog_*.npz).

Dropout: torch's RNG stream cannot be reproduced by a kernel, so the build
defines its own counter-based mask (``dropout_keep``) that the CUDA kernel and
this oracle share bit-for-bit; parity with the *reference* is checked at p=0.
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _lowbias32(x):
    """32-bit integer finaliser (xorshift-multiply), vectorised on uint64 holders."""
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def dropout_threshold(p):
    """16-bit integer threshold of the mask; the realised drop probability is thr / 65536 (0.100006 for p = 0.1)."""
    return int(round(float(p) * 65536.0))


def dropout_scale(p):
    """1 / (1 - realised drop probability): what kept probabilities are multiplied by (keeps the layer unbiased)."""
    return 65536.0 / (65536.0 - dropout_threshold(p))


def dropout_keep(seed, idx, p):
    """Keep-mask for flat element indices ``idx`` (uint64) under 64-bit ``seed``.

    One 32-bit hash serves two neighbouring elements: q = idx >> 1, h = f(q_lo + f(q_hi ^ seed_hi) ^ seed_lo) with
    f = lowbias32; element idx uses the low (even idx) or high (odd idx) 16 bits of h and is kept iff that field is
    >= round(p * 65536).  The CUDA kernels (csrc/xattn.cu: keep_mask(), csrc/xattn_tc.cu) evaluate exactly this.
    """
    idx = np.asarray(idx, dtype=np.uint64)
    seed = np.uint64(seed)
    s_lo, s_hi = seed & _M32, seed >> np.uint64(32)
    q = idx >> np.uint64(1)
    q_lo, q_hi = q & _M32, q >> np.uint64(32)
    h = _lowbias32(((q_lo + _lowbias32(q_hi ^ s_hi)) & _M32) ^ s_lo)
    field = np.where((idx & np.uint64(1)) == 1, h >> np.uint64(16), h & np.uint64(0xFFFF))
    return field >= np.uint64(dropout_threshold(p))


def _id(a):
    return a


def forward(P, target, source, value, n_heads, p_drop=0.0, seed=0, keep=False, q=_id, relu_mask=None):
    """ReprogrammingLayer.forward.  target (B,L,dm); source,value (S,dllm).  Returns (B,L,dllm).

    ``q`` quantises both operands of the four projection GEMMs (identity = exact reference;
    oracle.gwnet_np.bf16_round = quantisation-aware reference of the bf16 tensor-core mode).
    ``relu_mask`` (B, L, H*E) of 0/1 pins the gate pattern of the ReLU in front of the out projection (HOP.py:284) to an
    implementation's own pattern (see oracle/hop_torch.py::gwnet_forward for why)."""
    P = {k: np.asarray(v, dtype=np.float64) for k, v in P.items()}
    x = np.asarray(target, np.float64)
    src = np.asarray(source, np.float64)
    val = np.asarray(value, np.float64)
    B, L, _ = x.shape
    S = src.shape[0]
    H = n_heads
    Q = (q(x) @ q(P['query_projection.weight']).T + P['query_projection.bias']).reshape(B, L, H, -1)
    K = (q(src) @ q(P['key_projection.weight']).T + P['key_projection.bias']).reshape(S, H, -1)
    V = (q(val) @ q(P['value_projection.weight']).T + P['value_projection.bias']).reshape(S, H, -1)
    E = Q.shape[-1]
    scale = 1.0 / np.sqrt(E)
    sc = np.einsum('blhe,she->bhls', Q, K) * scale
    sc = sc - sc.max(axis=-1, keepdims=True)
    pr = np.exp(sc)
    pr /= pr.sum(axis=-1, keepdims=True)
    if p_drop > 0:
        idx = np.arange(B * H * L * S, dtype=np.uint64).reshape(B, H, L, S)
        mask = dropout_keep(seed, idx, p_drop) * dropout_scale(p_drop)
    else:
        mask = np.ones_like(pr)
    pd = pr * mask
    O = np.einsum('bhls,she->blhe', pd, V).reshape(B, L, H * E)
    gate = (O > 0) if relu_mask is None else (np.asarray(relu_mask) > 0)
    R = O * gate
    Y = q(R) @ q(P['out_projection.weight']).T + P['out_projection.bias']
    cache = dict(x=x, src=src, val=val, Q=Q, K=K, V=V, pr=pr, mask=mask, O=O, R=R, gate=gate, scale=scale) if keep else None
    return Y, cache


def backward(P, cache, dY, n_heads, q=_id):
    """Returns (dtarget, dsource, dvalue, grads-by-state_dict-name)."""
    P = {k: np.asarray(v, dtype=np.float64) for k, v in P.items()}
    c = cache
    dY = np.asarray(dY, np.float64)
    B, L, _ = dY.shape
    H = n_heads
    G = {}
    dYq = q(dY)
    G['out_projection.weight'] = np.einsum('blo,bli->oi', dYq, q(c['R']))
    G['out_projection.bias'] = dYq.sum(axis=(0, 1))      # bias gradient = all-ones column of the weight-gradient GEMM
    dR = dYq @ q(P['out_projection.weight'])
    dO = (dR * c['gate']).reshape(B, L, H, -1)
    dpd = np.einsum('blhe,she->bhls', dO, c['V'])
    dV = np.einsum('bhls,blhe->she', c['pr'] * c['mask'], dO)
    dpr = dpd * c['mask']
    dsc = c['pr'] * (dpr - (dpr * c['pr']).sum(axis=-1, keepdims=True))
    dQ = np.einsum('bhls,she->blhe', dsc, c['K']) * c['scale']
    dK = np.einsum('bhls,blhe->she', dsc, c['Q']) * c['scale']
    S = dK.shape[0]
    dQf, dKf, dVf = q(dQ.reshape(B * L, -1)), q(dK.reshape(S, -1)), q(dV.reshape(S, -1))
    xf = c['x'].reshape(B * L, -1)
    G['query_projection.weight'] = dQf.T @ q(xf)
    G['query_projection.bias'] = dQf.sum(0)
    G['key_projection.weight'] = dKf.T @ q(c['src'])
    G['key_projection.bias'] = dKf.sum(0)
    G['value_projection.weight'] = dVf.T @ q(c['val'])
    G['value_projection.bias'] = dVf.sum(0)
    dx = (dQf @ q(P['query_projection.weight'])).reshape(c['x'].shape)
    dsrc = dKf @ q(P['key_projection.weight'])
    dval = dVf @ q(P['value_projection.weight'])
    return dx, dsrc, dval, G


def init_params(rng, d_model=128, n_heads=8, d_keys=128, d_llm=768):
    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return rng.uniform(-b, b, size=(o, i)), rng.uniform(-b, b, size=(o,))
    P = {}
    for name, (o, i) in dict(query_projection=(d_keys * n_heads, d_model),
                             key_projection=(d_keys * n_heads, d_llm),
                             value_projection=(d_keys * n_heads, d_llm),
                             out_projection=(d_llm, d_keys * n_heads)).items():
        P[name + '.weight'], P[name + '.bias'] = lin(o, i)
    return P
