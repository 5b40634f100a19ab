"""GPU parity of the TMA + tcgen05 bf16 GEMM (csrc/gemm_tma.cu, hopk_gemm_bf16) through the C ABI.

Operands are bf16 tensors, the reference is float64 matmul of the same bf16 values, so the only differences are the
fp32 accumulation order (1e-5 relative to the output scale at these K) and, for bf16 outputs, the final rounding (2^-9)."""
import numpy as np
import pytest
import torch

from tests.util import Report, relerr

pytestmark = pytest.mark.gpu


def _gemm(A, B, M, N, K, a_mn=False, b_mn=False, bias=None, addend=None, out_bf16=False, act=None, slope=0.0, splits=1,
          ldc=None, accumulate_into=None):
    from hop_b200 import _lib
    L = _lib.lib()
    dev = A.device
    ldc = ldc or N
    flags = (_lib.GEMM_A_MN if a_mn else 0) | (_lib.GEMM_B_MN if b_mn else 0) | (_lib.GEMM_OUT_BF16 if out_bf16 else 0)
    flags |= {None: 0, 'relu': _lib.GEMM_RELU, 'leaky': _lib.GEMM_LEAKY, 'gelu': _lib.GEMM_GELU}[act]
    if accumulate_into is not None:
        C = accumulate_into
        flags |= _lib.GEMM_ACCUMULATE
    else:
        C = torch.full((M, ldc), float('nan'), device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    _lib.check(L.hopk_gemm_bf16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(addend), None, M, N, K,
                                A.stride(0), B.stride(0), ldc, flags, float(slope), splits, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return C[:, :N]


CASES = [
    # M, N, K, a_mn, b_mn
    (128, 128, 64, False, False),
    (4352, 1024, 128, False, False),      # query projection
    (4352, 768, 1024, False, False),      # out projection
    (1500, 1024, 768, False, False),      # key / value projection
    (300, 200, 72, False, False),         # ragged M, N, K (K tail zero-filled by TMA)
    (4352, 1056, 992, False, False),      # GRU input projection, layer 0 (TED)
    (4352, 992, 1056, False, True),       # dX = dG @ W_ih  (B MN-major)
    (1056, 992, 4352, True, True),        # dW = dG^T @ X   (both MN-major)
    (1500, 768, 30522, False, True),      # text prototypes: W_map @ WE (B MN-major), long K with a tail
    (130, 70, 200, True, False),          # A MN-major alone, ragged
]


@pytest.mark.parametrize('M,N,K,a_mn,b_mn', CASES)
def test_gemm_bf16_layouts(M, N, K, a_mn, b_mn, cuda):
    g = torch.Generator(device='cpu').manual_seed(M + 3 * N + 7 * K)
    pad8 = lambda n: (n + 7) // 8 * 8
    A = torch.randn((K, pad8(M)) if a_mn else (M, pad8(K)), generator=g).bfloat16().to(cuda)
    B = torch.randn((K, pad8(N)) if b_mn else (N, pad8(K)), generator=g).bfloat16().to(cuda)
    Ad = (A[:, :M].double().t() if a_mn else A[:, :K].double())
    Bd = (B[:, :N].double().t() if b_mn else B[:, :K].double())
    ref = (Ad @ Bd.t()).cpu().numpy()
    C = _gemm(A, B, M, N, K, a_mn, b_mn)
    rep = Report(f'gemm_tma_{M}_{N}_{K}_{int(a_mn)}{int(b_mn)}', 2e-5 if K < 8192 else 1e-4)     # fp32 accumulation over K terms
    rep.add('C', relerr(C.double().cpu().numpy(), ref))
    rep.finish()


def test_gemm_bf16_epilogues(cuda):
    g = torch.Generator(device='cpu').manual_seed(5)
    M, N, K = 700, 1050, 352
    A = torch.randn(M, K, generator=g).bfloat16().to(cuda)
    B = torch.randn(N, K, generator=g).bfloat16().to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    ref = A.double() @ B.double().t()
    rep = Report('gemm_tma_epilogues', 2e-5)
    # bias + padded leading dimension of C (N = 1050 is not a multiple of 8)
    C = _gemm(A, B, M, N, K, bias=bias, ldc=1056)
    rep.add('bias, ldc 1056', relerr(C.double().cpu().numpy(), (ref + bias.double()).cpu().numpy()))
    # unaligned rows (ldc = N = 1050): scalar store path
    C = _gemm(A, B, M, N, K, bias=bias)
    rep.add('bias, ldc 1050', relerr(C.double().cpu().numpy(), (ref + bias.double()).cpu().numpy()))
    for act, fn in (('relu', torch.relu), ('leaky', lambda t: torch.nn.functional.leaky_relu(t, 0.2)),
                    ('gelu', torch.nn.functional.gelu)):
        C = _gemm(A, B, M, N, K, bias=bias, act=act, slope=0.2, ldc=1056)
        rep.add(act, relerr(C.double().cpu().numpy(), fn(ref + bias.double()).cpu().numpy()))
    C = _gemm(A, B, M, N, K, bias=bias, out_bf16=True, ldc=1056)
    rep.add('bf16 out', relerr(C.double().cpu().numpy(), (ref + bias.double()).cpu().numpy()), tol=6e-3)
    add = torch.randn(M, 1056, generator=g).to(cuda)
    C = _gemm(A, B, M, N, K, addend=add, ldc=1056)
    rep.add('addend', relerr(C.double().cpu().numpy(), (ref + add[:, :N].double()).cpu().numpy()))
    rep.finish()


def test_gemm_bf16_split_k_and_accumulate(cuda):
    """Weight-gradient shape: small output, long contraction -> split-K with vector atomics; accumulate into an existing C."""
    g = torch.Generator(device='cpu').manual_seed(6)
    M, N, K = 350, 1056, 4352
    A = torch.randn(K, 352, generator=g).bfloat16().to(cuda)          # [K][M] (MN-major), ld 352
    B = torch.randn(K, N, generator=g).bfloat16().to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    ref = A[:, :M].double().t() @ B.double()
    rep = Report('gemm_tma_splitk', 2e-5)
    for splits in (2, 5, 8):
        C = _gemm(A, B, M, N, K, a_mn=True, b_mn=True, splits=splits, bias=bias)
        rep.add(f'splits {splits}', relerr(C.double().cpu().numpy(), (ref + bias.double()).cpu().numpy()))
    C0 = torch.randn(M, N, generator=g).to(cuda)
    C = _gemm(A, B, M, N, K, a_mn=True, b_mn=True, splits=3, accumulate_into=C0.clone())
    rep.add('accumulate', relerr(C.double().cpu().numpy(), (ref + C0.double()).cpu().numpy()))
    rep.finish()


def test_cast_and_colsum(cuda):
    from hop_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device='cpu').manual_seed(8)
    x = torch.randn(333, 350, generator=g).to(cuda)
    y = torch.full((333, 352), float('nan'), device=cuda, dtype=torch.bfloat16)
    _lib.check(L.hopk_cast_bf16(_lib.ptr(x), _lib.ptr(y), 333, 350, 350, 352, 352, 0, _lib.stream_ptr()))
    assert torch.equal(y[:, :350], x.bfloat16()) and float(y[:, 350:].abs().max()) == 0.0
    out = torch.empty(350, device=cuda)
    _lib.check(L.hopk_colsum(_lib.ptr(x), _lib.ptr(out), 333, 350, 350, 0, _lib.stream_ptr()))
    assert relerr(out.cpu().numpy(), x.double().sum(0).cpu().numpy()) < 1e-5
    _lib.check(L.hopk_colsum(_lib.ptr(y), _lib.ptr(out), 333, 350, 352, 1, _lib.stream_ptr()))
    assert relerr(out.cpu().numpy(), y[:, :350].double().sum(0).cpu().numpy()) < 1e-5
