"""TFLOP/s of hopk_gemm_bf16 (TMA + tcgen05) next to torch.matmul (cuBLAS bf16) on the step's GEMM shapes.

    python scripts/bench_gemm.py [out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hop_b200 import _lib  # noqa: E402

SHAPES = [  # name, M, N, K, a_mn, b_mn, splits
    ('out_projection 4352x768x1024', 4352, 768, 1024, 0, 0, 1),
    ('query_projection 4352x1024x128', 4352, 1024, 128, 0, 0, 1),
    ('kv_projection 1500x1024x768', 1500, 1024, 768, 0, 0, 1),
    ('mapping 1500x768x30522', 1500, 768, 30522, 0, 1, 8),
    ('mapping_wgrad 1500x30522x768', 1500, 30522, 768, 0, 0, 1),
    ('gru_in_l0 4352x2112x992', 4352, 2112, 992, 0, 0, 1),
    ('gru_in_l1 4352x2112x704', 4352, 2112, 704, 0, 0, 1),
    ('gru_dx 4352x992x2112', 4352, 992, 2112, 0, 1, 1),
    ('gru_dw 2112x992x4352', 2112, 992, 4352, 1, 1, 2),
    ('beat1 2048x1700x3400', 2048, 1700, 3400, 0, 0, 1),
    ('bert_ffn1 4352x3072x768', 4352, 3072, 768, 0, 0, 1),
    ('square 8192', 8192, 8192, 8192, 0, 0, 1),
]


def main():
    L = _lib.lib()
    dev = torch.device('cuda:0')
    rows = []
    pad8 = lambda n: (n + 7) // 8 * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = [a for a in sys.argv[1:] if a.startswith('--only=')]
    shapes = [sh for sh in SHAPES if not only or only[0][7:] in sh[0]]
    for name, M, N, K, a_mn, b_mn, splits in shapes:
        A = torch.randn((K, pad8(M)) if a_mn else (M, pad8(K)), device=dev).bfloat16()
        B = torch.randn((K, pad8(N)) if b_mn else (N, pad8(K)), device=dev).bfloat16()
        C = torch.empty(M, pad8(N), device=dev)
        flags = (_lib.GEMM_A_MN if a_mn else 0) | (_lib.GEMM_B_MN if b_mn else 0)

        def ours():
            _lib.check(L.hopk_gemm_bf16(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), None, None, None, M, N, K, A.stride(0), B.stride(0),
                                        C.stride(0), flags, 0.0, splits, _lib.stream_ptr()))
        Am = A[:, :M].t() if a_mn else A[:, :K]
        Bm = B[:, :N] if b_mn else B[:, :K].t()
        Co = torch.empty(M, N, device=dev, dtype=torch.bfloat16)

        def cublas():
            torch.matmul(Am, Bm, out=Co)
        res = {}
        warm = '--warm' in sys.argv      # back-to-back launches on L2-resident operands (20 per event pair) instead of one cold launch
        for nm, fn in (('ours', ours), ('cublas', cublas)):
            for _ in range(3):
                fn()
            if warm:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[nm] = e0.elapsed_time(e1) / 20
                continue
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[nm] = ts[len(ts) // 2]
        fl = 2.0 * M * N * K
        rows.append({'shape': name, 'ours_us': round(res['ours'] * 1e3, 1), 'cublas_us': round(res['cublas'] * 1e3, 1),
                     'ours_tflops': round(fl / res['ours'] / 1e9, 1), 'cublas_tflops': round(fl / res['cublas'] / 1e9, 1)})
        print(rows[-1], flush=True)
    if len(sys.argv) > 1 and not sys.argv[1].startswith('--'):
        json.dump(rows, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
