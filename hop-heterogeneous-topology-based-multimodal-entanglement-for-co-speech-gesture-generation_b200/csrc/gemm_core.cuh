// gemm_core.cuh -- CTA-tiled FP32 (FFMA) GEMM skeleton with functor loaders / epilogues.
//
// Every dense contraction on the fp32-exact path (1e-5 parity mode) goes through this one
// skeleton:  C[m][n] = sum_k A(m,k) * B(k,n).  The operands are *functors*, so gathers
// (dilated taps, BatchNorm folding, NCHW strides, segment concatenation) are fused into the
// tile load and bias / activation / residual / statistics are fused into the epilogue.
//
// Tile: BM x BN outputs per CTA (BM = 64*MG, BN = 64*NG), BK = 16, 256 threads laid out
// 16 x 16; each thread owns MG*4 rows x NG*4 cols (row groups 64 apart, col groups 64 apart)
// so shared-memory reads are 128-bit and conflict free.  Global->register prefetch of the next
// K chunk overlaps the FFMA loop on the current one.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace hopk {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_BK = 16;

// Loader contract:
//   struct L { static constexpr bool kFast;            // true: k is the contiguous index in memory
//              __device__ float operator()(int m_or_n, int k) const; }   // must return 0 outside range
// Epilogue contract (the kernel makes a thread-private copy, so it may hold partial sums):
//   __device__ void operator()(int m, int nb, const float (&v)[4*NG]);
//        row m (< M guaranteed); v[4*g + j] is column nb + 64*g + j (functor checks < N itself)
//   __device__ void flush(int nb);      // once per thread after all rows (column reductions)

template <int MG, int NG, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_kernel(int M, int N, int K, int k_per_split, ALoad aload, BLoad bload, Epi epi)
{
    constexpr int BM = 64 * MG, BN = 64 * NG, BK = GEMM_BK;
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int A_PER_T = BM * BK / GEMM_THREADS;   // 4*MG
    constexpr int B_PER_T = BN * BK / GEMM_THREADS;   // 4*NG
    __shared__ __align__(16) float As[BK * LDA];
    __shared__ __align__(16) float Bs[BK * LDB];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);

    float acc[4 * MG][4 * NG];
#pragma unroll
    for (int i = 0; i < 4 * MG; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NG; ++j) acc[i][j] = 0.f;

    float ra[A_PER_T], rb[B_PER_T];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
            int idx = tid + j * GEMM_THREADS;
            int mm, kk;
            if (ALoad::kFast) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
            int m = m0 + mm, k = k0 + kk;
            ra[j] = (m < M && k < k_end) ? aload(m, k) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < B_PER_T; ++j) {
            int idx = tid + j * GEMM_THREADS;
            int nn, kk;
            if (BLoad::kFast) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
            int n = n0 + nn, k = k0 + kk;
            rb[j] = (n < N && k < k_end) ? bload(n, k) : 0.f;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
            int idx = tid + j * GEMM_THREADS;
            int mm, kk;
            if (ALoad::kFast) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
            As[kk * LDA + mm] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < B_PER_T; ++j) {
            int idx = tid + j * GEMM_THREADS;
            int nn, kk;
            if (BLoad::kFast) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
            Bs[kk * LDB + nn] = rb[j];
        }
    };

    if (k_begin < k_end) fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        stash();
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4 * MG], b[4 * NG];
#pragma unroll
            for (int g = 0; g < MG; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&As[kk * LDA + g * 64 + ty * 4]);
                a[4 * g] = t.x; a[4 * g + 1] = t.y; a[4 * g + 2] = t.z; a[4 * g + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[kk * LDB + g * 64 + tx * 4]);
                b[4 * g] = t.x; b[4 * g + 1] = t.y; b[4 * g + 2] = t.z; b[4 * g + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 4 * MG; ++i)
#pragma unroll
                for (int j = 0; j < 4 * NG; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    Epi e = epi;   // thread-private copy: epilogues may carry per-thread partial sums
#pragma unroll
    for (int gi = 0; gi < MG; ++gi)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + gi * 64 + ty * 4 + i;
            if (m < M) e(m, n0 + tx * 4, acc[4 * gi + i]);
        }
    e.flush(n0 + tx * 4);
}

// ---- small helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace hopk
