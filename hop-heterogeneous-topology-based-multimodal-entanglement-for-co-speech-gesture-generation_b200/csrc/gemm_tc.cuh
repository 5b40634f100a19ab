// gemm_tc.cuh -- tcgen05 GEMM skeleton: C[m][n] = sum_k A(m,k) * B(n,k) with bf16 operands and fp32 TMEM accumulators.
//
// Same functor contracts as the FFMA skeleton (gemm_core.cuh): loaders return fp32 elements (so gathers, BatchNorm
// folding, dilated taps and concatenations fuse into the tile load exactly as on the fp32 path); the conversion to bf16
// happens on the way into shared memory.  One CTA = one 128 x 128 output tile; K is consumed in slabs of 64:
//   all 256 threads stage slab k+1 (global -> regs -> bf16 -> swizzled smem) while the tensor core works on slab k
//   (two smem stages, one mbarrier per stage armed by tcgen05.commit); one elected thread issues the UMMAs;
//   the epilogue reads the accumulator row-per-thread with tcgen05.ld (warps 0-3 = TMEM lane quarters).
// Epilogue contract:  void row32(int m, int n0, const float (&v)[32]);   // 32 consecutive columns of row m
//                     void finish();                                       // once per thread (column reductions)
#pragma once
#include "tc_core.cuh"

namespace hopk {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
// dynamic smem: 2 stages x (A slab + B slab) + 1024 B alignment slack
constexpr size_t TC_SMEM_BYTES = 2 * (tc::slab_bytes(TC_BM) + tc::slab_bytes(TC_BN)) + 1024;

template <class L>
__device__ __forceinline__ void tc_stage_slab(uint8_t* slab, const L& ld, int row0, int nrows_valid, int k0, int kmax)
{
    // 128 rows x 8 chunks; lane order follows the loader's contiguous index so global reads coalesce
#pragma unroll
    for (int it = 0; it < (128 * 8) / TC_THREADS; ++it) {
        int idx = threadIdx.x + it * TC_THREADS;
        int row, ch;
        if (L::kFast) { ch = idx & 7; row = idx >> 3; } else { row = idx & 127; ch = idx >> 7; }
        float f[8];
        int i = row0 + row;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int k = k0 + ch * 8 + j;
            f[j] = (i < nrows_valid && k < kmax) ? ld(i, k) : 0.f;
        }
        tc::slab_store8(slab, row, ch, f);
    }
}

template <class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(int M, int N, int K, ALoad aload, BLoad bload, Epi epi)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t A_BYTES = tc::slab_bytes(TC_BM), B_BYTES = tc::slab_bytes(TC_BN), STAGE = A_BYTES + B_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;

    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_smem;
    constexpr uint32_t idesc = tc::idesc_bf16(TC_BM, TC_BN, 0, 0);

    const int nslabs = (K + TC_BK - 1) / TC_BK;
    for (int ks = 0; ks < nslabs; ++ks) {
        const int buf = ks & 1;
        if (ks >= 2) tc::mbar_wait(&bars[buf], ((ks >> 1) - 1) & 1);      // UMMAs that read this stage are done
        uint8_t* sa = smem + buf * STAGE;
        uint8_t* sb = sa + A_BYTES;
        tc_stage_slab(sa, aload, m0, M, ks * TC_BK, K);
        tc_stage_slab(sb, bload, n0, N, ks * TC_BK, K);
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t a_addr = tc::smem_u32(sa), b_addr = tc::smem_u32(sb);
#pragma unroll
            for (int j = 0; j < TC_BK / 16; ++j)
                tc::mma_bf16(tmem, tc::desc_kmajor(a_addr, j), tc::desc_kmajor(b_addr, j), idesc, (ks | j) != 0);
            tc::mma_commit(&bars[buf]);
        }
    }
    if (nslabs > 0) {
        const int last = nslabs - 1;
        tc::mbar_wait(&bars[last & 1], (last >> 1) & 1);                  // commits complete in order
    }
    tc::fence_after_sync();

    Epi e = epi;
    if (warp < 4) {
        const int m = m0 + warp * 32 + (tid & 31);
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
            float v[32];
            if (nslabs > 0) tc::tmem_ld32(lane_addr + c * 32, v);
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (m < M) e.row32(m, n0 + c * 32, v);
        }
    }
    e.finish();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------- epilogues for the tensor-core skeleton
// flags: 1 = relu on output, 4 = mask by aux > 0
struct EpiStoreTC {
    float* out; long ld; const float* bias; const float* aux; int N; int flags;
    __device__ __forceinline__ void row32(int m, int n0, const float (&v)[32]) {
        float* row = out + (size_t)m * ld;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            if (n < N) {
                float x = v[j];
                if (bias) x += __ldg(bias + n);
                if (flags & 1) x = fmaxf(x, 0.f);
                if (flags & 4) x = (__ldg(aux + (size_t)m * ld + n) > 0.f) ? x : 0.f;
                row[n] = x;
            }
        }
    }
    __device__ __forceinline__ void finish() {}
};

template <class AL, class BL, class EP>
static cudaError_t launch_gemm_tc(int M, int N, int K, AL a, BL b, EP e, cudaStream_t st)
{
    auto kern = gemm_tc_kernel<AL, BL, EP>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    if (err != cudaSuccess) return err;
    dim3 grid((N + TC_BN - 1) / TC_BN, (M + TC_BM - 1) / TC_BM);
    kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(M, N, K, a, b, e);
    return cudaGetLastError();
}

}  // namespace hopk
