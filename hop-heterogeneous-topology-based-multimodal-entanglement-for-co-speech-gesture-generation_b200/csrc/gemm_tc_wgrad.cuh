// gemm_tc_wgrad.cuh -- tcgen05 skeleton for weight gradients:  out[i][j] = sum_r A(r, i) * B(r, j).
//
// The contraction runs over data rows r (B*T*V of them) and both operands are row-major in memory with their
// channels contiguous, i.e. they are *already* in the MN-major slab format ([K rows][64 M/N columns], 128-byte rows):
// they are staged with 128-bit loads exactly like a forward operand and handed to the tensor core through MN-major
// descriptors (a_major = b_major = 1) -- no transpose anywhere.  One CTA = one 128 x BN output tile over one split of
// the rows; slabs of 64 rows, two smem stages, accumulator in TMEM; partial tiles are combined with fp32 atomics by the
// epilogue (the output is zeroed by the caller).  Column sums of A (= the bias gradient) fall out of the staging loop.
// Loader contract:  int ncols;  void ld8(int r, int c0, float (&f)[8]) const;   // 8 consecutive columns, 0 outside
// Epilogue contract: row32 / finish as in gemm_tc.cuh, plus  void bias(int i, float v);  // column sum of A (atomic)
#pragma once
#include <math.h>
#include "common.cuh"
#include "gemm_tc.cuh"

namespace hopk {

constexpr int WG_ROWS = 64;                                   // contraction rows per slab
constexpr uint32_t WG_SLAB = tc::slab_bytes(WG_ROWS);         // 8 KB
template <int BN>
constexpr size_t wg_smem_bytes() { return 2 * (2 * WG_SLAB + (BN / 64) * WG_SLAB) + 1024; }

template <int BN, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_wgrad_kernel(int R, int Mo, int No, int r_per_split, ALoad aload, BLoad bload, Epi epi, int want_bias)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[256];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t A_BYTES = 2 * WG_SLAB, B_BYTES = (BN / 64) * WG_SLAB, STAGE = A_BYTES + B_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int i0 = blockIdx.y * TC_BM, j0 = blockIdx.x * BN;
    const int r_begin = blockIdx.z * r_per_split;
    const int r_end = min(R, r_begin + r_per_split);

    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_barrier_init(); }
    red[tid] = 0.f;
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, BN);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_smem;
    constexpr uint32_t idesc = tc::idesc_bf16(TC_BM, BN, 1, 1);
    const bool do_bias = want_bias && blockIdx.x == 0;
    float bsum[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) bsum[q] = 0.f;

    const int nslabs = r_end > r_begin ? (r_end - r_begin + WG_ROWS - 1) / WG_ROWS : 0;
    constexpr int A_ITEMS = (WG_ROWS * 16) / TC_THREADS, B_ITEMS = (WG_ROWS * (BN / 8)) / TC_THREADS;
    float fa[A_ITEMS][8], fb[B_ITEMS][8];
    // load phase: every global load of a slab in flight together (and across the barrier / UMMAs of the previous slab)
    auto load_regs = [&](int ks) {
        const int r0 = r_begin + ks * WG_ROWS;
#pragma unroll
        for (int it = 0; it < A_ITEMS; ++it) {                                // A: 64 rows x 16 chunks of 8 columns
            int idx = tid + it * TC_THREADS;
            int ch = idx & 15, row = idx >> 4;
            if (r0 + row < r_end) aload.ld8(r0 + row, i0 + ch * 8, fa[it]);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) fa[it][q] = 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < B_ITEMS; ++it) {                                // B: 64 rows x BN/8 chunks
            int idx = tid + it * TC_THREADS;
            int ch = idx % (BN / 8), row = idx / (BN / 8);
            if (r0 + row < r_end) bload.ld8(r0 + row, j0 + ch * 8, fb[it]);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) fb[it][q] = 0.f;
            }
        }
    };
    if (nslabs > 0) load_regs(0);
    for (int ks = 0; ks < nslabs; ++ks) {
        const int buf = ks & 1;
        if (ks >= 2) tc::mbar_wait(&bars[buf], ((ks >> 1) - 1) & 1);
        uint8_t* sa = smem + buf * STAGE;
        uint8_t* sb = sa + A_BYTES;
#pragma unroll
        for (int it = 0; it < A_ITEMS; ++it) {
            int idx = tid + it * TC_THREADS;
            int ch = idx & 15, row = idx >> 4;
            if (do_bias) {
#pragma unroll
                for (int q = 0; q < 8; ++q) bsum[q] += fa[it][q];
            }
            tc::slab_store8(sa + (ch >> 3) * WG_SLAB, row, ch & 7, fa[it]);
        }
#pragma unroll
        for (int it = 0; it < B_ITEMS; ++it) {
            int idx = tid + it * TC_THREADS;
            int ch = idx % (BN / 8), row = idx / (BN / 8);
            tc::slab_store8(sb + (ch >> 3) * WG_SLAB, row, ch & 7, fb[it]);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t a_addr = tc::smem_u32(sa), b_addr = tc::smem_u32(sb);
#pragma unroll
            for (int t = 0; t < WG_ROWS / 16; ++t)
                tc::mma_bf16(tmem, tc::desc_mnmajor(a_addr, WG_SLAB, t), tc::desc_mnmajor(b_addr, WG_SLAB, t), idesc, (ks | t) != 0);
            tc::mma_commit(&bars[buf]);
        }
        // next slab's global loads fly under the UMMAs (issued after the proxy fence: MEMBAR would wait for them)
        if (ks + 1 < nslabs) load_regs(ks + 1);
    }
    if (nslabs > 0) {
        const int last = nslabs - 1;
        tc::mbar_wait(&bars[last & 1], (last >> 1) & 1);
    }
    tc::fence_after_sync();

    Epi e = epi;
    if (do_bias) {                                   // thread's column chunk is fixed: (tid & 15) * 8
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int i = i0 + (tid & 15) * 8 + q;
            if (i < Mo && bsum[q] != 0.f) e.bias(i, bsum[q]);
        }
    }
    if (warp < 4) {
        const int i = i0 + warp * 32 + (tid & 31);
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            if (nslabs > 0) tc::tmem_ld32(lane_addr + c * 32, v);
            else {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0.f;
            }
            e.row32(i, i < Mo, j0 + c * 32, v, red);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    e.finish(red);
    if (warp == 0) tc::tmem_dealloc(tmem, BN);
}

constexpr int WG_MIN_SLABS = 16;                              // contraction slabs (64 rows each) per split-K CTA at least
template <int BN, class AL, class BL, class EP>
static cudaError_t launch_gemm_tc_wgrad(int R, int Mo, int No, AL a, BL b, EP e, bool want_bias, cudaStream_t st)
{
    auto kern = gemm_tc_wgrad_kernel<BN, AL, BL, EP>;
    {
        cudaError_t err = configure_smem_once((const void*)kern, wg_smem_bytes<BN>());      // once per (device, instantiation)
        if (err != cudaSuccess) return err;
    }
    const int tiles = ((Mo + TC_BM - 1) / TC_BM) * ((No + BN - 1) / BN);
    // Split-K factor, measured on the whole backward (TED B = 128 / Expressive / B = 1024, ms): 4 slabs per split 1.14 / 3.92 /
    // 5.15, 8: 1.00 / 3.42 / -, 16: 0.94 / 3.40 / 4.18, 32: 1.38 / 4.47 / 4.10; a sqrt(R) rule fitted to TED loses 1.3 ms on
    // Expressive.  These kernels run on side streams beside the dependent dx / dy chain: a grid that fills every SM slot
    // delays the chain's own small kernels (and adds reduction traffic), a grid of a few dozen CTAs leaves room for them.
    int splits = (2 * 148 + tiles - 1) / tiles;                       // about two CTAs per SM at most
    int max_splits = (R + WG_MIN_SLABS * WG_ROWS - 1) / (WG_MIN_SLABS * WG_ROWS);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rper = (((R + splits - 1) / splits + WG_ROWS - 1) / WG_ROWS) * WG_ROWS;
    splits = (R + rper - 1) / rper;
    dim3 grid((No + BN - 1) / BN, (Mo + TC_BM - 1) / TC_BM, splits);
    kern<<<grid, TC_THREADS, wg_smem_bytes<BN>(), st>>>(R, Mo, No, rper, a, b, e, want_bias ? 1 : 0);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- row-major operand loaders (8 columns at a time)
// mode 0 plain, 1 relu(value), 2 value masked by aux > 0
struct W8Plain {
    const float* p; const float* aux; long ld; int ncols; int mode;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        const float* row = p + (size_t)r * ld;
        if (c0 + 8 <= ncols && ((ld | c0) & 3) == 0) {
            float4 a = __ldg(reinterpret_cast<const float4*>(row + c0)), b = __ldg(reinterpret_cast<const float4*>(row + c0 + 4));
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = c0 + q < ncols ? __ldg(row + c0 + q) : 0.f;
        }
        if (mode == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
        } else if (mode == 2) {
            const float* arow = aux + (size_t)r * ld;
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = (c0 + q < ncols && __ldg(arow + c0 + q) > 0.f) ? f[q] : 0.f;
        }
    }
};

}  // namespace hopk
