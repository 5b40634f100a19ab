// gemm_tc.cuh -- tcgen05 GEMM skeleton: C[m][n] = sum_k A(m,k) * B(n,k) with bf16 operands and fp32 TMEM accumulators.
//
// Same loader contract as the FFMA skeleton (gemm_core.cuh): loaders return fp32 elements (so gathers, BatchNorm
// folding, dilated taps and concatenations fuse into the tile load exactly as on the fp32 path); the conversion to bf16
// happens on the way into shared memory.  One CTA = one 128 x BN output tile (BN = 64 or 128) over one K split;
// K is consumed in slabs of 64:
//   all 256 threads stage slab k+1 (global -> regs -> bf16 -> swizzled smem) while the tensor core works on slab k
//   (two smem stages, one mbarrier per stage armed by tcgen05.commit); one elected thread issues the UMMAs;
//   the epilogue reads the accumulator row-per-thread with tcgen05.ld (warps 0-3 = TMEM lane quarters).
// Epilogue contract (thread-private copy, like the FFMA skeleton):
//   void row32(int m, bool valid, int n0, float (&v)[32], float* red);  // 32 consecutive columns of row m; called by
//                                                                       // every lane of warps 0-3 (valid = m < M)
//   void finish(float* red);        // by all 256 threads after a __syncthreads; red = 1024 zero-initialised smem floats
//                                   // (convention: warp w of the four epilogue warps owns red[256 w .. 256 w + 255])
#pragma once
#include "common.cuh"
#include <type_traits>
#include <utility>
#include "tc_core.cuh"

namespace hopk {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128, TC_BK = 64;
template <int BN>
constexpr size_t tc_smem_bytes() { return 2 * (tc::slab_bytes(TC_BM) + tc::slab_bytes(BN)) + 1024; }

// optional vector interfaces of a loader:
//   void ld8(int i, int k, int kmax, float (&f)[8]) const;     K-contiguous operand: elements (i, k .. k+7), 0 for k >= kmax
//   void ld8mn(int k, int i, int imax, float (&f)[8]) const;   M/N-contiguous operand: elements (i .. i+7, k), 0 for i >= imax
template <class T, class = void> struct has_ld8 : std::false_type {};
template <class T>
struct has_ld8<T, std::void_t<decltype(std::declval<const T&>().ld8(0, 0, 0, std::declval<float (&)[8]>()))>> : std::true_type {};
template <class T, class = void> struct has_ld8mn : std::false_type {};
template <class T>
struct has_ld8mn<T, std::void_t<decltype(std::declval<const T&>().ld8mn(0, 0, 0, std::declval<float (&)[8]>()))>> : std::true_type {};

// Staging is split into a LOAD phase (global -> registers) and a STORE phase (registers -> bf16 -> swizzled smem) so
// that all global loads of a slab are in flight together (one memory latency per slab instead of one per 16-byte
// chunk), and so that the loads of slab k+1 can be issued before the barrier / UMMAs of slab k.
template <int ROWS>
struct SlabRegs { float f[(ROWS * 8) / TC_THREADS][8]; };

// B operand whose N index is contiguous in memory: staged as [64 K-rows][BN columns] slabs and read MN-major
template <int BN, class L>
__device__ __forceinline__ void tc_load_slab_mn(SlabRegs<BN>& r, const L& ld, int n0, int N, int k0, int kmax)
{
#pragma unroll
    for (int it = 0; it < (TC_BK * (BN / 8)) / TC_THREADS; ++it) {
        int idx = threadIdx.x + it * TC_THREADS;
        int ch = idx % (BN / 8), krow = idx / (BN / 8);
        if (k0 + krow < kmax) ld.ld8mn(k0 + krow, n0 + ch * 8, N, r.f[it]);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r.f[it][j] = 0.f;
        }
    }
}
template <int BN>
__device__ __forceinline__ void tc_store_slab_mn(uint8_t* slabs, const SlabRegs<BN>& r)
{
#pragma unroll
    for (int it = 0; it < (TC_BK * (BN / 8)) / TC_THREADS; ++it) {
        int idx = threadIdx.x + it * TC_THREADS;
        int ch = idx % (BN / 8), krow = idx / (BN / 8);
        tc::slab_store8(slabs + (ch >> 3) * tc::slab_bytes(TC_BK), krow, ch & 7, r.f[it]);
    }
}

// item -> (row, chunk): lane order follows the loader's contiguous index so global reads coalesce
template <int ROWS, class L>
__device__ __forceinline__ void tc_item(int it, int& row, int& ch)
{
    int idx = threadIdx.x + it * TC_THREADS;
    if (has_ld8<L>::value || L::kFast) { ch = idx & 7; row = idx >> 3; } else { row = idx % ROWS; ch = idx / ROWS; }
}

template <int ROWS, class L>
__device__ __forceinline__ void tc_load_slab(SlabRegs<ROWS>& r, const L& ld, int row0, int nrows_valid, int k0, int kmax)
{
#pragma unroll
    for (int it = 0; it < (ROWS * 8) / TC_THREADS; ++it) {
        int row, ch;
        tc_item<ROWS, L>(it, row, ch);
        const int i = row0 + row;
        if constexpr (has_ld8<L>::value) {
            if (i < nrows_valid) ld.ld8(i, k0 + ch * 8, kmax, r.f[it]);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) r.f[it][j] = 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int k = k0 + ch * 8 + j;
                r.f[it][j] = (i < nrows_valid && k < kmax) ? ld(i, k) : 0.f;
            }
        }
    }
}
template <int ROWS, class L>
__device__ __forceinline__ void tc_store_slab(uint8_t* slab, const SlabRegs<ROWS>& r)
{
#pragma unroll
    for (int it = 0; it < (ROWS * 8) / TC_THREADS; ++it) {
        int row, ch;
        tc_item<ROWS, L>(it, row, ch);
        tc::slab_store8(slab, row, ch, r.f[it]);
    }
}

// sum over the 32 lanes of v[j], result for column j = lane delivered to that lane (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32])
{
    const int lane = threadIdx.x & 31;
#define HOPK_TS_STEP(OFF)                                                          \
    {                                                                              \
        const bool upper = lane & OFF;                                             \
        _Pragma("unroll") for (int j = 0; j < OFF; ++j) {                          \
            float send = upper ? v[j] : v[j + OFF];                                \
            float keep = upper ? v[j + OFF] : v[j];                                \
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);                 \
        }                                                                          \
    }
    HOPK_TS_STEP(16) HOPK_TS_STEP(8) HOPK_TS_STEP(4) HOPK_TS_STEP(2) HOPK_TS_STEP(1)
#undef HOPK_TS_STEP
    return v[0];
}

template <int BN, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(int M, int N, int K, int k_per_split, ALoad aload, BLoad bload, Epi epi)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[1024];                      // 4 lane-quarter warps x 256: per-warp slots, combined in a fixed order (deterministic)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr uint32_t A_BYTES = tc::slab_bytes(TC_BM), B_BYTES = tc::slab_bytes(BN), STAGE = A_BYTES + B_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);

    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
    }
    for (int i = tid; i < 1024; i += TC_THREADS) red[i] = 0.f;
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, BN);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_smem;
    constexpr bool B_MN = !BLoad::kFast && has_ld8mn<BLoad>::value;       // N-contiguous B: no transpose, MN-major view
    constexpr uint32_t idesc = tc::idesc_bf16(TC_BM, BN, 0, B_MN ? 1 : 0);

    const int nslabs = k_end > k_begin ? (k_end - k_begin + TC_BK - 1) / TC_BK : 0;
    SlabRegs<TC_BM> ra;
    SlabRegs<BN> rb;
    auto load_regs = [&](int ks) {
        tc_load_slab<TC_BM>(ra, aload, m0, M, k_begin + ks * TC_BK, k_end);
        if constexpr (B_MN) tc_load_slab_mn<BN>(rb, bload, n0, N, k_begin + ks * TC_BK, k_end);
        else tc_load_slab<BN>(rb, bload, n0, N, k_begin + ks * TC_BK, k_end);
    };
    if (nslabs > 0) load_regs(0);
    for (int ks = 0; ks < nslabs; ++ks) {
        const int buf = ks & 1;
        if (ks >= 2) tc::mbar_wait(&bars[buf], ((ks >> 1) - 1) & 1);      // UMMAs that read this stage are done
        uint8_t* sa = smem + buf * STAGE;
        uint8_t* sb = sa + A_BYTES;
        tc_store_slab<TC_BM, ALoad>(sa, ra);
        if constexpr (B_MN) tc_store_slab_mn<BN>(sb, rb);
        else tc_store_slab<BN, BLoad>(sb, rb);
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t a_addr = tc::smem_u32(sa), b_addr = tc::smem_u32(sb);
#pragma unroll
            for (int j = 0; j < TC_BK / 16; ++j)
                tc::mma_bf16(tmem, tc::desc_kmajor(a_addr, j),
                             B_MN ? tc::desc_mnmajor(b_addr, tc::slab_bytes(TC_BK), j) : tc::desc_kmajor(b_addr, j), idesc, (ks | j) != 0);
            tc::mma_commit(&bars[buf]);
        }
        // next slab's global loads fly under the UMMAs (issued after the proxy fence: MEMBAR would wait for them)
        if (ks + 1 < nslabs) load_regs(ks + 1);
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
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            if (nslabs > 0) tc::tmem_ld32(lane_addr + c * 32, v);
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            e.row32(m, m < M, n0 + c * 32, v, red);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    e.finish(red);
    if (warp == 0) tc::tmem_dealloc(tmem, BN);
}

template <int BN, class AL, class BL, class EP>
static cudaError_t launch_gemm_tc(int M, int N, int K, int splits, AL a, BL b, EP e, cudaStream_t st)
{
    auto kern = gemm_tc_kernel<BN, AL, BL, EP>;
    {
        cudaError_t err = configure_smem_once((const void*)kern, tc_smem_bytes<BN>());      // once per (device, instantiation)
        if (err != cudaSuccess) return err;
    }
    int kper = K;
    if (splits > 1) { kper = (((K + splits - 1) / splits + TC_BK - 1) / TC_BK) * TC_BK; splits = (K + kper - 1) / kper; }
    if (splits < 1) splits = 1;
    dim3 grid((N + BN - 1) / BN, (M + TC_BM - 1) / TC_BM, splits);
    kern<<<grid, TC_THREADS, tc_smem_bytes<BN>(), st>>>(M, N, K, kper, a, b, e);
    return cudaGetLastError();
}

}  // namespace hopk
