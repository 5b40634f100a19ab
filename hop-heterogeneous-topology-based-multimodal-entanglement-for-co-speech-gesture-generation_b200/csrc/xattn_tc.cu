// xattn_tc.cu -- reprogramming cross-attention on the 5th-generation tensor cores (tcgen05 + TMEM), dtype 1.
//
// Same contract as csrc/xattn.cu (reference model/HOP.py:289-299 and its backward) with bf16 operands and fp32
// accumulation.  Head dim E = 128 (HOP: d_keys = d_ff = 128).  A CTA owns 128 query rows of one head (the flattened
// (b, l) axis) and streams the S text prototypes in tiles of 128:
//     S_j  = Q K_j^T              UMMA 128x128x128, both operands K-major slabs, accumulator in TMEM cols [0,128)
//     P_j  = online softmax       one thread per row: tcgen05.ld of its own TMEM lane, no shuffles; dropout hash
//     O   += P_j V_j              P_j written back to smem as bf16 (K-major A), V_j read MN-major (its natural layout)
// The running output lives in TMEM cols [128,256) and is rescaled in place (tcgen05.ld/st) only when a row's maximum
// moved.  Backward = two passes that recompute P from the saved log-sum-exp (dQ: loop over S; dK/dV: loop over rows),
// every transposed operand being just an MN-major *view* of a row-major slab, so nothing is physically transposed.
#include "tc_core.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

#ifdef HOPK_DEBUG                               // progress markers in mapped host memory (nvcc -DHOPK_DEBUG; never in the release library)
__device__ volatile int* g_dbg = nullptr;
#define DBG(slot, val) do { if (g_dbg && blockIdx.x == 0 && blockIdx.y == 0) { g_dbg[slot] = (val); __threadfence_system(); } } while (0)
#else
#define DBG(slot, val) do { } while (0)
#endif

constexpr int AT = 128;                       // tile edge: query rows, prototypes per tile, head dim
constexpr uint32_t AT_SLAB = tc::slab_bytes(AT);        // 16 KB: [128 rows][64 bf16]


__device__ __forceinline__ uint32_t lowbias32_tc(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool keep_mask_tc(uint64_t seed, uint64_t idx, uint32_t thr)
{
    uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
    uint64_t q = idx >> 1;
    uint32_t q_lo = (uint32_t)q, q_hi = (uint32_t)(q >> 32);
    uint32_t h = lowbias32_tc((q_lo + lowbias32_tc(q_hi ^ s_hi)) ^ s_lo);
    uint32_t field = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
    return field >= thr;
}
// dropout on N consecutive elements starting at an EVEN flat index idx0: one hash per pair
template <int N>
__device__ __forceinline__ void dropout_run_even(float (&p)[N], uint64_t idx0, uint64_t seed, uint32_t thr, float inv_keep)
{
    const uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
    const uint64_t q0 = idx0 >> 1;
    const uint32_t lo0 = (uint32_t)q0, hi0 = (uint32_t)(q0 >> 32);
    const uint32_t in0 = lowbias32_tc(hi0 ^ s_hi), in1 = lowbias32_tc((hi0 + 1) ^ s_hi);
#pragma unroll
    for (int t = 0; t < N / 2; ++t) {
        uint32_t lo = lo0 + (uint32_t)t;
        uint32_t h = lowbias32_tc((lo + (lo < lo0 ? in1 : in0)) ^ s_lo);
        p[2 * t] = (h & 0xFFFFu) >= thr ? p[2 * t] * inv_keep : 0.f;
        p[2 * t + 1] = (h >> 16) >= thr ? p[2 * t + 1] * inv_keep : 0.f;
    }
}

// stage a [128 rows][128 cols] fp32 tile of a (rows, H, E=128) tensor (head h) as two bf16 slabs (cols 0-63 | 64-127)
__device__ __forceinline__ void stage_rows_f32(uint8_t* slabs, const float* __restrict__ src, int row0, int nrows, int H, int h)
{
#pragma unroll
    for (int it = 0; it < (AT * 16) / 256; ++it) {
        int idx = threadIdx.x + it * 256;
        int ch16 = idx & 15, row = idx >> 4;              // 16 chunks of 8 floats per row
        float f[8];
        if (row0 + row < nrows) {
            const float4* p = reinterpret_cast<const float4*>(src + ((size_t)(row0 + row) * H + h) * AT + ch16 * 8);
            float4 a = __ldg(p), b = __ldg(p + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
        tc::slab_store8(slabs + (ch16 >> 3) * AT_SLAB, row, ch16 & 7, f);
    }
}

// ------------------------------------------------------------------------------------------------ forward
// v2: prototypes streamed in tiles of 64 so that one CTA needs 81 KB of shared memory and 256 TMEM columns ->
// TWO CTAs per SM, whose stage / UMMA / softmax phases overlap each other.  Two threads per query row (warps w and
// w+4 share TMEM lane quarter w&3 and split the 64 score columns), log2-domain softmax (one FFMA + one MUFU.EX2 per
// element), dropout hash with the per-row high word hoisted.
constexpr int FS = 64;                                   // prototypes per tile
constexpr uint32_t FS_SLAB = tc::slab_bytes(FS);         // 8 KB: [64 rows][64 bf16]

__device__ __forceinline__ float ex2f(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// stage a [64 rows][128 cols] fp32 tile as two bf16 slabs of 64 rows
__device__ __forceinline__ void stage_rows64_f32(uint8_t* slabs, const float* __restrict__ src, int row0, int nrows, int H, int h)
{
#pragma unroll
    for (int it = 0; it < (FS * 16) / 256; ++it) {
        int idx = threadIdx.x + it * 256;
        int ch16 = idx & 15, row = idx >> 4;
        float f[8];
        if (row0 + row < nrows) {
            const float4* p = reinterpret_cast<const float4*>(src + ((size_t)(row0 + row) * H + h) * AT + ch16 * 8);
            float4 a = __ldg(p), b = __ldg(p + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
        tc::slab_store8(slabs + (ch16 >> 3) * FS_SLAB, row, ch16 & 7, f);
    }
}

// ------------------------------------------------------------------------------------------------ forward v3
// K / V are packed ONCE per call into bf16 slab images, one contiguous 32 KB record per (head, 64-prototype tile):
// [K e<64 | K e>=64 | V e<64 | V e>=64], byte-for-byte what the UMMA descriptors expect in shared memory.  The main
// kernel then never touches K / V with its threads: one elected thread streams the records with cp.async.bulk
// (TMA bulk copy, completion on an mbarrier via complete_tx) into a 3-stage ring, and S is double-buffered in TMEM so
// that the QK^T UMMA of tile j+1 runs while the 256 threads do the softmax of tile j.
constexpr int KV_STAGES = 3;
constexpr uint32_t KV_REC = 4 * FS_SLAB;                 // 32 KB per (head, tile)

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void xattn_pack_kv_kernel(const float* __restrict__ K, const float* __restrict__ V, uint8_t* __restrict__ pack,
                                     int S, int H, int ntiles)
{
    const int j = blockIdx.x, h = blockIdx.y;
    uint8_t* rec = pack + ((size_t)h * ntiles + j) * KV_REC;
    const int s0 = j * FS;
    for (int idx = threadIdx.x; idx < 2 * FS * 16; idx += blockDim.x) {       // 2 tensors x 64 rows x 16 chunks of 8
        int which = idx / (FS * 16), r = idx - which * FS * 16;
        int ch16 = r & 15, row = r >> 4;
        const float* src = which ? V : K;
        float f[8];
        if (s0 + row < S) {
            const float4* p4 = reinterpret_cast<const float4*>(src + ((size_t)(s0 + row) * H + h) * AT + ch16 * 8);
            float4 a = __ldg(p4), b = __ldg(p4 + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
        }
        tc::slab_store8(rec + (which * 2 + (ch16 >> 3)) * FS_SLAB, row, ch16 & 7, f);
    }
}

// Warp-specialised: 16 softmax warps (four threads per query row: warps w, w+4, w+8, w+12 share TMEM lane quarter
// w & 3 and take 16 score columns each) + 1 control warp whose lane 0 streams the K/V records (cp.async.bulk) and
// issues every UMMA.  The only synchronisation inside the loop is mbarriers (S ready, P ready, PV done) and a
// 128-thread named barrier per lane quarter for the row-maximum exchange.
constexpr int FWD3_THREADS = 512 + 32;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(FWD3_THREADS, 1)
xattn_fwd_tc3_kernel(const float* __restrict__ Q, const uint8_t* __restrict__ kvpack, float* __restrict__ O,
                     float* __restrict__ LSE, int M, int L, int H, int S, float scale, float inv_keep, uint32_t thr,
                     uint64_t seed, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[KV_STAGES], bar_s[2], bar_p[2], bar_o[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float xch[2][4][AT];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* Qs = smem; uint8_t* Ps = Qs + 2 * AT_SLAB; uint8_t* ring = Ps + 2 * AT_SLAB;      // P is double-buffered
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y, m0 = blockIdx.x * AT;
    const int ntiles = (S + FS - 1) / FS;
    const uint8_t* recs = kvpack + (size_t)h * ntiles * KV_REC;

    if (tid == 512) {
#pragma unroll
        for (int i = 0; i < KV_STAGES; ++i) tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_s[0], 1); tc::mbar_init(&bar_s[1], 1);
        tc::mbar_init(&bar_p[0], 512); tc::mbar_init(&bar_p[1], 512); tc::mbar_init(&bar_o[0], 1); tc::mbar_init(&bar_o[1], 1);
        tc::fence_barrier_init();
        for (int t = 0; t < KV_STAGES && t < ntiles; ++t) {            // fill the ring
            mbar_expect_tx(&bar_full[t], KV_REC);
            bulk_g2s(ring + t * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t]);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 256);
    if (tid < 256) stage_rows_f32(Qs, Q, m0, M, H, h);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_s0 = tmem_base_smem, tmem_o = tmem_base_smem + 128;     // S buffers at columns 0 and 64

    if (warp == 16) {
        // ------------------------------------------------------------------ control warp
        if (lane == 0) {
            constexpr uint32_t idesc_qk = tc::idesc_bf16(AT, FS, 0, 0);
            constexpr uint32_t idesc_pv = tc::idesc_bf16(AT, AT, 0, 1);
            const uint32_t qa = tc::smem_u32(Qs), pa = tc::smem_u32(Ps);
            auto issue_s = [&](int t) {                                  // S_t = Q K_t^T into TMEM buffer t & 1
                tc::mbar_wait(&bar_full[t % KV_STAGES], (t / KV_STAGES) & 1);
                tc::fence_after_sync();
                const uint32_t ka = tc::smem_u32(ring + (t % KV_STAGES) * KV_REC);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(tmem_s0 + (t & 1) * 64, tc::desc_kmajor(qa + c * AT_SLAB, k), tc::desc_kmajor(ka + c * FS_SLAB, k),
                                     idesc_qk, (c | k) != 0);
                tc::mma_commit(&bar_s[t & 1]);
            };
            issue_s(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) issue_s(j + 1);                      // runs while the softmax warps work on tile j
                if (j > 0 && j + 2 < ntiles) {                           // P V_{j-1} done -> ring stage (j-1) % 3 is free: refill
                    tc::mbar_wait(&bar_o[(j - 1) & 1], ((j - 1) >> 1) & 1);
                    const int t = j + 2;
                    mbar_expect_tx(&bar_full[t % KV_STAGES], KV_REC);
                    bulk_g2s(ring + (t % KV_STAGES) * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t % KV_STAGES]);
                }
                tc::mbar_wait(&bar_p[j & 1], (j >> 1) & 1);              // P_j written (and O rescaled) by all softmax threads
                tc::fence_after_sync();
                const uint32_t va = tc::smem_u32(ring + (j % KV_STAGES) * KV_REC + 2 * FS_SLAB);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    tc::mma_bf16(tmem_o, tc::desc_kmajor(pa + (j & 1) * AT_SLAB, t), tc::desc_mnmajor(va, FS_SLAB, t), idesc_pv,
                                 (j | t) != 0);
                tc::mma_commit(&bar_o[j & 1]);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps
        const int row = (warp & 3) * 32 + lane;
        const int quad = warp >> 2;                              // which 16 of the 64 score columns / which 32 of the 128 outputs
        const int m = m0 + row;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint64_t drop_base = m < M ? (((uint64_t)(m / L) * H + h) * (uint64_t)L + (uint64_t)(m % L)) * (uint64_t)S : 0;
        const uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
        const float sc2 = scale * 1.4426950408889634f;
        float mrow = -INFINITY, lrow = 0.f;
        for (int j = 0; j < ntiles; ++j) {
            const int s0 = j * FS;
            tc::mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
            tc::fence_after_sync();
            float v[16];
            tc::tmem_ld16(tmem_s0 + (j & 1) * 64 + lane_off + quad * 16, v);
            const int sb = s0 + quad * 16;
            if (s0 + FS > S) {                                   // ragged last tile: mask the missing prototypes
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = sb + i < S ? v[i] : -INFINITY;
            }
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
            mx *= sc2;                                           // log2-domain maximum (sc2 > 0 commutes with max)
            float (*x)[AT] = xch[j & 1];
            x[quad][row] = mx;
            named_bar_sync(1 + (warp & 3), 128);                 // the four warps that share this lane quarter
            const float mnew = fmaxf(fmaxf(mrow, mx), fmaxf(fmaxf(x[quad ^ 1][row], x[quad ^ 2][row]), x[quad ^ 3][row]));
            const float corr = ex2f(mrow - mnew);
            if (j >= 2) tc::mbar_wait(&bar_o[j & 1], ((j >> 1) - 1) & 1);    // P V_{j-2} done: P buffer j & 1 is free
            if (j > 0 && __any_sync(0xffffffffu, mnew > mrow)) {     // a row of this warp moved its maximum: rescale O
                tc::mbar_wait(&bar_o[(j - 1) & 1], ((j - 1) >> 1) & 1);      // needs P V_{j-1} complete
                tc::fence_after_sync();
                {
                    float o[32];
                    tc::tmem_ld32(tmem_o + lane_off + quad * 32, o);
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] *= corr;
                    tc::tmem_st32(tmem_o + lane_off + quad * 32, o);
                }
            }
            float rs = 0.f;
            const uint64_t idx0 = drop_base + (uint64_t)sb;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = ex2f(fmaf(v[i], sc2, -mnew));
                rs += v[i];
            }
            if (thr) {
                if ((idx0 & 1) == 0) dropout_run_even<16>(v, idx0, seed, thr, inv_keep);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = keep_mask_tc(seed, idx0 + (uint64_t)i, thr) ? v[i] * inv_keep : 0.f;
                }
            }
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = v[q8 * 8 + i];
                tc::slab_store8(Ps + (j & 1) * AT_SLAB, row, quad * 2 + q8, f);
            }
            lrow = lrow * corr + rs;
            mrow = mnew;
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&bar_p[j & 1]);
        }
        if (ntiles >= 2) tc::mbar_wait(&bar_o[(ntiles - 2) & 1], ((ntiles - 2) >> 1) & 1);
        tc::mbar_wait(&bar_o[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
        tc::fence_after_sync();
        float (*x)[AT] = xch[ntiles & 1];
        x[quad][row] = lrow;
        named_bar_sync(1 + (warp & 3), 128);
        const float ltot = x[0][row] + x[1][row] + x[2][row] + x[3][row];
        const float inv = 1.f / ltot;
        float* orow = O + ((size_t)(m < M ? m : 0) * H + h) * AT + quad * 32;
        float o[32];
        tc::tmem_ld32(tmem_o + lane_off + quad * 32, o);
        if (m < M) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(orow + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
        }
        if (m < M && quad == 0) LSE[((size_t)(m / L) * H + h) * L + (m % L)] = (mrow + log2f(ltot)) * 0.6931471805599453f;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 256);
}

// ------------------------------------------------------------------------------------------------ backward: dQ, v3
// Same organisation as the v3 forward: K / V arrive as packed bf16 records (the forward's pack is reused) through a
// 3-stage cp.async.bulk ring, S and dP are double-buffered in TMEM (2 x 64 columns each), dS is double-buffered in
// shared memory, and a control warp issues every UMMA so that  S_{j+1} = Q K_{j+1}^T,  dP_{j+1} = dO V_{j+1}^T  run while
// the 16 softmax warps (four threads per query row, 16 prototypes each) turn tile j into dS_j.
// TMEM columns: S [0,128)  dP [128,256)  dQ [256,384).
__global__ void __launch_bounds__(FWD3_THREADS, 1)
xattn_bwd_dq_tc3_kernel(const float* __restrict__ Q, const uint8_t* __restrict__ kvpack, const float* __restrict__ O,
                        const float* __restrict__ LSE, const float* __restrict__ dO, float* __restrict__ dQ,
                        float* __restrict__ delta, uint4* __restrict__ rowstat, int M, int L, int H, int S, float scale,
                        float inv_keep, uint32_t thr, uint64_t seed, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[KV_STAGES], bar_s[2], bar_p[2], bar_q[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float xch[4][AT];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* Qs = smem; uint8_t* dOs = Qs + 2 * AT_SLAB; uint8_t* dSs = dOs + 2 * AT_SLAB;   // dS: 2 buffers of [128][64]
    uint8_t* ring = dSs + 2 * AT_SLAB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y, m0 = blockIdx.x * AT;
    const int ntiles = (S + FS - 1) / FS;
    const uint8_t* recs = kvpack + (size_t)h * ntiles * KV_REC;

    if (tid == 512) {
#pragma unroll
        for (int i = 0; i < KV_STAGES; ++i) tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_s[0], 1); tc::mbar_init(&bar_s[1], 1);
        tc::mbar_init(&bar_p[0], 512); tc::mbar_init(&bar_p[1], 512); tc::mbar_init(&bar_q[0], 1); tc::mbar_init(&bar_q[1], 1);
        tc::fence_barrier_init();
        for (int t = 0; t < KV_STAGES && t < ntiles; ++t) {            // fill the ring
            mbar_expect_tx(&bar_full[t], KV_REC);
            bulk_g2s(ring + t * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t]);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 512);
    if (tid < 256) stage_rows_f32(Qs, Q, m0, M, H, h);
    else if (tid < 512) {                                              // warps 8-15 stage dO with the same item map
#pragma unroll
        for (int it = 0; it < (AT * 16) / 256; ++it) {
            int idx = (tid - 256) + it * 256;
            int ch16 = idx & 15, row = idx >> 4;
            float f[8];
            if (m0 + row < M) tc::ldg256(dO + ((size_t)(m0 + row) * H + h) * AT + ch16 * 8, f);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.f;
            }
            tc::slab_store8(dOs + (ch16 >> 3) * AT_SLAB, row, ch16 & 7, f);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_s0 = tmem_base_smem, tmem_dp0 = tmem_base_smem + 128, tmem_dq = tmem_base_smem + 256;

    if (warp == 16) {
        // ------------------------------------------------------------------ control warp
        if (lane == 0) {
            constexpr uint32_t idesc_sk = tc::idesc_bf16(AT, FS, 0, 0);
            constexpr uint32_t idesc_dq = tc::idesc_bf16(AT, AT, 0, 1);
            const uint32_t qa = tc::smem_u32(Qs), da = tc::smem_u32(dOs), sa = tc::smem_u32(dSs);
            auto issue_sdp = [&](int t) {                               // S_t, dP_t into TMEM buffer t & 1
                tc::mbar_wait(&bar_full[t % KV_STAGES], (t / KV_STAGES) & 1);
                tc::fence_after_sync();
                const uint32_t ka = tc::smem_u32(ring + (t % KV_STAGES) * KV_REC), va = ka + 2 * FS_SLAB;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(tmem_s0 + (t & 1) * 64, tc::desc_kmajor(qa + c * AT_SLAB, k), tc::desc_kmajor(ka + c * FS_SLAB, k),
                                     idesc_sk, (c | k) != 0);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(tmem_dp0 + (t & 1) * 64, tc::desc_kmajor(da + c * AT_SLAB, k), tc::desc_kmajor(va + c * FS_SLAB, k),
                                     idesc_sk, (c | k) != 0);
                tc::mma_commit(&bar_s[t & 1]);
            };
            issue_sdp(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) issue_sdp(j + 1);                    // runs while the softmax warps work on tile j
                if (j > 0 && j + 2 < ntiles) {                           // dQ UMMA of tile j-1 done -> its ring stage is free
                    tc::mbar_wait(&bar_q[(j - 1) & 1], ((j - 1) >> 1) & 1);
                    const int t = j + 2;
                    mbar_expect_tx(&bar_full[t % KV_STAGES], KV_REC);
                    bulk_g2s(ring + (t % KV_STAGES) * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t % KV_STAGES]);
                }
                tc::mbar_wait(&bar_p[j & 1], (j >> 1) & 1);              // dS_j written by all softmax threads
                tc::fence_after_sync();
                const uint32_t ka = tc::smem_u32(ring + (j % KV_STAGES) * KV_REC);
#pragma unroll
                for (int t = 0; t < 4; ++t)       // dQ[row][e] += sum_s dS[row][s] K[s][e]  (K read MN-major)
                    tc::mma_bf16(tmem_dq, tc::desc_kmajor(sa + (j & 1) * AT_SLAB, t), tc::desc_mnmajor(ka, FS_SLAB, t), idesc_dq,
                                 (j | t) != 0);
                tc::mma_commit(&bar_q[j & 1]);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps
        const int row = (warp & 3) * 32 + lane;
        const int quad = warp >> 2;                              // which 16 of the 64 prototypes / which 32 of the 128 outputs
        const int m = m0 + row;
        const bool rvalid = m < M;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const size_t li = rvalid ? ((size_t)(m / L) * H + h) * L + (m % L) : 0;
        const uint64_t drop_base = (uint64_t)li * (uint64_t)S;
        const float sc2 = scale * 1.4426950408889634f;
        // delta = rowsum(dO * O) in fp32 from global: each of the row's four threads sums 32 columns
        float part = 0.f;
        if (rvalid) {
            const float* o = O + ((size_t)m * H + h) * AT + quad * 32;
            const float* d = dO + ((size_t)m * H + h) * AT + quad * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a[8], b[8];
                tc::ldg256(o + 8 * i, a); tc::ldg256(d + 8 * i, b);
#pragma unroll
                for (int k = 0; k < 8; ++k) part = fmaf(a[k], b[k], part);
            }
        }
        xch[quad][row] = part;
        named_bar_sync(1 + (warp & 3), 128);
        const float dl = xch[0][row] + xch[1][row] + xch[2][row] + xch[3][row];
        const float lse2 = rvalid ? __ldg(LSE + li) * 1.4426950408889634f : 0.f;
        if (rvalid && quad == 0) {
            delta[li] = dl;
            // per-(head, row) record for the dK/dV pass: {lse * log2(e), delta, dropout base index}
            if (rowstat) rowstat[(size_t)h * M + m] = make_uint4(__float_as_uint(lse2), __float_as_uint(dl), (uint32_t)drop_base, (uint32_t)(drop_base >> 32));
        }
        for (int j = 0; j < ntiles; ++j) {
            const int s0 = j * FS;
            tc::mbar_wait(&bar_s[j & 1], (j >> 1) & 1);
            tc::fence_after_sync();
            float sv[16], dv[16];
            tc::tmem_ld16(tmem_s0 + (j & 1) * 64 + lane_off + quad * 16, sv);
            tc::tmem_ld16(tmem_dp0 + (j & 1) * 64 + lane_off + quad * 16, dv);
            const int sb = s0 + quad * 16;
            if (thr) {
                const uint64_t idx0 = drop_base + (uint64_t)sb;
                if ((idx0 & 1) == 0) dropout_run_even<16>(dv, idx0, seed, thr, inv_keep);
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) dv[i] = keep_mask_tc(seed, idx0 + (uint64_t)i, thr) ? dv[i] * inv_keep : 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float p = ex2f(fmaf(sv[i], sc2, -lse2));
                sv[i] = (rvalid && sb + i < S) ? p * (dv[i] - dl) * scale : 0.f;
            }
            if (j >= 2) tc::mbar_wait(&bar_q[j & 1], ((j >> 1) - 1) & 1);    // dQ UMMA of tile j-2 done: dS buffer j & 1 is free
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = sv[q8 * 8 + i];
                tc::slab_store8(dSs + (j & 1) * AT_SLAB, row, quad * 2 + q8, f);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&bar_p[j & 1]);
        }
        if (ntiles >= 2) tc::mbar_wait(&bar_q[(ntiles - 2) & 1], ((ntiles - 2) >> 1) & 1);
        tc::mbar_wait(&bar_q[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
        tc::fence_after_sync();
        float v[32];
        tc::tmem_ld32(tmem_dq + lane_off + quad * 32, v);
        if (rvalid) {
            float* qrow = dQ + ((size_t)m * H + h) * AT + quad * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) tc::stg256(qrow + i, v + i);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 512);
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV, v3
// One CTA per (128 prototypes, head, chunk of query-row tiles).  Everything is computed TRANSPOSED so that the CTA's own
// prototypes are the TMEM lanes:   S^T = K Q_i^T,  dP^T = V dO_i^T   (M = 128 prototypes, N = 64 query rows per tile),
// which makes P~^T and dS^T plain K-major A operands written one row per thread, and dV += P~^T dO_i, dK += dS^T Q_i
// read Q_i / dO_i MN-major from the very records that fed the first two products.  Q / dO are packed once per call into
// bf16 records (same format and kernel as the K / V records) and streamed through a 3-stage cp.async.bulk ring; S^T and
// dP^T are double-buffered in TMEM, so the UMMAs of tile i+1 run under the softmax arithmetic of tile i.
// TMEM columns: S^T [0,128)  dP^T [128,256)  dK [256,384)  dV [384,512).  Row chunks combine with 128-bit atomics.
__global__ void __launch_bounds__(FWD3_THREADS, 1)
xattn_bwd_dkv_tc3_kernel(const float* __restrict__ K, const float* __restrict__ V, const uint8_t* __restrict__ qdopack,
                         const uint4* __restrict__ rowstat, float* __restrict__ dK, float* __restrict__ dV, int M, int H, int S,
                         int tiles_per_chunk, float scale, float inv_keep, uint32_t thr, uint64_t seed, int use_atomics, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[KV_STAGES], bar_s[2], bar_p, bar_kv;
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* Ks = smem; uint8_t* Vs = Ks + 2 * AT_SLAB; uint8_t* Ps = Vs + 2 * AT_SLAB; uint8_t* dSs = Ps + AT_SLAB;
    uint8_t* ring = dSs + AT_SLAB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.y, s0 = blockIdx.x * AT;
    const int ntiles_all = (M + FS - 1) / FS;
    const int t_begin = blockIdx.z * tiles_per_chunk;
    const int ntiles = min(tiles_per_chunk, ntiles_all - t_begin);           // >= 1 by construction of the grid
    const uint8_t* recs = qdopack + ((size_t)h * ntiles_all + t_begin) * KV_REC;

    if (tid == 512) {
#pragma unroll
        for (int i = 0; i < KV_STAGES; ++i) tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_s[0], 1); tc::mbar_init(&bar_s[1], 1);
        tc::mbar_init(&bar_p, 512); tc::mbar_init(&bar_kv, 1);
        tc::fence_barrier_init();
        for (int t = 0; t < KV_STAGES && t < ntiles; ++t) {
            mbar_expect_tx(&bar_full[t], KV_REC);
            bulk_g2s(ring + t * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t]);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 512);
    if (tid < 256) stage_rows_f32(Ks, K, s0, S, H, h);
    else if (tid < 512) {                                              // warps 8-15 stage V with the same item map
#pragma unroll
        for (int it = 0; it < (AT * 16) / 256; ++it) {
            int idx = (tid - 256) + it * 256;
            int ch16 = idx & 15, row = idx >> 4;
            float f[8];
            if (s0 + row < S) tc::ldg256(V + ((size_t)(s0 + row) * H + h) * AT + ch16 * 8, f);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.f;
            }
            tc::slab_store8(Vs + (ch16 >> 3) * AT_SLAB, row, ch16 & 7, f);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_s0 = tmem_base_smem, tmem_dp0 = tmem_base_smem + 128, tmem_dk = tmem_base_smem + 256,
                   tmem_dv = tmem_base_smem + 384;

    if (warp == 16) {
        // ------------------------------------------------------------------ control warp
        if (lane == 0) {
            constexpr uint32_t idesc_st = tc::idesc_bf16(AT, FS, 0, 0);      // [128 protos] x [64 rows], both K-major (K = e)
            constexpr uint32_t idesc_kv = tc::idesc_bf16(AT, AT, 0, 1);      // A = P~^T / dS^T K-major (K = rows), B = dO / Q MN-major
            const uint32_t ka = tc::smem_u32(Ks), va = tc::smem_u32(Vs), pa = tc::smem_u32(Ps), sa = tc::smem_u32(dSs);
            auto issue_sdp = [&](int t) {
                tc::mbar_wait(&bar_full[t % KV_STAGES], (t / KV_STAGES) & 1);
                tc::fence_after_sync();
                const uint32_t qa = tc::smem_u32(ring + (t % KV_STAGES) * KV_REC), da = qa + 2 * FS_SLAB;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(tmem_s0 + (t & 1) * 64, tc::desc_kmajor(ka + c * AT_SLAB, k), tc::desc_kmajor(qa + c * FS_SLAB, k),
                                     idesc_st, (c | k) != 0);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(tmem_dp0 + (t & 1) * 64, tc::desc_kmajor(va + c * AT_SLAB, k), tc::desc_kmajor(da + c * FS_SLAB, k),
                                     idesc_st, (c | k) != 0);
                tc::mma_commit(&bar_s[t & 1]);
            };
            issue_sdp(0);
            for (int i = 0; i < ntiles; ++i) {
                if (i + 1 < ntiles) issue_sdp(i + 1);                    // runs under the softmax arithmetic of tile i
                if (i > 0 && i + 2 < ntiles) {                           // dK/dV UMMAs of tile i-1 done -> its ring stage is free
                    tc::mbar_wait(&bar_kv, (i - 1) & 1);
                    const int t = i + 2;
                    mbar_expect_tx(&bar_full[t % KV_STAGES], KV_REC);
                    bulk_g2s(ring + (t % KV_STAGES) * KV_REC, recs + (size_t)t * KV_REC, KV_REC, &bar_full[t % KV_STAGES]);
                }
                tc::mbar_wait(&bar_p, i & 1);                            // P~^T_i, dS^T_i written by all softmax threads
                tc::fence_after_sync();
                const uint32_t qa = tc::smem_u32(ring + (i % KV_STAGES) * KV_REC), da = qa + 2 * FS_SLAB;
#pragma unroll
                for (int t = 0; t < 4; ++t)       // dV[s][e] += sum_row P~^T[s][row] dO[row][e]
                    tc::mma_bf16(tmem_dv, tc::desc_kmajor(pa, t), tc::desc_mnmajor(da, FS_SLAB, t), idesc_kv, (i | t) != 0);
#pragma unroll
                for (int t = 0; t < 4; ++t)       // dK[s][e] += sum_row dS^T[s][row] Q[row][e]
                    tc::mma_bf16(tmem_dk, tc::desc_kmajor(sa, t), tc::desc_mnmajor(qa, FS_SLAB, t), idesc_kv, (i | t) != 0);
                tc::mma_commit(&bar_kv);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps: thread = (prototype, 16 query rows)
        const int srow = (warp & 3) * 32 + lane;                 // prototype inside the tile = TMEM lane
        const int quad = warp >> 2;
        const int s = s0 + srow;
        const bool svalid = s < S;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const float sc2 = scale * 1.4426950408889634f;
        const uint4* rs_h = rowstat + (size_t)h * M;
        for (int i = 0; i < ntiles; ++i) {
            const int mb = (t_begin + i) * FS + quad * 16;      // first of this thread's 16 query rows
            tc::mbar_wait(&bar_s[i & 1], (i >> 1) & 1);
            tc::fence_after_sync();
            float sv[16], dv[16];
            tc::tmem_ld16(tmem_s0 + (i & 1) * 64 + lane_off + quad * 16, sv);
            tc::tmem_ld16(tmem_dp0 + (i & 1) * 64 + lane_off + quad * 16, dv);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int m = mb + k;
                float pt = 0.f, ds = 0.f;
                if (svalid && m < M) {
                    const uint4 r = __ldg(rs_h + m);             // same address across the warp: one broadcast transaction
                    const float p = ex2f(fmaf(sv[k], sc2, -__uint_as_float(r.x)));
                    float d = dv[k];
                    pt = p;
                    if (thr) {
                        const uint64_t idx = (((uint64_t)r.w << 32) | (uint64_t)r.z) + (uint64_t)s;
                        const bool kp = keep_mask_tc(seed, idx, thr);
                        pt = kp ? p * inv_keep : 0.f;
                        d = kp ? d * inv_keep : 0.f;
                    }
                    ds = p * (d - __uint_as_float(r.y)) * scale;
                }
                sv[k] = pt; dv[k] = ds;
            }
            if (i > 0) tc::mbar_wait(&bar_kv, (i - 1) & 1);      // dK/dV UMMAs of tile i-1 done: P / dS buffers are free
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
                float f[8], g[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { f[k] = sv[q8 * 8 + k]; g[k] = dv[q8 * 8 + k]; }
                tc::slab_store8(Ps, srow, quad * 2 + q8, f);
                tc::slab_store8(dSs, srow, quad * 2 + q8, g);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&bar_p);
        }
        tc::mbar_wait(&bar_kv, (ntiles - 1) & 1);
        tc::fence_after_sync();
        float a[32], b[32];
        tc::tmem_ld32(tmem_dk + lane_off + quad * 32, a);
        tc::tmem_ld32(tmem_dv + lane_off + quad * 32, b);
        if (svalid) {
            float* kr = dK + ((size_t)s * H + h) * AT + quad * 32;
            float* vr = dV + ((size_t)s * H + h) * AT + quad * 32;
            if (use_atomics) {
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    atomicAdd(reinterpret_cast<float4*>(kr + k), make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]));
                    atomicAdd(reinterpret_cast<float4*>(vr + k), make_float4(b[k], b[k + 1], b[k + 2], b[k + 3]));
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; k += 8) { tc::stg256(kr + k, a + k); tc::stg256(vr + k, b + k); }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 512);
}

}  // namespace hopk
using namespace hopk;

#ifdef HOPK_DEBUG
extern "C" int hopk_debug_set(void* mapped_host_ints)
{
    int* p = (int*)mapped_host_ints;
    return cudaMemcpyToSymbol(hopk::g_dbg, &p, sizeof(p)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" size_t hopk_xattn_pack_bytes(int S, int H) { return (size_t)H * ((S + FS - 1) / FS) * KV_REC + 1024; }
// backward scratch: Q / dO records (same format, rows = B*L) followed by the per-(head, row) statistics of the dK/dV pass
static size_t qdo_pack_bytes(int M, int H) { return (size_t)H * ((M + FS - 1) / FS) * KV_REC; }
extern "C" size_t hopk_xattn_bwd_scratch_bytes(int B, int L, int H) { return qdo_pack_bytes(B * L, H) + (size_t)H * B * L * sizeof(uint4) + 1024; }

extern "C" int hopk_xattn_fwd_tc(const float* q, const float* k, const float* v, float* o, float* lse, void* kv_pack, int B,
                                 int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream)
{
    HOPK_REQUIRE(B > 0 && L > 0 && H > 0 && S > 0, "xattn sizes");
    HOPK_REQUIRE(E == 128, "tensor-core attention is specialised for head dim 128");
    HOPK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout p in [0,1)");
    HOPK_REQUIRE(kv_pack != nullptr, "kv_pack scratch (hopk_xattn_pack_bytes) required");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = B * L;
    uint32_t thr = (uint32_t)lrintf(p_drop * 65536.f);
    const float inv_keep_h = 65536.f / (65536.f - (float)thr);
    const unsigned long long* epoch = drop_epoch_ptr();
    HOPK_REQUIRE(epoch != nullptr, "dropout epoch symbol");
    // packed bf16 K/V records + bulk-copy ring + double-buffered S
    const size_t smem3 = 4 * AT_SLAB + KV_STAGES * KV_REC + 1024;
    HOPK_CUDA(configure_smem_once((const void*)xattn_fwd_tc3_kernel, smem3));
    uint8_t* pack = reinterpret_cast<uint8_t*>(((uintptr_t)kv_pack + 1023) & ~uintptr_t(1023));
    const int ntiles = cdiv(S, FS);
    xattn_pack_kv_kernel<<<dim3(ntiles, H), 256, 0, st>>>(k, v, pack, S, H, ntiles);
    HOPK_LAUNCH_CHECK("xattn_pack_kv");
    xattn_fwd_tc3_kernel<<<dim3(cdiv(M, AT), H), FWD3_THREADS, smem3, st>>>(q, pack, o, lse, M, L, H, S, 1.f / sqrtf((float)E),
                                                                   inv_keep_h, thr, seed, epoch);
    HOPK_LAUNCH_CHECK("xattn_fwd_tc3");
    return 0;
}

extern "C" int hopk_xattn_bwd_tc(const float* q, const float* k, const float* v, const float* o, const float* lse,
                                 const float* dout, float* dq, float* dk, float* dv, float* delta, void* kv_pack, void* scratch,
                                 int B, int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream)
{
    HOPK_REQUIRE(B > 0 && L > 0 && H > 0 && S > 0, "xattn sizes");
    HOPK_REQUIRE(E == 128, "tensor-core attention is specialised for head dim 128");
    HOPK_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout p in [0,1)");
    HOPK_REQUIRE(delta != nullptr, "delta scratch (B*H*L floats) required");
    HOPK_REQUIRE(kv_pack != nullptr && scratch != nullptr, "the forward's kv_pack and hopk_xattn_bwd_scratch_bytes() of scratch required");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = B * L;
    uint32_t thr = (uint32_t)lrintf(p_drop * 65536.f);
    const float inv_keep = 65536.f / (65536.f - (float)thr), scale = 1.f / sqrtf((float)E);
    const unsigned long long* epoch = drop_epoch_ptr();
    HOPK_REQUIRE(epoch != nullptr, "dropout epoch symbol");
    // K / V records packed by the forward (same K, V), bulk-copy rings
    const size_t smem3 = 6 * AT_SLAB + KV_STAGES * KV_REC + 1024;
    HOPK_CUDA(configure_smem_once((const void*)xattn_bwd_dq_tc3_kernel, smem3));
    HOPK_CUDA(configure_smem_once((const void*)xattn_bwd_dkv_tc3_kernel, smem3));
    const uint8_t* pack = reinterpret_cast<const uint8_t*>(((uintptr_t)kv_pack + 1023) & ~uintptr_t(1023));
    uint8_t* qdo = reinterpret_cast<uint8_t*>(((uintptr_t)scratch + 1023) & ~uintptr_t(1023));
    uint4* rowstat = reinterpret_cast<uint4*>(qdo + qdo_pack_bytes(M, H));
    const int mt = cdiv(M, FS);
    xattn_pack_kv_kernel<<<dim3(mt, H), 256, 0, st>>>(q, dout, qdo, M, H, mt);       // Q / dO records
    HOPK_LAUNCH_CHECK("xattn_pack_qdo");
    xattn_bwd_dq_tc3_kernel<<<dim3(cdiv(M, AT), H), FWD3_THREADS, smem3, st>>>(q, pack, o, lse, dout, dq, delta, rowstat, M, L, H,
                                                                            S, scale, inv_keep, thr, seed, epoch);
    HOPK_LAUNCH_CHECK("xattn_bwd_dq_tc3");
    const int kvt = cdiv(S, AT);
    int chunks = (2 * 148 + kvt * H / 2) / (kvt * H);                   // about two waves of CTAs
    if (chunks < 1) chunks = 1;
    if (chunks > mt) chunks = mt;
    const int per = cdiv(mt, chunks);
    chunks = cdiv(mt, per);
    if (chunks > 1) {
        HOPK_CUDA(cudaMemsetAsync(dk, 0, (size_t)S * H * E * sizeof(float), st));
        HOPK_CUDA(cudaMemsetAsync(dv, 0, (size_t)S * H * E * sizeof(float), st));
    }
    xattn_bwd_dkv_tc3_kernel<<<dim3(kvt, H, chunks), FWD3_THREADS, smem3, st>>>(k, v, qdo, rowstat, dk, dv, M, H, S, per, scale,
                                                                              inv_keep, thr, seed, chunks > 1 ? 1 : 0, epoch);
    HOPK_LAUNCH_CHECK("xattn_bwd_dkv_tc3");
    return 0;
}
