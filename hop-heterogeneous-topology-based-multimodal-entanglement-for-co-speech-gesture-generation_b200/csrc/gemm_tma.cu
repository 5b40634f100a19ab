// gemm_tma.cu -- dense bf16 GEMM on tcgen05 with TMA tensor copies: the workhorse behind every plain dense contraction of
// the step (projections of the reprogramming layer HOP.py:276-285, the text-prototype mapping GEMM HOP.py:200, the beat MLP
// HOP.py:130-134, the align layer HOP.py:202-203, the input / weight-gradient GEMMs of the GRU decoder HOP.py:166-167).
//
//     C[M][N] (+)= sum_k A(m, k) * B(n, k)  (+ bias[n]) -> act
//
// Operands are bf16 matrices in global memory, each either "K-major" (the contraction index is contiguous: A is [M][K],
// B is [N][K], i.e. x @ W^T with PyTorch's Linear weight) or "MN-major" (the contraction index is the row: A is [K][M],
// B is [K][N]) -- so dX = dY @ W (B MN-major) and dW = dY^T @ X (both MN-major) need no transposed copies.
//
// Kernel anatomy (persistent, warp specialised, one CTA per SM):
//   warp 0   producer: cp.async.bulk.tensor.2d (TMA, 128B swizzle) fills a ring of STAGES x (A 16 KB | B 16 / 24 / 32 KB)
//            slabs, completion by mbarrier complete_tx
//   warp 1   MMA issuer: one lane issues tcgen05.mma (M 128, N = BN, K 16, 4 per 64-wide K block) into one of two TMEM
//            accumulators; tcgen05.commit releases the ring slot / publishes the accumulator
//   warps 2-9  epilogue (two warps per TMEM lane quarter, interleaved 32-column chunks): tcgen05.ld (row per thread), bias /
//            residual / activation / derivative mask / conversion, then a TMA tensor store out of two alternating swizzled
//            staging boxes per warp (per-thread stores when C's rows are not 16-byte multiples; vector reductions for split-K)
// Launched with programmatic stream serialisation: the prologue overlaps the preceding kernel's tail.
// The second accumulator lets the MMAs of tile i+1 run under the epilogue of tile i.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "tc_core.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

constexpr int GT_BM = 128, GT_BK = 64;
constexpr int GT_EPI_WARPS = 8;                         // 2 per TMEM lane quarter (16 were measured: no gain)
constexpr int GT_THREADS = 64 + 32 * GT_EPI_WARPS;      // producer warp, MMA warp, epilogue warps

template <int BN> constexpr int gt_stages() { return BN == 128 ? 6 : 4; }
template <int BN> constexpr uint32_t gt_stage_bytes() { return tc::slab_bytes(GT_BM) + tc::slab_bytes(BN); }
constexpr uint32_t GT_STG_BYTES = 4096;                 // per epilogue warp: two 32-row x 64-byte store boxes (32 bf16 / 16 fp32 columns)
template <int BN> constexpr size_t gt_smem_bytes() { return (size_t)gt_stages<BN>() * gt_stage_bytes<BN>() + (size_t)GT_EPI_WARPS * GT_STG_BYTES + 1024 + 256; }

struct GtArgs {
    void* C; const float* bias; const void* addend; const void* mask;
    int bias_row, mask_bf16, mask_gelu;                            // bias indexed by the output row m instead of the column n; mask stored as bf16
    int a_kshift, b_kshift;                             // added to the contraction (row) coordinate of an MN-major operand; rows
                                                        // that fall outside the matrix read as zero (TMA out-of-bounds fill)
    int M, N, K;
    long ldc;
    int a_mn, b_mn;                                     // 1: operand is MN-major ([K][M] / [K][N] row-major)
    int out_bf16, accumulate, act, splits, kper;        // act: 0 none, 1 relu, 2 leaky relu (slope), 3 gelu (erf)
    float slope;
    int tiles_m, tiles_n;
    int tma_store;                                      // the epilogue writes C through tmC (32-row x 64-byte boxes) instead of per-thread stores
};

__device__ __forceinline__ float gt_act(float x, int act, float slope)
{
    if (act == 1) return fmaxf(x, 0.f);
    if (act == 2) return x > 0.f ? x : slope * x;
    if (act == 3) return gelu_fast(x);
    return x;
}

// The operand majors are template parameters: the single MMA-issuing thread is the busiest warp of the kernel (ncu, round 2:
// ~180 instructions per K block with run-time layout branches against 512 tensor-pipe cycles), so its loop is straight-line:
// one 64-bit add per descriptor, four tcgen05.mma, one commit.
// (A cluster-pair variant with a multicast B tile was measured too: no gain at cluster size 2 -- TMA multicast only
// de-duplicates L2 reads for larger clusters -- and it was removed again.  Knock-out timings at 4352 x 3072 x 768: fixed
// cost 12 us, + main loop 11 us, + TMEM reads of the epilogue 6 us, + its stores 9 us: the epilogue of a tile is as long as
// its main loop and competes with the next tile's MMAs for TMEM bandwidth; per-thread stores, a shared-memory transpose and
// the TMA store all time the same with ONE staging box per warp; alternating two boxes
// (the store of box i reads its buffer while box i + 1 is filled) gained 12 % at 4352 x 3072 x 768.  Issuing the next chunk's
// tcgen05.ld before processing the current one made the kernel 30 % slower: more TMEM reads in flight beside the MMAs.)
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const GtArgs g)
{
    constexpr int STAGES = gt_stages<BN>();
    constexpr uint32_t A_BYTES = tc::slab_bytes(GT_BM), B_BYTES = tc::slab_bytes(BN), STAGE = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* staging = smem + (size_t)STAGES * STAGE;       // GT_EPI_WARPS x GT_STG_BYTES, 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + (size_t)GT_EPI_WARPS * GT_STG_BYTES);
    uint64_t* full = bars;                       // [STAGES] TMA -> MMA
    uint64_t* empty = bars + STAGES;             // [STAGES] MMA -> TMA
    uint64_t* acc_full = bars + 2 * STAGES;      // [2]      MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2; // [2]      epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], GT_EPI_WARPS); }
        tc::fence_barrier_init();
    }
    constexpr uint32_t TMEM_COLS = BN <= 128 ? 256 : 512;  // two accumulators of BN columns (allocation sizes are powers of two)
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    // programmatic dependent launch: everything above overlaps the tail of the preceding kernel in the stream; global
    // memory is only touched after that kernel has completed.  Our own dependents may start their prologue right away.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int per_split = g.tiles_m * g.tiles_n;
    const int ntiles = per_split * g.splits;
    const int w0 = (int)blockIdx.x, wstride = (int)gridDim.x;
    auto tile_m0 = [&](int t2) { return (t2 / g.tiles_n) * GT_BM; };
    if (warp == 0) {
        // ===================================================== producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = w0; tile < ntiles; tile += wstride) {
                const int sp = tile / per_split, t2 = tile - sp * per_split;
                const int m0 = tile_m0(t2), n0 = (t2 % g.tiles_n) * BN;
                const int k_begin = sp * g.kper, k_end = min(g.K, k_begin + g.kper);
                for (int k = k_begin; k < k_end; k += GT_BK) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * STAGE;
                    uint8_t* sb = sa + A_BYTES;
                    tc::mbar_expect_tx(&full[stage], STAGE);
                    if (!A_MN) tc::tma_load_2d(sa, &tmA, k, m0, &full[stage]);                    // box {64 k, 128 rows}
                    else {
#pragma unroll
                        for (int s = 0; s < GT_BM / 64; ++s) tc::tma_load_2d(sa + s * tc::slab_bytes(GT_BK), &tmA, m0 + 64 * s, k + g.a_kshift, &full[stage]);   // box {64 m, 64 k-rows}
                    }
                    if (!B_MN) tc::tma_load_2d(sb, &tmB, k, n0, &full[stage]);
                    else {
#pragma unroll
                        for (int s = 0; s < BN / 64; ++s) tc::tma_load_2d(sb + s * tc::slab_bytes(GT_BK), &tmB, n0 + 64 * s, k + g.b_kshift, &full[stage]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = tc::idesc_bf16(GT_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            // descriptors of stage 0; a later stage / K = 16 step only moves the 14-bit start-address field (bytes >> 4)
            const uint32_t s0 = tc::smem_u32(smem);
            const uint64_t dA0 = A_MN ? tc::desc_mnmajor(s0, tc::slab_bytes(GT_BK), 0) : tc::desc_kmajor(s0, 0);
            const uint64_t dB0 = B_MN ? tc::desc_mnmajor(s0 + A_BYTES, tc::slab_bytes(GT_BK), 0) : tc::desc_kmajor(s0 + A_BYTES, 0);
            constexpr uint64_t STEP_A = (A_MN ? 2048u : 32u) >> 4, STEP_B = (B_MN ? 2048u : 32u) >> 4, STEP_STAGE = STAGE >> 4;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = w0; tile < ntiles; tile += wstride) {
                const int sp = tile / per_split;
                const int k_begin = sp * g.kper, k_end = min(g.K, k_begin + g.kper);
                tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc::fence_after_sync();
                const uint32_t d = tmem + acc * BN;
                uint32_t accumulate = 0;
                for (int k = k_begin; k < k_end; k += GT_BK) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::fence_after_sync();
                    const uint64_t da = dA0 + (uint64_t)stage * STEP_STAGE, db = dB0 + (uint64_t)stage * STEP_STAGE;
                    tc::mma_bf16(d, da, db, idesc, accumulate != 0);
                    tc::mma_bf16(d, da + STEP_A, db + STEP_B, idesc, true);
                    tc::mma_bf16(d, da + 2 * STEP_A, db + 2 * STEP_B, idesc, true);
                    tc::mma_bf16(d, da + 3 * STEP_A, db + 3 * STEP_B, idesc, true);
                    accumulate = 1;
                    tc::mma_commit(&empty[stage]);                 // ring slot free once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc::mma_commit(&acc_full[acc]);                    // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue: warps 2..; warp w drains TMEM lane quarter w & 3 (the
        // quarter a warp may access) and, of the tile's 32-column chunks, every (GT_EPI_WARPS / 4)-th one starting at (w - 2) >> 2
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        uint8_t* stg = staging + (size_t)(warp - 2) * GT_STG_BYTES;
        uint32_t box = 0;                                  // TMA store boxes issued by this warp (alternates the two staging halves)
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = w0; tile < ntiles; tile += wstride) {
            const int t2 = tile % per_split;
            const bool first_split = tile < per_split;                     // bias / addend enter once, with K-split 0
            const int m0 = tile_m0(t2), n0 = (t2 % g.tiles_n) * BN;
            const int m = m0 + row;
            tc::mbar_wait(&acc_full[acc], acc_phase);
            tc::fence_after_sync();
            const uint32_t src = tmem + acc * BN + lane_off;
#pragma unroll 1
            for (int c = chalf; c < BN / 32; c += GT_EPI_WARPS / 4) {
                const int nb = n0 + c * 32;
                if (nb >= g.N) break;                              // uniform across the warp
                float v[32];
                tc::tmem_ld32(src + c * 32, v);
                const bool fullw = nb + 32 <= g.N;
                if (m < g.M) {
                    if (g.bias && first_split) {
                        if (g.bias_row) {
                            const float bm = __ldg(g.bias + m);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += bm;
                        } else if (fullw && ((reinterpret_cast<uintptr_t>(g.bias + nb) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + nb + j));
                                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (fullw || nb + j < g.N) v[j] += __ldg(g.bias + nb + j);
                        }
                    }
                    if (g.addend && first_split) {                                // fp32 residual / accumulate-from tensor with C's layout
                        const float* ad = reinterpret_cast<const float*>(g.addend) + (size_t)m * g.ldc + nb;
                        if (fullw && ((reinterpret_cast<uintptr_t>(ad) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 aa = __ldg(reinterpret_cast<const float4*>(ad + j));
                                v[j] += aa.x; v[j + 1] += aa.y; v[j + 2] += aa.z; v[j + 3] += aa.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (fullw || nb + j < g.N) v[j] += __ldg(ad + j);
                        }
                    }
                    if (g.act) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gt_act(v[j], g.act, g.slope);
                    }
                    if (g.mask) {                                  // out *= mask > 0 ? 1 : slope  (gradient of a fused (Leaky)ReLU)
                        if (g.mask_gelu) {                         // out *= gelu'(mask): the mask is the saved bf16 pre-activation
                            const __nv_bfloat16* mk = reinterpret_cast<const __nv_bfloat16*>(g.mask) + (size_t)m * g.ldc + nb;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (fullw || nb + j < g.N) {
                                const float xx = __bfloat162float(mk[j]);
                                v[j] *= gelu_grad_fast(xx);
                            }
                        } else if (g.mask_bf16) {
                            const __nv_bfloat16* mk = reinterpret_cast<const __nv_bfloat16*>(g.mask) + (size_t)m * g.ldc + nb;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (fullw || nb + j < g.N) v[j] = __bfloat162float(mk[j]) > 0.f ? v[j] : g.slope * v[j];
                        } else {
                            const float* mk = reinterpret_cast<const float*>(g.mask) + (size_t)m * g.ldc + nb;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (fullw || nb + j < g.N) v[j] = __ldg(mk + j) > 0.f ? v[j] : g.slope * v[j];
                        }
                    }
                }
                if (g.tma_store) {
                    // row-per-thread registers -> swizzled staging box of 32 rows x 64 bytes (32 bf16 or 16 fp32 columns) -> one TMA
                    // store (full lines, clipped at M / N).  Two boxes per warp alternate: the store of box i reads its buffer
                    // while box i + 1 is being filled (a single buffer serialised every chunk on the TMA read latency).
                    auto put_box = [&](const uint4 (&u)[4], int col) {
                        uint8_t* buf = stg + (box & 1) * (GT_STG_BYTES / 2);
                        ++box;
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the box before last has left
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = u[q];
                        tc::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(&tmC), "r"(tc::smem_u32(buf)), "r"(col), "r"(m0 + quarter * 32) : "memory");
                            tc::bulk_commit();
                        }
                    };
                    if (g.out_bf16) {
                        uint4 u[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            u[q].x = tc::pack_bf16x2(v[8 * q], v[8 * q + 1]); u[q].y = tc::pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                            u[q].z = tc::pack_bf16x2(v[8 * q + 4], v[8 * q + 5]); u[q].w = tc::pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                        }
                        put_box(u, nb);
                    } else {
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            if (nb + 16 * hh >= g.N) break;         // uniform
                            uint4 u[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                u[q].x = __float_as_uint(v[16 * hh + 4 * q]); u[q].y = __float_as_uint(v[16 * hh + 4 * q + 1]);
                                u[q].z = __float_as_uint(v[16 * hh + 4 * q + 2]); u[q].w = __float_as_uint(v[16 * hh + 4 * q + 3]);
                            }
                            put_box(u, nb + 16 * hh);
                        }
                    }
                } else if (m < g.M) {
                    if (g.out_bf16) {
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.C) + (size_t)m * g.ldc + nb;
                        if (fullw && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 q;
                                q.x = tc::pack_bf16x2(v[j], v[j + 1]); q.y = tc::pack_bf16x2(v[j + 2], v[j + 3]);
                                q.z = tc::pack_bf16x2(v[j + 4], v[j + 5]); q.w = tc::pack_bf16x2(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(dst + j) = q;
                            }
                        } else {
                            for (int j = 0; j < 32 && nb + j < g.N; ++j) dst[j] = __float2bfloat16_rn(v[j]);
                        }
                    } else {
                        float* dst = reinterpret_cast<float*>(g.C) + (size_t)m * g.ldc + nb;
                        if (g.splits > 1 || g.accumulate) {        // split-K partials / accumulation: vector reductions
                            if (fullw && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                                                 "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                            } else {
                                for (int j = 0; j < 32 && nb + j < g.N; ++j) atomicAdd(dst + j, v[j]);
                            }
                        } else if (fullw && ((reinterpret_cast<uintptr_t>(dst) & 31) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) tc::stg256(dst + j, v + j);
                        } else if (fullw && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {      // rows that are only 8-byte aligned
#pragma unroll
                            for (int j = 0; j < 32; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
                        } else {
                            for (int j = 0; j < 32 && nb + j < g.N; ++j) dst[j] = v[j];
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (g.tma_store && lane == 0) tc::bulk_wait_read();        // shared memory stays valid until the last box has been read
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ small helper kernels
// fp32 -> bf16 copy of a (rows x cols) matrix with independent leading dimensions (cols padded with zeros up to cols_out)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long rows, int cols, long lds,
                                 int cols_out, long ldd, int relu)
{
    const long n8 = rows * (cols_out / 8);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (cols_out / 8); const int c = (int)(i - r * (cols_out / 8)) * 8;
        float f[8];
        const float* s = src + r * lds + c;
        if (c + 8 <= cols && ((lds | c) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = c + j < cols ? __ldg(s + j) : 0.f;
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint4 q;
        q.x = tc::pack_bf16x2(f[0], f[1]); q.y = tc::pack_bf16x2(f[2], f[3]); q.z = tc::pack_bf16x2(f[4], f[5]); q.w = tc::pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(dst + r * ldd + c) = q;
    }
}

// overlapping windows of a 1-D signal as a bf16 matrix: out[(b*nwin + k)][c] = x[b][k*hop + c]  (Tensor.unfold + cast)
__global__ void unfold_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int nwin, int win, int hop,
                                   long ldx, long ldo)
{
    const long n = (long)B * nwin * (ldo / 8);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (ldo / 8); const int c = (int)(i - r * (ldo / 8)) * 8;
        const int b = (int)(r / nwin), k = (int)(r % nwin);
        const float* s = x + (long)b * ldx + (long)k * hop + c;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = c + j < win ? __ldg(s + j) : 0.f;
        uint4 q;
        q.x = tc::pack_bf16x2(f[0], f[1]); q.y = tc::pack_bf16x2(f[2], f[3]); q.z = tc::pack_bf16x2(f[4], f[5]); q.w = tc::pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(out + r * ldo + c) = q;
    }
}

// column sums of a bf16 or fp32 (rows x cols) matrix into fp32 out[cols] (bias gradients); out must be zeroed by the caller.
// A thread owns 8 (bf16) / 4 (fp32) consecutive columns = one 16-byte load per row; blockIdx.y strides over row blocks.
template <class T>
__global__ void colsum_kernel(const T* __restrict__ src, float* __restrict__ out, long rows, int cols, long ld, long rows_per_block)
{
    constexpr int W = 16 / sizeof(T);
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) * W;
    if (c >= cols) return;
    const long r0 = (long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc[W];
#pragma unroll
    for (int j = 0; j < W; ++j) acc[j] = 0.f;
    const bool vec = c + W <= cols && (ld % W) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    auto add16 = [&](const uint4& u) {
        if constexpr (sizeof(T) == 2) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[2 * j] += __low2float(h[j]); acc[2 * j + 1] += __high2float(h[j]); }
        } else {
            acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y); acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
        }
    };
    long r = r0;
    if (vec) {
        for (; r + 8 <= r1; r += 8) {                              // eight independent 16-byte loads in flight per thread
            uint4 u[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(src + (r + k) * ld + c));
#pragma unroll
            for (int k = 0; k < 8; ++k) add16(u[k]);
        }
        for (; r < r1; ++r) add16(__ldg(reinterpret_cast<const uint4*>(src + r * ld + c)));
    } else {
        for (; r < r1; ++r) {
            const T* p = src + r * ld + c;
#pragma unroll
            for (int j = 0; j < W; ++j) if (c + j < cols) acc[j] += (float)p[j];
        }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) if (c + j < cols) atomicAdd(out + c + j, acc[j]);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// row-major matrix (rows x cols, pitch ld elements of elt_bytes = 2: bf16, 4: fp32) as a 2-D tensor map with {box_cols, box_rows}
// boxes and the 128B / 64B swizzle
int make_tensor_map_2d(void* tmap, const void* base, int elt_bytes, long rows, long cols, long ld, int box_cols, int box_rows, int swizzle_bytes)
{
    CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(tmap);
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(3, "cuTensorMapEncodeTiled", "driver entry point not found");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elt_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[96];
        snprintf(msg, sizeof(msg), "CUresult %d (rows %ld cols %ld ld %ld)", (int)r, rows, cols, ld);
        return fail(3, "cuTensorMapEncodeTiled failed:", msg);
    }
    return 0;
}
int make_tensor_map_bf16(void* tmap, const void* base, long rows, long cols, long ld, int box_rows)
{
    return make_tensor_map_2d(tmap, base, 2, rows, cols, ld, 64, box_rows, 128);
}

static int make_map(CUtensorMap* tm, const void* base, long rows, long cols, long ld, int box_rows)
{
    return make_tensor_map_bf16(tm, base, rows, cols, ld, box_rows);
}

int gemm_bf16_launch(const void* A, const void* B, void* C, const float* bias, const void* addend, int M, int N, int K, long lda,
                     long ldb, long ldc, int a_mn, int b_mn, int out_bf16, int accumulate, int act, float slope, int splits,
                     cudaStream_t st, const void* mask, int a_kshift, int b_kshift, int bias_row, int mask_bf16, int mask_gelu)
{
    HOPK_REQUIRE((a_kshift == 0 || a_mn) && (b_kshift == 0 || b_mn), "a row shift needs an MN-major operand");
    HOPK_REQUIRE(!(mask && (accumulate || splits > 1)), "a mask cannot be combined with accumulation / split-K");
    HOPK_REQUIRE(M > 0 && N > 0 && K > 0, "gemm sizes");
    HOPK_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "operand leading dimensions must be multiples of 8 bf16 (16 bytes, TMA)");
    HOPK_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0, "operands must be 16-byte aligned");
    HOPK_REQUIRE(!(out_bf16 && (accumulate || splits > 1)), "accumulation needs an fp32 destination");
    HOPK_REQUIRE(!(act && (accumulate || splits > 1)), "an activation cannot be combined with accumulation / split-K");
    CUtensorMap tmA, tmB, tmC;
    // K-major: matrix is (M rows x K cols), box {64 k, 128 rows}; MN-major: matrix is (K rows x M cols), box {64 m, 64 k-rows}
    if (int rc = a_mn ? make_map(&tmA, A, K, M, lda, GT_BK) : make_map(&tmA, A, M, K, lda, GT_BM)) return rc;
    // tile width: 128 x 256 tiles read less of A per flop but cost ~1.75x a 128 x 128 tile; the persistent grid runs
    // ceil(tiles / SMs) waves, so take whichever finishes first (wave quantisation decides at these sizes)
    int dev0 = 0, sms0 = 148;
    cudaGetDevice(&dev0);
    cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
    // tile width: a wider tile reads less of A per flop but costs more per tile (about 0.25 + 0.75 BN / 128 of a 128-wide
    // one); the persistent grid runs ceil(tiles / SMs) rounds, so take whichever finishes first -- wave quantisation decides
    // at these sizes (N = 768 fills 136 of 148 SMs with 192-wide tiles, 102 with 256-wide ones)
    const long sp = splits > 1 ? splits : 1;
    int BN = 128;
    double best = 1e30;
    for (int bn = 128; bn <= 256; bn += 64) {
        if (bn > 128 && N <= bn - 64) break;
        const long tiles = (long)cdiv(M, GT_BM) * cdiv(N, bn) * sp;
        const double cost = (double)cdiv(tiles, sms0) * (0.25 + 0.75 * bn / 128.0);
        if (cost < best - 1e-9) { best = cost; BN = bn; }
    }
    if (splits < 1) splits = 1;
    int kper = ((cdiv(K, splits) + GT_BK - 1) / GT_BK) * GT_BK;
    splits = cdiv(K, kper);
    const int tiles_m = cdiv(M, GT_BM), tiles_n = cdiv(N, BN);
    if (int rc = b_mn ? make_map(&tmB, B, K, N, ldb, GT_BK) : make_map(&tmB, B, N, K, ldb, BN)) return rc;
    GtArgs g;
    g.C = C; g.bias = bias; g.addend = addend; g.mask = mask; g.bias_row = bias_row; g.mask_bf16 = mask_bf16; g.mask_gelu = mask_gelu;
    g.a_kshift = a_kshift; g.b_kshift = b_kshift; g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.a_mn = a_mn; g.b_mn = b_mn;
    g.out_bf16 = out_bf16; g.accumulate = accumulate; g.act = act; g.slope = slope;
    g.splits = splits; g.kper = kper;
    g.tiles_m = tiles_m; g.tiles_n = tiles_n;
    // plain stores go through the TMA engine when C's rows are 16-byte multiples (32 x 32 boxes: 128-byte fp32 / 64-byte bf16 rows)
    const int elt = out_bf16 ? 2 : 4;
    g.tma_store = splits == 1 && !accumulate && ((size_t)ldc * elt) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    if (g.tma_store) {
        if (int rc = make_tensor_map_2d(&tmC, C, elt, M, N, ldc, out_bf16 ? 32 : 16, 32, 64)) return rc;      // 32 rows x 64 bytes
    } else tmC = tmA;
    if (splits > 1 && !accumulate) HOPK_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
    const long ntiles = (long)tiles_m * tiles_n * splits;
    const int grid = (int)(ntiles < sms0 ? ntiles : sms0);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GT_THREADS); cfg.stream = st; cfg.attrs = pdl; cfg.numAttrs = 1;
#define HOPK_GT_LAUNCH(BNV, AM, BM_)                                                                                      \
    do {                                                                                                                  \
        HOPK_CUDA(configure_smem_once((const void*)gemm_tma_kernel<BNV, AM, BM_>, gt_smem_bytes<BNV>()));                  \
        cfg.dynamicSmemBytes = gt_smem_bytes<BNV>();                                                                      \
        HOPK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tma_kernel<BNV, AM, BM_>, tmA, tmB, tmC, g));                             \
    } while (0)
    const int variant = (BN == 256 ? 8 : BN == 192 ? 4 : 0) | (a_mn ? 2 : 0) | (b_mn ? 1 : 0);
    switch (variant) {
        case 0: HOPK_GT_LAUNCH(128, false, false); break;
        case 1: HOPK_GT_LAUNCH(128, false, true); break;
        case 2: HOPK_GT_LAUNCH(128, true, false); break;
        case 3: HOPK_GT_LAUNCH(128, true, true); break;
        case 4: HOPK_GT_LAUNCH(192, false, false); break;
        case 5: HOPK_GT_LAUNCH(192, false, true); break;
        case 6: HOPK_GT_LAUNCH(192, true, false); break;
        case 7: HOPK_GT_LAUNCH(192, true, true); break;
        case 8: HOPK_GT_LAUNCH(256, false, false); break;
        case 9: HOPK_GT_LAUNCH(256, false, true); break;
        case 10: HOPK_GT_LAUNCH(256, true, false); break;
        default: HOPK_GT_LAUNCH(256, true, true); break;
    }
#undef HOPK_GT_LAUNCH
    HOPK_LAUNCH_CHECK("gemm_tma");
    return 0;
}

}  // namespace hopk

using namespace hopk;

extern "C" int hopk_gemm_bf16(const void* A, const void* B, void* C, const float* bias, const float* addend, const void* mask,
                              int M, int N, int K, long lda, long ldb, long ldc, int flags, float slope, int splits, void* stream)
{
    const int a_mn = (flags & HOPK_GEMM_A_MN) ? 1 : 0, b_mn = (flags & HOPK_GEMM_B_MN) ? 1 : 0;
    const int act = (flags & HOPK_GEMM_RELU) ? 1 : (flags & HOPK_GEMM_LEAKY) ? 2 : (flags & HOPK_GEMM_GELU) ? 3 : 0;
    return gemm_bf16_launch(A, B, C, bias, addend, M, N, K, lda, ldb, ldc, a_mn, b_mn, (flags & HOPK_GEMM_OUT_BF16) ? 1 : 0,
                            (flags & HOPK_GEMM_ACCUMULATE) ? 1 : 0, act, slope, splits, (cudaStream_t)stream, mask, 0, 0,
                            (flags & HOPK_GEMM_BIAS_ROW) ? 1 : 0, (flags & HOPK_GEMM_MASK_BF16) ? 1 : 0,
                            (flags & HOPK_GEMM_MASK_GELU) ? 1 : 0);
}

extern "C" int hopk_cast_bf16(const float* src, void* dst, long rows, int cols, long lds, int cols_out, long ldd, int relu, void* stream)
{
    HOPK_REQUIRE(rows > 0 && cols > 0 && cols_out >= cols && cols_out % 8 == 0 && ldd % 8 == 0, "cast_bf16: cols_out and ldd must be multiples of 8");
    const long n8 = rows * (cols_out / 8);
    int blocks = (int)((n8 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, rows, cols, lds, cols_out, ldd, relu);
    HOPK_LAUNCH_CHECK("cast_bf16");
    return 0;
}

extern "C" int hopk_unfold_bf16(const float* x, void* out, int B, int nwin, int win, int hop, long ldx, long ldo, void* stream)
{
    HOPK_REQUIRE(B > 0 && nwin > 0 && win > 0 && hop > 0 && ldo % 8 == 0 && ldo >= win, "unfold_bf16 sizes (ldo multiple of 8, >= win)");
    HOPK_REQUIRE((long)(nwin - 1) * hop + win <= ldx, "unfold_bf16: windows exceed the signal length");
    const long n = (long)B * nwin * (ldo / 8);
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    unfold_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, B, nwin, win, hop, ldx, ldo);
    HOPK_LAUNCH_CHECK("unfold_bf16");
    return 0;
}

extern "C" int hopk_colsum(const void* src, float* out, long rows, int cols, long ld, int src_bf16, void* stream)
{
    HOPK_REQUIRE(rows > 0 && cols > 0, "colsum sizes");
    cudaStream_t st = (cudaStream_t)stream;
    HOPK_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
    const int W = src_bf16 ? 8 : 4;                       // columns per thread
    const int tx = cdiv(cols, W) < 64 ? 32 : 64;
    long per = rows / 128 > 16 ? (rows / 128 + 7) / 8 * 8 : 16;   // about 128 row blocks, batches of 8 rows
    dim3 grid(cdiv(cdiv(cols, W), tx), cdiv(rows, per));
    if (src_bf16) colsum_kernel<__nv_bfloat16><<<grid, tx, 0, st>>>((const __nv_bfloat16*)src, out, rows, cols, ld, per);
    else colsum_kernel<float><<<grid, tx, 0, st>>>((const float*)src, out, rows, cols, ld, per);
    HOPK_LAUNCH_CHECK("colsum");
    return 0;
}
