// gwnet_fused.cuh -- one kernel per Graph-WaveNet layer (forward, dtype 1, C = 64).
//
// A CTA owns a tile of G = floor(128 / V) complete (b, t) node groups (126 rows for V = 9 and for V = 42) and runs the
// whole layer on it without leaving the SM:
//   1. stage [x_t | x_{t+d}] (BatchNorm_{i-1} folded into the load) as bf16 slabs; copy the pre-packed gate weights
//   2. UMMA 128x128x128  ->  TMEM (filter | gate pre-activations)
//   3. tanh * sigmoid (two threads per row), y kept in shared memory (fp32) and written out with tanh/sigmoid
//      (backward needs them) + the skip slice
//   4. diffusion x1 = A^T y, x2 = A^T x1 per node group, entirely in shared memory (fp32), re-quantised into slabs
//   5. UMMA 128x64x192 on [y | x1 | x2]  ->  TMEM
//   6. + bias + residual (BN-folded), u written out, per-channel sum / sum^2 (shuffle transpose-reduce -> smem -> fp64 atomics)
// Reference lines fused: model/gwnet.py:186-200 (gated conv), :209-220 (skip slice), :12-14,:35-46 (gcn), :233 (residual),
// and the statistics half of :237 (BatchNorm).
#pragma once
#include "functors.cuh"

namespace hopk {

constexpr int FZ_C = 64;
constexpr uint32_t FZ_WG_BYTES = 2 * tc::slab_bytes(128);            // gate weights [128 n][128 k]: 32 KB
constexpr uint32_t FZ_WM_BYTES = 3 * tc::slab_bytes(64);             // mlp weights  [64 n][192 k]: 24 KB
constexpr uint32_t FZ_PACK_BYTES = FZ_WG_BYTES + FZ_WM_BYTES;        // per layer
constexpr int FZ_LDF = FZ_C + 4;                                     // fp32 row pitch of the y / x1 tiles in smem

// bf16 slab images of one layer's weights, byte-for-byte what the kernel copies into 1024-aligned shared memory
__global__ void fz_pack_weights_kernel(const HopkGwnetParams p, int L, uint8_t* __restrict__ pack)
{
    const int l = blockIdx.x;
    if (l >= L) return;
    uint8_t* img = pack + (size_t)l * FZ_PACK_BYTES;
    const float* wf = p.filter_w[l]; const float* wg = p.gate_w[l]; const float* wm = p.mlp_w[l];
    constexpr int C = FZ_C;
    for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {      // gate: n -> (fg, o) in blocks [16 f | 16 g], k = tap*C + c
        int n = idx >> 7, k = idx & 127;
        int fg = (n >> 4) & 1, o = (n >> 5) * 16 + (n & 15);
        int tap = k >= C, c = k - tap * C;
        float v = (fg ? wg : wf)[(size_t)o * 2 * C + 2 * c + tap];
        int slab = k >> 6, col = k & 63;
        uint32_t off = slab * tc::slab_bytes(128) + tc::slab_chunk_off(n, col >> 3) + (col & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
    }
    for (int idx = threadIdx.x; idx < 64 * 192; idx += blockDim.x) {       // mlp: [o][k], k = seg*C + c (native order)
        int n = idx / 192, k = idx - n * 192;
        float v = wm[(size_t)n * 192 + k];
        int slab = k >> 6, col = k & 63;
        uint32_t off = FZ_WG_BYTES + slab * tc::slab_bytes(64) + tc::slab_chunk_off(n, col >> 3) + (col & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(v);
    }
}

// bf16 slab image of the block-diagonal diffusion operator of one 128-row tile: BD[(g, w)][(g, v)] = A[v][w]
// (x1 = BD . y applies A^T inside every complete (b, t) node group of the tile; gwnet.py:12-14 on the rows layout)
constexpr uint32_t FZ_BD_BYTES = 2 * tc::slab_bytes(128);            // [128 rows][128 k]: 32 KB
__global__ void fz_pack_bd_kernel(const float* __restrict__ A, int V, uint8_t* __restrict__ img)
{
    const int G = 128 / V;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * 128; idx += gridDim.x * blockDim.x) {
        int r = idx >> 7, k = idx & 127;
        int gr = r / V, w = r - gr * V, gk = k / V, v = k - gk * V;
        float val = (gr == gk && gr < G) ? A[v * V + w] : 0.f;
        int slab = k >> 6, col = k & 63;
        uint32_t off = slab * tc::slab_bytes(128) + tc::slab_chunk_off(r, col >> 3) + (col & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16_rn(val);
    }
}

struct FzArgs {
    const float* up; const float* ss;                 // layer input (pre-BN of the previous layer or start conv) + folded scale/shift
    const uint8_t* pack;                              // this layer's packed weights
    const uint8_t* bd;                                // block-diagonal diffusion operator (fz_pack_bd_kernel)
    const float* bf; const float* bg; const float* bm;
    float* TF; float* SG; float* Y; float* X1; float* X2; float* U; float* ycat; double* stats;
    LayerGeom g; int layer, L, Tl, groups, gpt;       // groups = B*To, gpt = groups per tile
    // BatchNorm finalize by the last CTA to finish (gwnet.py:120,237)
    unsigned int* ticket; double count; const float* gamma; const float* beta; float* rmean; float* rvar; long long* nbt;
    float* mr; float* ss_next; int training;
    float bn_momentum, bn_eps;                        // of the module's BatchNorm2d (gwnet.py:120)
#ifdef HOPK_DEBUG
    int stop;                                         // timing experiments only (HOPK_FZ_STOP): leave the kernel after stage `stop`
#endif
};

constexpr uint32_t FZ_R1 = 3 * tc::slab_bytes(128);                  // gate weights, then the diffusion operator
constexpr uint32_t FZ_R2 = FZ_R1 + FZ_WG_BYTES;                      // mlp weights
constexpr size_t fz_smem_bytes() { return FZ_R2 + FZ_WM_BYTES + 1024; }   // 105 KB: two CTAs / SM

__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// one diffusion hop on the tensor core: D[128 x 64] = BD[128 x 128] . X[128 x 64], X = slab `src` read MN-major
__device__ __forceinline__ void fz_issue_hop(uint32_t tmem_d, uint32_t bd_addr, uint32_t src_addr, uint64_t* bar)
{
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t SL128 = tc::slab_bytes(128);
#pragma unroll
    for (int t = 0; t < 8; ++t)
        tc::mma_bf16(tmem_d, tc::desc_kmajor(bd_addr + (t >> 2) * SL128, t & 3), tc::desc_mnmajor(src_addr, SL128, t), idesc, t != 0);
    tc::mma_commit(bar);
}

#ifdef HOPK_DEBUG
#define FZ_STOP_AT(k)                                                                               \
    if (a.stop == (k)) {                                                                            \
        if (tid == 0) { tc::mbar_wait(&bars[0], 0); tc::mbar_wait(&bars[1], 0); if ((k) >= 3) tc::mbar_wait(&bars[2], 0); } \
        tc::fence_before_sync();                                                                    \
        __syncthreads();                                                                            \
        if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 256);                                       \
        return;                                                                                     \
    }
#define FZ_STOP_IS(k) (a.stop == (k))
#else
#define FZ_STOP_AT(k)
#define FZ_STOP_IS(k) false
#endif

__global__ void __launch_bounds__(256, 2) fz_layer_fwd_kernel(const FzArgs a)
{
#ifdef HOPK_DEBUG
    if (a.stop == -1) return;
#endif
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[4];                      // 0: gate weights, 1: mlp weights, 2: diffusion operator, 3: UMMA completion
    __shared__ uint32_t tmem_base_smem;
    __shared__ float red[256];
    __shared__ __align__(16) float cst[5 * FZ_C];     // filter bias | gate bias | mlp bias | BN scale | BN shift
    __shared__ unsigned int is_last;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int C = FZ_C;
    constexpr uint32_t SL128 = tc::slab_bytes(128), SL64 = tc::slab_bytes(64);
    uint8_t* A1 = smem;                       // phase 1: [x_t | x_{t+d}] 2 slabs        phase 2: [y | x1 | x2] 3 slabs
    uint8_t* A2 = smem;
    uint8_t* Wg = smem + FZ_R1;               // phase 1: gate weights; phase 2: diffusion operator BD
    uint8_t* Wm = smem + FZ_R2;               // mlp weights 3 x 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = a.g.V;
    const int g0 = blockIdx.x * a.gpt;
    const int ng = min(a.gpt, a.groups - g0);
    const int r0 = g0 * V, nrows = ng * V;               // rows of this tile in the layer's output rows layout

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&bars[0], FZ_WG_BYTES);
        tc::bulk_g2s(Wg, a.pack, FZ_WG_BYTES, &bars[0]);
        tc::mbar_expect_tx(&bars[1], FZ_WM_BYTES);
        tc::bulk_g2s(Wm, a.pack + FZ_WG_BYTES, FZ_WM_BYTES, &bars[1]);
    }
    red[tid] = 0.f;
    if (tid < FZ_C) {
        cst[tid] = __ldg(a.bf + tid); cst[FZ_C + tid] = __ldg(a.bg + tid); cst[2 * FZ_C + tid] = __ldg(a.bm + tid);
        cst[3 * FZ_C + tid] = __ldg(a.ss + tid); cst[4 * FZ_C + tid] = __ldg(a.ss + FZ_C + tid);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 256);
    // ---- 1. stage the gate operands
    {
        W8GateX ld{a.up, a.ss, a.g, 2 * C};
        float f[(128 * 16) / 256][8];
#pragma unroll
        for (int it = 0; it < (128 * 16) / 256; ++it) {      // load phase: all global loads in flight together
            int idx = tid + it * 256;
            int ch = idx & 15, row = idx >> 4;
            if (row < nrows) ld.ld8(r0 + row, ch * 8, f[it]);
            else {
#pragma unroll
                for (int q = 0; q < 8; ++q) f[it][q] = 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < (128 * 16) / 256; ++it) {
            int idx = tid + it * 256;
            int ch = idx & 15, row = idx >> 4;
            tc::slab_store8(A1 + (ch >> 3) * SL128, row, ch & 7, f[it]);
        }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_g = tmem_base_smem, tmem_h = tmem_base_smem + 128, tmem_x = tmem_base_smem + 192;
    uint32_t mma_phase = 0;
    FZ_STOP_AT(1)
    // ---- 2. gate UMMA
    if (tid == 0) {
        tc::mbar_wait(&bars[0], 0);                                   // gate weights landed
        constexpr uint32_t idesc = tc::idesc_bf16(128, 128, 0, 0);
        const uint32_t aa = tc::smem_u32(A1), wa = tc::smem_u32(Wg);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int t = 0; t < 4; ++t)
                tc::mma_bf16(tmem_g, tc::desc_kmajor(aa + c * SL128, t), tc::desc_kmajor(wa + c * SL128, t), idesc, (c | t) != 0);
        tc::mma_commit(&bars[3]);
    }
    tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
    tc::fence_after_sync();
    // the gate weights are dead: the diffusion operator takes their place while the gating epilogue runs
    if (tid == 0) {
        tc::mbar_expect_tx(&bars[2], FZ_BD_BYTES);
        tc::bulk_g2s(Wg, a.bd, FZ_BD_BYTES, &bars[2]);
    }
    FZ_STOP_AT(3)
    // ---- 3..6: warp roles.  Warps 0-3 ("feeders") turn TMEM results into the next bf16 operand slabs and execute the
    // proxy fences; warps 4-7 ("writers") read the same TMEM lanes and own every global store plus the statistics.
    // The split matters because fence.proxy.async is MEMBAR.ALL.CTA in SASS: a thread that fences with global stores
    // in flight waits for all of them, which serialised every stage on a store round trip.
    const int row = (warp & 3) * 32 + lane;
    const bool writer = warp >= 4;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const bool rvalid = row < nrows;
    const int m = r0 + row;                                   // global output row
    int tt = -1; size_t ycat_off = 0;
    if (rvalid) {
        int vv = m % V; int bt = m / V; int t = bt % a.g.To; int b = bt / a.g.To;
        tt = t - (a.g.To - a.Tl);
        if (tt >= 0) ycat_off = ((size_t)(b * a.Tl + tt) * V + vv) * ((size_t)a.L * C) + (size_t)a.layer * C;
    }
    // residual operand of the last stage: requested now, consumed at the very end (writers only)
    float4 xr[16];
    if (writer) {
        const long rr = rvalid ? a.g.in_row(m) + (long)a.g.d * V : 0;
#pragma unroll
        for (int q = 0; q < 16; ++q)
            xr[q] = rvalid ? __ldg(reinterpret_cast<const float4*>(a.up + rr * C) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- gating: 4 column blocks of [16 filter | 16 gate] pre-activations per row
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
        const int c0 = blk * 16;
        float v[32];
        tc::tmem_ld32(tmem_g + lane_off + blk * 32, v);
        float y[16];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 b1 = *reinterpret_cast<const float4*>(cst + c0 + j4), b2 = *reinterpret_cast<const float4*>(cst + FZ_C + c0 + j4);
            const float bfv[4] = {b1.x, b1.y, b1.z, b1.w}, bgv[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j4 + q;
                float tf = tanh_fast(v[j] + bfv[q]);                                    // MUFU.TANH (2^-11 rel. error << bf16)
                float sg = fmaf(0.5f, tanh_fast(0.5f * (v[16 + j] + bgv[q])), 0.5f);    // sigmoid(x) = (1 + tanh(x/2)) / 2
                y[j] = rvalid ? tf * sg : 0.f;
                v[j] = tf; v[16 + j] = sg;
            }
        }
        if (FZ_STOP_IS(31)) continue;                           // timing experiment: TMEM reads + math only
        if (!writer) {
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {                  // y as bf16 into slab 0 of the mlp A operand
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = y[q8 * 8 + j];
                tc::slab_store8(A2, row, (c0 >> 3) + q8, f);
            }
        } else if (rvalid && !FZ_STOP_IS(32)) {
            const size_t o = (size_t)m * C + c0;
#pragma unroll
            for (int j = 0; j < 16; j += 8) {                 // full 32-byte sectors per store
                tc::stg256(a.TF + o + j, v + j);
                tc::stg256(a.SG + o + j, v + 16 + j);
                if (a.Y) tc::stg256(a.Y + o + j, y + j);
                if (tt >= 0) tc::stg256(a.ycat + ycat_off + c0 + j, y + j);
            }
        }
    }
    FZ_STOP_AT(4) FZ_STOP_AT(31) FZ_STOP_AT(32)
    // ---- 4. diffusion on the tensor core: x1 = BD . y, x2 = BD . x1 (bf16 operands, fp32 accumulate, like the mlp)
#pragma unroll 1
    for (int hop = 0; hop < 2; ++hop) {
        if (!writer) tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            if (hop == 0) tc::mbar_wait(&bars[2], 0);                 // diffusion operator landed
            fz_issue_hop(tmem_x, tc::smem_u32(Wg), tc::smem_u32(A2 + hop * SL128), &bars[3]);
        }
        tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
        tc::fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            float v[32];
            tc::tmem_ld32(tmem_x + lane_off + c * 32, v);
            if (!writer) {
#pragma unroll
                for (int q8 = 0; q8 < 4; ++q8) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = v[q8 * 8 + j];
                    tc::slab_store8(A2 + (1 + hop) * SL128, row, c * 4 + q8, f);
                }
            } else if (rvalid && a.X1) {
                float* go = (hop == 0 ? a.X1 : a.X2) + (size_t)m * C + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) tc::stg256(go + j, v + j);
            }
        }
    }
    FZ_STOP_AT(5)
    // ---- 5. mlp UMMA: [y | x1 | x2] (K = 192) x Wm^T (N = 64)
    if (!writer) tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        tc::mbar_wait(&bars[1], 0);                                   // mlp weights landed (long ago)
        constexpr uint32_t idesc = tc::idesc_bf16(128, 64, 0, 0);
        const uint32_t aa = tc::smem_u32(A2), wa = tc::smem_u32(Wm);
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int t = 0; t < 4; ++t)
                tc::mma_bf16(tmem_h, tc::desc_kmajor(aa + s * SL128, t), tc::desc_kmajor(wa + s * SL64, t), idesc, (s | t) != 0);
        tc::mma_commit(&bars[3]);
    }
    tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
    tc::fence_after_sync();
    FZ_STOP_AT(6)
    // ---- 6. bias + residual + statistics (writers): thread = one row, two blocks of 32 output channels
    if (writer) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            float uv[32], w[32];
            tc::tmem_ld32(tmem_h + lane_off + c * 32, uv);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = c * 32 + 4 * q;
                const float4 sc = *reinterpret_cast<const float4*>(cst + 3 * FZ_C + n), sh = *reinterpret_cast<const float4*>(cst + 4 * FZ_C + n);
                const float4 bm = *reinterpret_cast<const float4*>(cst + 2 * FZ_C + n);
                const float4 x = c == 0 ? xr[q] : xr[8 + q];
                float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
                if (rvalid) {
                    u0 = uv[4 * q] + bm.x + fmaf(x.x, sc.x, sh.x); u1 = uv[4 * q + 1] + bm.y + fmaf(x.y, sc.y, sh.y);
                    u2 = uv[4 * q + 2] + bm.z + fmaf(x.z, sc.z, sh.z); u3 = uv[4 * q + 3] + bm.w + fmaf(x.w, sc.w, sh.w);
                }
                uv[4 * q] = u0; uv[4 * q + 1] = u1; uv[4 * q + 2] = u2; uv[4 * q + 3] = u3;
                w[4 * q] = u0 * u0; w[4 * q + 1] = u1 * u1; w[4 * q + 2] = u2 * u2; w[4 * q + 3] = u3 * u3;
            }
            if (rvalid) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) tc::stg256(a.U + (size_t)m * C + c * 32 + j, uv + j);
            }
            const float s2 = warp_transpose_sum(w);
            const float s1 = warp_transpose_sum(uv);
            // per-warp slots in the (now dead) operand region, combined below in a fixed order: bit-reproducible statistics
            float* part = reinterpret_cast<float*>(A2) + (warp & 3) * 128;
            part[c * 32 + lane] = s1;
            part[64 + c * 32 + lane] = s2;
        }
    }
    FZ_STOP_AT(7)
    tc::fence_before_sync();
    __syncthreads();
    if (tid < C) {                                                    // feeder threads: no global stores of their own in flight
        const float* part = reinterpret_cast<const float*>(A2);
        atomicAdd(a.stats + tid, (double)((part[tid] + part[128 + tid]) + (part[256 + tid] + part[384 + tid])));
        atomicAdd(a.stats + C + tid, (double)((part[64 + tid] + part[192 + tid]) + (part[320 + tid] + part[448 + tid])));
        __threadfence();                                              // statistics visible before the ticket
    }
    if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 256);
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    // ---- 7. the last CTA to arrive turns the statistics into mean / rstd / folded scale+shift / running stats
    if (is_last && tid < C) {
        __threadfence();
        const int c = tid;
        float mean, rstd;
        if (a.training) {
            const double s1 = __ldcg(a.stats + c), s2 = __ldcg(a.stats + C + c);
            double mu = s1 / a.count;
            double var = s2 / a.count - mu * mu;
            if (var < 0) var = 0;
            mean = (float)mu;
            rstd = (float)(1.0 / sqrt(var + (double)a.bn_eps));
            double unb = a.count > 1 ? var * a.count / (a.count - 1) : var;
            a.rmean[c] = (1.f - a.bn_momentum) * a.rmean[c] + a.bn_momentum * mean;
            a.rvar[c] = (1.f - a.bn_momentum) * a.rvar[c] + a.bn_momentum * (float)unb;
            if (c == 0 && a.nbt) *a.nbt += 1;
        } else {
            mean = a.rmean[c];
            rstd = 1.f / sqrtf(a.rvar[c] + a.bn_eps);
        }
        a.mr[c] = mean; a.mr[C + c] = rstd;
        float sc = a.gamma[c] * rstd;
        a.ss_next[c] = sc; a.ss_next[C + c] = a.beta[c] - mean * sc;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole stack of layers as ONE persistent kernel: the grid (every CTA resident: cooperative launch) walks the layers,
// a CTA takes tiles blockIdx.x, blockIdx.x + gridDim.x, ... of each, and a grid barrier on a per-layer counter separates
// a layer's BatchNorm statistics from their use: after it every CTA folds the fp64 sums into the scale / shift of the next
// layer's loads itself (CTA 0 also writes the running statistics and the mean / rstd / scale / shift backward reads).
// Activations written by other CTAs earlier in the same launch are read with ld.global.cg (L2), never the read-only path.
// Saves, per layer, a kernel launch + drain, the TMEM allocation and the last-CTA finalize hop of fz_layer_fwd_kernel.
constexpr int FZ_NET_LAYERS = 8;                                    // the reference stack (4 blocks x 2 layers); deeper stacks run per layer
struct FzNetArgs {
    FzArgs layer[FZ_NET_LAYERS];
    int L;
};

__device__ __forceinline__ void fz_grid_barrier(unsigned int* counter, unsigned int expected)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                              // this CTA's stores / atomics before its arrival
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen < expected) __nanosleep(64);
        } while (seen < expected);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 2) fz_net_fwd_kernel(const __grid_constant__ FzNetArgs na)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[4];                      // 0: gate weights, 1: mlp weights, 2: diffusion operator, 3: UMMA completion
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float cst[5 * FZ_C];     // filter bias | gate bias | mlp bias | BN scale | BN shift (of the layer INPUT)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int C = FZ_C;
    constexpr uint32_t SL128 = tc::slab_bytes(128), SL64 = tc::slab_bytes(64);
    uint8_t* A1 = smem;                       // phase 1: [x_t | x_{t+d}] 2 slabs        phase 2: [y | x1 | x2] 3 slabs
    uint8_t* A2 = smem;
    uint8_t* Wg = smem + FZ_R1;               // phase 1: gate weights; phase 2: diffusion operator BD
    uint8_t* Wm = smem + FZ_R2;               // mlp weights 3 x 8 KB
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_g = tmem_base_smem, tmem_h = tmem_base_smem + 128, tmem_x = tmem_base_smem + 192;
    uint32_t mma_phase = 0, it = 0;           // UMMA barrier parity; tiles this CTA has processed (parity of the three load barriers)
    const int row = (warp & 3) * 32 + lane;
    const bool writer = warp >= 4;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;

#pragma unroll 1
    for (int l = 0; l < na.L; ++l) {
        const FzArgs& a = na.layer[l];
        const int V = a.g.V;
        if (tid < FZ_C) {
            cst[tid] = __ldg(a.bf + tid); cst[FZ_C + tid] = __ldg(a.bg + tid); cst[2 * FZ_C + tid] = __ldg(a.bm + tid);
            if (l == 0) { cst[3 * FZ_C + tid] = __ldg(a.ss + tid); cst[4 * FZ_C + tid] = __ldg(a.ss + FZ_C + tid); }
        }
        __syncthreads();
        const int ntiles = (a.groups + a.gpt - 1) / a.gpt;
#pragma unroll 1
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int g0 = tile * a.gpt;
            const int ng = min(a.gpt, a.groups - g0);
            const int r0 = g0 * V, nrows = ng * V;               // rows of this tile in the layer's output rows layout
            const uint32_t ld_phase = it & 1;
            if (tid == 0) {
                tc::mbar_expect_tx(&bars[0], FZ_WG_BYTES);
                tc::bulk_g2s(Wg, a.pack, FZ_WG_BYTES, &bars[0]);
                tc::mbar_expect_tx(&bars[1], FZ_WM_BYTES);
                tc::bulk_g2s(Wm, a.pack + FZ_WG_BYTES, FZ_WM_BYTES, &bars[1]);
            }
            // ---- 1. stage the gate operands: BatchNorm of the previous layer folded into the load (scale / shift in cst)
            {
                float f[(128 * 16) / 256][8];
#pragma unroll
                for (int i2 = 0; i2 < (128 * 16) / 256; ++i2) {      // load phase: all global loads in flight together
                    const int idx = tid + i2 * 256;
                    const int ch = idx & 15, rr = idx >> 4;
                    if (rr < nrows) {
                        const int c0 = ch * 8, tap = c0 >= C, c = c0 - tap * C;
                        const float* p = a.up + (size_t)(a.g.in_row(r0 + rr) + (long)tap * a.g.d * V) * C + c;
                        const float4 x0 = __ldcg(reinterpret_cast<const float4*>(p)), x1 = __ldcg(reinterpret_cast<const float4*>(p + 4));
                        const float* sc = cst + 3 * FZ_C + c; const float* sh = cst + 4 * FZ_C + c;
                        f[i2][0] = fmaf(x0.x, sc[0], sh[0]); f[i2][1] = fmaf(x0.y, sc[1], sh[1]); f[i2][2] = fmaf(x0.z, sc[2], sh[2]); f[i2][3] = fmaf(x0.w, sc[3], sh[3]);
                        f[i2][4] = fmaf(x1.x, sc[4], sh[4]); f[i2][5] = fmaf(x1.y, sc[5], sh[5]); f[i2][6] = fmaf(x1.z, sc[6], sh[6]); f[i2][7] = fmaf(x1.w, sc[7], sh[7]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) f[i2][q] = 0.f;
                    }
                }
#pragma unroll
                for (int i2 = 0; i2 < (128 * 16) / 256; ++i2) {
                    const int idx = tid + i2 * 256;
                    const int ch = idx & 15, rr = idx >> 4;
                    tc::slab_store8(A1 + (ch >> 3) * SL128, rr, ch & 7, f[i2]);
                }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
            // ---- 2. gate UMMA
            if (tid == 0) {
                tc::mbar_wait(&bars[0], ld_phase);                            // gate weights landed
                constexpr uint32_t idesc = tc::idesc_bf16(128, 128, 0, 0);
                const uint32_t aa = tc::smem_u32(A1), wa = tc::smem_u32(Wg);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        tc::mma_bf16(tmem_g, tc::desc_kmajor(aa + c * SL128, t), tc::desc_kmajor(wa + c * SL128, t), idesc, (c | t) != 0);
                tc::mma_commit(&bars[3]);
            }
            tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
            tc::fence_after_sync();
            // the gate weights are dead: the diffusion operator takes their place while the gating epilogue runs
            if (tid == 0) {
                tc::mbar_expect_tx(&bars[2], FZ_BD_BYTES);
                tc::bulk_g2s(Wg, a.bd, FZ_BD_BYTES, &bars[2]);
            }
            // ---- 3..6: warp roles as in fz_layer_fwd_kernel (feeders: TMEM -> bf16 slabs + proxy fences; writers: global stores)
            const bool rvalid = row < nrows;
            const int m = r0 + row;                                   // global output row
            int tt = -1; size_t ycat_off = 0;
            if (rvalid) {
                int vv = m % V; int bt = m / V; int t = bt % a.g.To; int b = bt / a.g.To;
                tt = t - (a.g.To - a.Tl);
                if (tt >= 0) ycat_off = ((size_t)(b * a.Tl + tt) * V + vv) * ((size_t)a.L * C) + (size_t)a.layer * C;
            }
            float4 xr[16];                                            // residual operand of the last stage (writers)
            if (writer) {
                const long rr = rvalid ? a.g.in_row(m) + (long)a.g.d * V : 0;
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    xr[q] = rvalid ? __ldcg(reinterpret_cast<const float4*>(a.up + rr * C) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                const int c0 = blk * 16;
                float v[32];
                tc::tmem_ld32(tmem_g + lane_off + blk * 32, v);
                float y[16];
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const float4 b1 = *reinterpret_cast<const float4*>(cst + c0 + j4), b2 = *reinterpret_cast<const float4*>(cst + FZ_C + c0 + j4);
                    const float bfv[4] = {b1.x, b1.y, b1.z, b1.w}, bgv[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = j4 + q;
                        float tf = tanh_fast(v[j] + bfv[q]);
                        float sg = fmaf(0.5f, tanh_fast(0.5f * (v[16 + j] + bgv[q])), 0.5f);
                        y[j] = rvalid ? tf * sg : 0.f;
                        v[j] = tf; v[16 + j] = sg;
                    }
                }
                if (!writer) {
#pragma unroll
                    for (int q8 = 0; q8 < 2; ++q8) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = y[q8 * 8 + j];
                        tc::slab_store8(A2, row, (c0 >> 3) + q8, f);
                    }
                } else if (rvalid) {
                    const size_t o = (size_t)m * C + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        tc::stg256(a.TF + o + j, v + j);
                        tc::stg256(a.SG + o + j, v + 16 + j);
                        if (tt >= 0) tc::stg256(a.ycat + ycat_off + c0 + j, y + j);
                    }
                }
            }
            // ---- 4. diffusion on the tensor core: x1 = BD . y, x2 = BD . x1
#pragma unroll 1
            for (int hop = 0; hop < 2; ++hop) {
                if (!writer) tc::fence_async_smem();
                tc::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    tc::fence_after_sync();
                    if (hop == 0) tc::mbar_wait(&bars[2], ld_phase);          // diffusion operator landed
                    fz_issue_hop(tmem_x, tc::smem_u32(Wg), tc::smem_u32(A2 + hop * SL128), &bars[3]);
                }
                tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
                tc::fence_after_sync();
                if (!writer) {
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        float v[32];
                        tc::tmem_ld32(tmem_x + lane_off + c * 32, v);
#pragma unroll
                        for (int q8 = 0; q8 < 4; ++q8) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = v[q8 * 8 + j];
                            tc::slab_store8(A2 + (1 + hop) * SL128, row, c * 4 + q8, f);
                        }
                    }
                }
            }
            // ---- 5. mlp UMMA: [y | x1 | x2] (K = 192) x Wm^T (N = 64)
            if (!writer) tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                tc::mbar_wait(&bars[1], ld_phase);                            // mlp weights landed (long ago)
                constexpr uint32_t idesc = tc::idesc_bf16(128, 64, 0, 0);
                const uint32_t aa = tc::smem_u32(A2), wa = tc::smem_u32(Wm);
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2)
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        tc::mma_bf16(tmem_h, tc::desc_kmajor(aa + s2 * SL128, t), tc::desc_kmajor(wa + s2 * SL64, t), idesc, (s2 | t) != 0);
                tc::mma_commit(&bars[3]);
            }
            tc::mbar_wait(&bars[3], mma_phase); mma_phase ^= 1;
            tc::fence_after_sync();
            // ---- 6. bias + residual + statistics (writers): thread = one row, two blocks of 32 output channels
            if (writer) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float uv[32], w[32];
                    tc::tmem_ld32(tmem_h + lane_off + c * 32, uv);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int n = c * 32 + 4 * q;
                        const float4 sc = *reinterpret_cast<const float4*>(cst + 3 * FZ_C + n), sh = *reinterpret_cast<const float4*>(cst + 4 * FZ_C + n);
                        const float4 bm = *reinterpret_cast<const float4*>(cst + 2 * FZ_C + n);
                        const float4 x = c == 0 ? xr[q] : xr[8 + q];
                        float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
                        if (rvalid) {
                            u0 = uv[4 * q] + bm.x + fmaf(x.x, sc.x, sh.x); u1 = uv[4 * q + 1] + bm.y + fmaf(x.y, sc.y, sh.y);
                            u2 = uv[4 * q + 2] + bm.z + fmaf(x.z, sc.z, sh.z); u3 = uv[4 * q + 3] + bm.w + fmaf(x.w, sc.w, sh.w);
                        }
                        uv[4 * q] = u0; uv[4 * q + 1] = u1; uv[4 * q + 2] = u2; uv[4 * q + 3] = u3;
                        w[4 * q] = u0 * u0; w[4 * q + 1] = u1 * u1; w[4 * q + 2] = u2 * u2; w[4 * q + 3] = u3 * u3;
                    }
                    if (rvalid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) tc::stg256(a.U + (size_t)m * C + c * 32 + j, uv + j);
                    }
                    const float s2 = warp_transpose_sum(w);
                    const float s1 = warp_transpose_sum(uv);
                    float* part = reinterpret_cast<float*>(A2) + (warp & 3) * 128;      // fixed combination order: reproducible
                    part[c * 32 + lane] = s1;
                    part[64 + c * 32 + lane] = s2;
                }
            }
            tc::fence_before_sync();
            __syncthreads();
            if (tid < C) {                                                    // feeder threads: no global stores of their own in flight
                const float* part = reinterpret_cast<const float*>(A2);
                atomicAdd(a.stats + tid, (double)((part[tid] + part[128 + tid]) + (part[256 + tid] + part[384 + tid])));
                atomicAdd(a.stats + C + tid, (double)((part[64 + tid] + part[192 + tid]) + (part[320 + tid] + part[448 + tid])));
            }
            __syncthreads();                                                  // the operand region is reused by the next tile
        }
        // ---- 7. every tile of the layer is out: BatchNorm statistics -> scale / shift of the next layer's loads
        fz_grid_barrier(a.ticket, gridDim.x);
        if (tid < C) {
            const int c = tid;
            float mean, rstd;
            if (a.training) {
                const double s1 = __ldcg(a.stats + c), s2 = __ldcg(a.stats + C + c);
                double mu = s1 / a.count;
                double var = s2 / a.count - mu * mu;
                if (var < 0) var = 0;
                mean = (float)mu;
                rstd = (float)(1.0 / sqrt(var + (double)a.bn_eps));
                if (blockIdx.x == 0) {
                    double unb = a.count > 1 ? var * a.count / (a.count - 1) : var;
                    a.rmean[c] = (1.f - a.bn_momentum) * a.rmean[c] + a.bn_momentum * mean;
                    a.rvar[c] = (1.f - a.bn_momentum) * a.rvar[c] + a.bn_momentum * (float)unb;
                    if (c == 0 && a.nbt) *a.nbt += 1;
                }
            } else {
                mean = __ldg(a.rmean + c);
                rstd = 1.f / sqrtf(__ldg(a.rvar + c) + a.bn_eps);
            }
            const float sc = __ldg(a.gamma + c) * rstd, sh = __ldg(a.beta + c) - mean * sc;
            cst[3 * FZ_C + c] = sc; cst[4 * FZ_C + c] = sh;
            if (blockIdx.x == 0) { a.mr[c] = mean; a.mr[C + c] = rstd; a.ss_next[c] = sc; a.ss_next[C + c] = sh; }
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_smem, 256);
}

}  // namespace hopk
