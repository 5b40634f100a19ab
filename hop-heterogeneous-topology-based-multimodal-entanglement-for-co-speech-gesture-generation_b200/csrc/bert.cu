// bert.cu -- the small kernels of the frozen BERT-6 encoder between its dense GEMMs (reference model/HOP.py:202-204:
// `self.llm_model(inputs_embeds=llama_enc_out).last_hidden_state`; the encoder is frozen, HOP.py:90-91, so backward only
// needs dX).  The GEMMs themselves (QKV / output / FFN projections and their dX products) run on gemm_tma.cu; here:
//   layer norm forward / backward (dX)        BertEmbeddings.LayerNorm, BertSelfOutput.LayerNorm, BertOutput.LayerNorm
//   self-attention forward / backward         12 heads x 64, sequence 34: two (sample, head) pairs per CTA on tcgen05 UMMAs
//   GELU                                      BertIntermediate (exact erf form)
// dtype-1 arithmetic: bf16 operands for the tensor-core GEMMs, fp32 statistics / softmax / residual stream.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_core.cuh"
#include "../../include/hopk.h"

namespace hopk {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t bf2(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------- layer norm: one warp per row, C = 128 * NV (NV <= 8)
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ add, int period,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                     float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, float* __restrict__ stat, int M)
{
    constexpr int C = 128 * NV;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
    const float4* ar = add ? reinterpret_cast<const float4*>(add + (size_t)(row % period) * C) : nullptr;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = __ldg(xr + lane + 32 * i);
        if (ar) { const float4 a = __ldg(ar + lane + 32 * i); v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w; }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    if (lane == 0 && stat) { stat[2 * row] = mean; stat[2 * row + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i), b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * C)[lane + 32 * i] = o;
        if (y16) reinterpret_cast<uint2*>(y16 + (size_t)row * C)[lane + 32 * i] = make_uint2(bf2(o.x, o.y), bf2(o.z, o.w));
    }
}

// dX of layer norm (gamma / beta frozen): dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ add,
                                                     int period, const float* __restrict__ gamma, const float* __restrict__ stat,
                                                     float* __restrict__ dx32, __nv_bfloat16* __restrict__ dx16, int M)
{
    constexpr int C = 128 * NV;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float mean = stat[2 * row], rstd = stat[2 * row + 1];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
    const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)row * C);
    const float4* ar = add ? reinterpret_cast<const float4*>(add + (size_t)(row % period) * C) : nullptr;
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 v = __ldg(xr + lane + 32 * i);
        if (ar) { const float4 a = __ldg(ar + lane + 32 * i); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        const float4 d = __ldg(dr + lane + 32 * i), gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        xh[i] = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
        g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 o;
        o.x = rstd * (g[i].x - m1 - xh[i].x * m2); o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
        o.z = rstd * (g[i].z - m1 - xh[i].z * m2); o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
        if (dx32) reinterpret_cast<float4*>(dx32 + (size_t)row * C)[lane + 32 * i] = o;
        if (dx16) reinterpret_cast<uint2*>(dx16 + (size_t)row * C)[lane + 32 * i] = make_uint2(bf2(o.x, o.y), bf2(o.z, o.w));
    }
}

// ---------------------------------------------------------------- GELU (exact), bf16 -> bf16, 8 elements per thread
__global__ void gelu_bf16_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ out, size_t n8)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(pre) + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = __low2float(h[j]), b = __high2float(h[j]);
            o[j] = bf2(gelu_fast(a), gelu_fast(b));
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------- self-attention on the tensor core
// qkv: bf16 [B*S][3*H*D] rows = (b, t): [q heads | k heads | v heads]; ctx: bf16 [B*S][H*D]   (BertSelfAttention, D = 64)
// A CTA owns two consecutive (sample, head) pairs; pair g occupies rows 64 g .. 64 g + S - 1 of every 128-row operand slab
// (S <= 64), rows beyond S are zero.
//   forward:  S = Q K^T as one 128 x 128 x 64 UMMA (only the two 64 x 64 diagonal blocks are used) -> TMEM; thread = row:
//             softmax over its S columns in registers; P (bf16) becomes a block-diagonal A operand; O = P V (128 x 64 x 128)
//             -> TMEM -> bf16 context rows.
//   backward: recomputes S and the softmax (only qkv is saved), dP = dO V^T beside it; dS = P (dP - sum P dP) / 8;
//             dQ = dS K, dK = dS^T Q, dV = P^T dO read P / dS through K-major and MN-major descriptors of one image.
// A block-diagonal operand is stored compactly as [block 0: 64 rows][64 zero rows][block 1: 64 rows] (8 KB each): its
// k < 64 half is the 128-row view at byte 0 and its k >= 64 half the 128-row view at byte 8192.
constexpr int AT_MAXS = 64, AT_D = 64;
constexpr uint32_t AT_SLAB = tc::slab_bytes(128);             // 16 KB
constexpr uint32_t AT_HALF = tc::slab_bytes(64);              // 8 KB
constexpr uint32_t AT_BD = 3 * AT_HALF;                       // block-diagonal operand
constexpr size_t AT_FWD_SMEM = 3 * AT_SLAB + 1024, AT_BWD_SMEM = 3 * AT_SLAB + 2 * AT_BD + 1024;   // P starts on the V slab

constexpr int AT_PITCH = 144;                                 // staging row pitch: 128 data bytes + 16 (conflict-free 16-byte stores)

// Q, K, V (and dO) of one pair are 64-row boxes of the row-major activations loaded by the TMA engine straight into the 128B
// swizzled slabs.  Rows S..63 of a box belong to the next sample (or are zero past the end): finite values that only meet
// zero probabilities / masked score columns, so nothing has to be cleared.
__device__ __forceinline__ void at_zero_half(uint8_t* p)     // 8 KB
{
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(p)[threadIdx.x + i * 128] = make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ float ex2_fast(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int SP>
__device__ __forceinline__ void at_ld_row(uint32_t taddr, float (&v)[SP])
{
    float a[32];
    tc::tmem_ld32(taddr, a);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = a[i];
    if constexpr (SP == 64) {
        tc::tmem_ld32(taddr + 32, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[32 + i] = a[i];
    } else {
        float c[8];
        tc::tmem_ld8(taddr + 32, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[32 + i] = c[i];
    }
}
// a thread's row of a block-diagonal operand: 64 bf16 (columns >= SP are zero) into block g
template <int SP>
__device__ __forceinline__ void at_store_row(uint8_t* bd, int g, int t, const float (&v)[SP])
{
    uint8_t* blk = bd + (g ? 2 * AT_HALF : 0);
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (ch * 8 < SP) u = make_uint4(bf2(v[ch * 8], v[ch * 8 + 1]), bf2(v[ch * 8 + 2], v[ch * 8 + 3]), bf2(v[ch * 8 + 4], v[ch * 8 + 5]), bf2(v[ch * 8 + 6], v[ch * 8 + 7]));
        *reinterpret_cast<uint4*>(blk + tc::slab_chunk_off(t, ch)) = u;
    }
}
// softmax(s / 8) over the first S entries, in place (entries >= S become 0; an invalid row becomes all zero)
template <int SP>
__device__ __forceinline__ void at_softmax(float (&s)[SP], int S, bool valid)
{
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < SP; ++c) if (c < S) m = fmaxf(m, s[c]);
    constexpr float kf = 0.125f * 1.4426950408889634f;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < SP; ++c) { const float e = c < S ? ex2_fast((s[c] - m) * kf) : 0.f; s[c] = e; sum += e; }
    const float inv = valid ? 1.f / sum : 0.f;
#pragma unroll
    for (int c = 0; c < SP; ++c) s[c] *= inv;
}
// TMEM row (64 fp32) -> 64 bf16 in the thread's staging row -> one 128-byte bulk store to global memory
__device__ __forceinline__ void at_store_out(uint32_t taddr, uint8_t* stage_row, __nv_bfloat16* dst, bool valid)
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float o[32];
        tc::tmem_ld32(taddr + half * 32, o);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            reinterpret_cast<uint4*>(stage_row + half * 64)[q] = make_uint4(bf2(o[8 * q], o[8 * q + 1]), bf2(o[8 * q + 2], o[8 * q + 3]),
                                                                            bf2(o[8 * q + 4], o[8 * q + 5]), bf2(o[8 * q + 6], o[8 * q + 7]));
    }
    tc::fence_async_smem();                                    // this thread's row -> visible to the bulk copy it issues
    if (valid) tc::bulk_s2g(dst, stage_row, 128);
}

template <int SP>
__global__ void __launch_bounds__(128) bert_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx,
                                                            float* __restrict__ P, int S, int H, int npairs)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];                               // 0: operands landed, 1: UMMA completion
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sK = smem + AT_SLAB, *sV = smem + 2 * AT_SLAB;
    uint8_t* sP = smem;                                        // over Q and the first half of K once S = Q K^T is complete
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pair0 = blockIdx.x * 2;
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&bars[0], 6 * AT_HALF);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int pr = pair0 + g, b = pr / H, h = pr - b * H;
            tc::tma_load_2d(sQ + g * AT_HALF, &tmQKV, h * AT_D, b * S, &bars[0]);
            tc::tma_load_2d(sK + g * AT_HALF, &tmQKV, (H + h) * AT_D, b * S, &bars[0]);
            tc::tma_load_2d(sV + g * AT_HALF, &tmQKV, (2 * H + h) * AT_D, b * S, &bars[0]);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_base_smem;
    if (tid == 0) {
        tc::mbar_wait(&bars[0], 0);
        constexpr uint32_t idesc = tc::idesc_bf16(128, 128, 0, 0);
        const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ), 0), dk = tc::desc_kmajor(tc::smem_u32(sK), 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::mma_bf16(tm, tc::desc_adv(dq, 32 * k), tc::desc_adv(dk, 32 * k), idesc, k != 0);
        tc::mma_commit(&bars[1]);
    }
    const int g = tid >> 6, t = tid & 63, pr = pair0 + g;
    const bool valid = t < S && pr < npairs;
    const int b = pr / H, h = pr - b * H;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    tc::mbar_wait(&bars[1], 0);
    tc::fence_after_sync();
    float s[SP];
    at_ld_row<SP>(tm + lane_off + g * 64, s);
    at_softmax<SP>(s, S, valid);
    if (P && valid) {
        float* prow = P + ((size_t)pr * S + t) * S;
#pragma unroll
        for (int c = 0; c < SP; ++c) if (c < S) prow[c] = s[c];
    }
    at_zero_half(sP + AT_HALF);
    at_store_row<SP>(sP, g, t, s);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        constexpr uint32_t idesc = tc::idesc_bf16(128, 64, 0, 1);
        const uint64_t dp = tc::desc_kmajor(tc::smem_u32(sP), 0), dv = tc::desc_mnmajor(tc::smem_u32(sV), 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) tc::mma_bf16(tm, tc::desc_adv(dp, (k >> 2) * AT_HALF + (k & 3) * 32), tc::desc_adv(dv, 2048 * k), idesc, k != 0);
        tc::mma_commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 1);
    tc::fence_after_sync();
    at_store_out(tm + lane_off, smem + tid * AT_PITCH, ctx + ((size_t)b * S + t) * ((size_t)H * AT_D) + h * AT_D, valid);
    tc::bulk_commit();
    tc::bulk_wait_read();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 128);
}

template <int SP>
__global__ void __launch_bounds__(128) bert_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                                                            __nv_bfloat16* __restrict__ dqkv, int S, int H, int npairs)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sK = smem + AT_SLAB, *sD = smem + 2 * AT_SLAB, *sV = smem + 3 * AT_SLAB;
    uint8_t* sP = sV;                                          // P: over V (dead after dP = dO V^T) and 8 KB more
    uint8_t* sS = sP + AT_BD;                                  // dS
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pair0 = blockIdx.x * 2;
    const size_t ld = (size_t)3 * H * AT_D;
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&bars[0], 8 * AT_HALF);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int pr = pair0 + g, b = pr / H, h = pr - b * H;
            tc::tma_load_2d(sQ + g * AT_HALF, &tmQKV, h * AT_D, b * S, &bars[0]);
            tc::tma_load_2d(sK + g * AT_HALF, &tmQKV, (H + h) * AT_D, b * S, &bars[0]);
            tc::tma_load_2d(sV + g * AT_HALF, &tmQKV, (2 * H + h) * AT_D, b * S, &bars[0]);
            tc::tma_load_2d(sD + g * AT_HALF, &tmDO, h * AT_D, b * S, &bars[0]);
        }
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_smem, 256);
    at_zero_half(sS + AT_HALF);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_base_smem;
    if (tid == 0) {
        tc::mbar_wait(&bars[0], 0);
        constexpr uint32_t idesc = tc::idesc_bf16(128, 128, 0, 0);
        const uint64_t dq = tc::desc_kmajor(tc::smem_u32(sQ), 0), dk = tc::desc_kmajor(tc::smem_u32(sK), 0);
        const uint64_t dd = tc::desc_kmajor(tc::smem_u32(sD), 0), dv = tc::desc_kmajor(tc::smem_u32(sV), 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::mma_bf16(tm, tc::desc_adv(dq, 32 * k), tc::desc_adv(dk, 32 * k), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::mma_bf16(tm + 128, tc::desc_adv(dd, 32 * k), tc::desc_adv(dv, 32 * k), idesc, k != 0);
        tc::mma_commit(&bars[1]);
    }
    const int g = tid >> 6, t = tid & 63, pr = pair0 + g;
    const bool valid = t < S && pr < npairs;
    const int b = pr / H, h = pr - b * H;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    tc::mbar_wait(&bars[1], 0);
    tc::fence_after_sync();
    float p[SP], dp[SP];
    at_ld_row<SP>(tm + lane_off + g * 64, p);
    at_softmax<SP>(p, S, valid);
    at_ld_row<SP>(tm + lane_off + 128 + g * 64, dp);
    float delta = 0.f;
#pragma unroll
    for (int c = 0; c < SP; ++c) delta = fmaf(p[c], dp[c], delta);
#pragma unroll
    for (int c = 0; c < SP; ++c) dp[c] = p[c] * (dp[c] - delta) * 0.125f;
    at_zero_half(sP + AT_HALF);                                // rows 64..127 of the V slab
    at_store_row<SP>(sP, g, t, p);
    at_store_row<SP>(sS, g, t, dp);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        constexpr uint32_t i_kn = tc::idesc_bf16(128, 64, 0, 1), i_nn = tc::idesc_bf16(128, 64, 1, 1);
        const uint64_t ds_k = tc::desc_kmajor(tc::smem_u32(sS), 0), ds_n = tc::desc_mnmajor(tc::smem_u32(sS), AT_HALF, 0);
        const uint64_t dp_n = tc::desc_mnmajor(tc::smem_u32(sP), AT_HALF, 0);
        const uint64_t bk = tc::desc_mnmajor(tc::smem_u32(sK), 0, 0), bq = tc::desc_mnmajor(tc::smem_u32(sQ), 0, 0), bd = tc::desc_mnmajor(tc::smem_u32(sD), 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)                            // dQ = dS K
            tc::mma_bf16(tm, tc::desc_adv(ds_k, (k >> 2) * AT_HALF + (k & 3) * 32), tc::desc_adv(bk, 2048 * k), i_kn, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)                            // dK = dS^T Q
            tc::mma_bf16(tm + 64, tc::desc_adv(ds_n, 2048 * k), tc::desc_adv(bq, 2048 * k), i_nn, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)                            // dV = P^T dO
            tc::mma_bf16(tm + 128, tc::desc_adv(dp_n, 2048 * k), tc::desc_adv(bd, 2048 * k), i_nn, k != 0);
        tc::mma_commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 1);
    tc::fence_after_sync();
    __nv_bfloat16* o = dqkv + ((size_t)b * S + t) * ld + h * AT_D;
    uint8_t* stage = smem + tid * AT_PITCH;
    at_store_out(tm + lane_off, stage, o, valid);
    at_store_out(tm + lane_off + 64, stage + 128 * AT_PITCH, o + H * AT_D, valid);
    at_store_out(tm + lane_off + 128, stage + 256 * AT_PITCH, o + 2 * H * AT_D, valid);
    tc::bulk_commit();
    tc::bulk_wait_read();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}

static int grid_for(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : b); }

}  // namespace hopk
using namespace hopk;

extern "C" int hopk_ln_fwd(const float* x, const float* add, int period, const float* gamma, const float* beta, float eps, float* y32,
                           void* y16, float* stat, int M, int C, void* stream)
{
    HOPK_REQUIRE(M > 0 && C % 128 == 0 && C >= 128 && C <= 1024, "layer norm: C must be a multiple of 128, <= 1024");
    HOPK_REQUIRE(!add || period > 0, "layer norm: period of the row-periodic addend");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = cdiv(M, 8);
    __nv_bfloat16* y = (__nv_bfloat16*)y16;
    switch (C / 128) {
        case 1: ln_fwd_kernel<1><<<blocks, 256, 0, st>>>(x, add, period, gamma, beta, eps, y32, y, stat, M); break;
        case 2: ln_fwd_kernel<2><<<blocks, 256, 0, st>>>(x, add, period, gamma, beta, eps, y32, y, stat, M); break;
        case 4: ln_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, add, period, gamma, beta, eps, y32, y, stat, M); break;
        case 6: ln_fwd_kernel<6><<<blocks, 256, 0, st>>>(x, add, period, gamma, beta, eps, y32, y, stat, M); break;
        case 8: ln_fwd_kernel<8><<<blocks, 256, 0, st>>>(x, add, period, gamma, beta, eps, y32, y, stat, M); break;
        default: return fail(2, "bad argument:", "layer norm: C / 128 must be 1, 2, 4, 6 or 8");
    }
    HOPK_LAUNCH_CHECK("ln_fwd");
    return 0;
}

extern "C" int hopk_ln_bwd(const float* dy, const float* x, const float* add, int period, const float* gamma, const float* stat,
                           float* dx32, void* dx16, int M, int C, void* stream)
{
    HOPK_REQUIRE(M > 0 && C % 128 == 0 && C >= 128 && C <= 1024, "layer norm: C must be a multiple of 128, <= 1024");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = cdiv(M, 8);
    __nv_bfloat16* d = (__nv_bfloat16*)dx16;
    switch (C / 128) {
        case 1: ln_bwd_kernel<1><<<blocks, 256, 0, st>>>(dy, x, add, period, gamma, stat, dx32, d, M); break;
        case 2: ln_bwd_kernel<2><<<blocks, 256, 0, st>>>(dy, x, add, period, gamma, stat, dx32, d, M); break;
        case 4: ln_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, x, add, period, gamma, stat, dx32, d, M); break;
        case 6: ln_bwd_kernel<6><<<blocks, 256, 0, st>>>(dy, x, add, period, gamma, stat, dx32, d, M); break;
        case 8: ln_bwd_kernel<8><<<blocks, 256, 0, st>>>(dy, x, add, period, gamma, stat, dx32, d, M); break;
        default: return fail(2, "bad argument:", "layer norm: C / 128 must be 1, 2, 4, 6 or 8");
    }
    HOPK_LAUNCH_CHECK("ln_bwd");
    return 0;
}

extern "C" int hopk_gelu_bf16(const void* pre, void* out, long n, void* stream)
{
    HOPK_REQUIRE(n > 0 && n % 8 == 0, "gelu: element count must be a multiple of 8");
    gelu_bf16_kernel<<<grid_for((size_t)n / 8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)pre, (__nv_bfloat16*)out, (size_t)n / 8);
    HOPK_LAUNCH_CHECK("gelu");
    return 0;
}

extern "C" int hopk_bert_attn_fwd(const void* qkv, void* ctx, float* P, int B, int S, int H, int D, void* stream)
{
    HOPK_REQUIRE(B > 0 && H > 0 && S >= 1 && S <= AT_MAXS && D == AT_D, "bert attention: head dim 64, sequence <= 64");
    const int npairs = B * H, grid = (npairs + 1) / 2;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tq;
    if (int rc = make_tensor_map_bf16(&tq, qkv, (long)B * S, 3L * H * D, 3L * H * D, 64)) return rc;
    if (S <= 40) {
        HOPK_CUDA(configure_smem_once((const void*)bert_attn_fwd_kernel<40>, AT_FWD_SMEM));
        bert_attn_fwd_kernel<40><<<grid, 128, AT_FWD_SMEM, st>>>(tq, (__nv_bfloat16*)ctx, P, S, H, npairs);
    } else {
        HOPK_CUDA(configure_smem_once((const void*)bert_attn_fwd_kernel<64>, AT_FWD_SMEM));
        bert_attn_fwd_kernel<64><<<grid, 128, AT_FWD_SMEM, st>>>(tq, (__nv_bfloat16*)ctx, P, S, H, npairs);
    }
    HOPK_LAUNCH_CHECK("bert_attn_fwd");
    return 0;
}

extern "C" int hopk_bert_attn_bwd(const void* qkv, const void* dctx, void* dqkv, int B, int S, int H, int D, void* stream)
{
    HOPK_REQUIRE(B > 0 && H > 0 && S >= 1 && S <= AT_MAXS && D == AT_D, "bert attention backward: head dim 64, sequence <= 64");
    const int npairs = B * H, grid = (npairs + 1) / 2;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tq, td;
    if (int rc = make_tensor_map_bf16(&tq, qkv, (long)B * S, 3L * H * D, 3L * H * D, 64)) return rc;
    if (int rc = make_tensor_map_bf16(&td, dctx, (long)B * S, (long)H * D, (long)H * D, 64)) return rc;
    if (S <= 40) {
        HOPK_CUDA(configure_smem_once((const void*)bert_attn_bwd_kernel<40>, AT_BWD_SMEM));
        bert_attn_bwd_kernel<40><<<grid, 128, AT_BWD_SMEM, st>>>(tq, td, (__nv_bfloat16*)dqkv, S, H, npairs);
    } else {
        HOPK_CUDA(configure_smem_once((const void*)bert_attn_bwd_kernel<64>, AT_BWD_SMEM));
        bert_attn_bwd_kernel<64><<<grid, 128, AT_BWD_SMEM, st>>>(tq, td, (__nv_bfloat16*)dqkv, S, H, npairs);
    }
    HOPK_LAUNCH_CHECK("bert_attn_bwd");
    return 0;
}
