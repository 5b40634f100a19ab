// bert.cu -- the small kernels of the frozen BERT-6 encoder between its dense GEMMs (reference model/HOP.py:202-204:
// `self.llm_model(inputs_embeds=llama_enc_out).last_hidden_state`; the encoder is frozen, HOP.py:90-91, so backward only
// needs dX).  The GEMMs themselves (QKV / output / FFN projections and their dX products) run on gemm_tma.cu; here:
//   layer norm forward / backward (dX)        BertEmbeddings.LayerNorm, BertSelfOutput.LayerNorm, BertOutput.LayerNorm
//   self-attention forward / backward         12 heads x 64, sequence 34: one CTA per (sample, head), everything in smem
//   GELU                                      BertIntermediate (exact erf form)
// dtype-1 arithmetic: bf16 operands for the tensor-core GEMMs, fp32 statistics / softmax / residual stream.
#include <cuda_bf16.h>
#include "common.cuh"
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
            o[j] = bf2(0.5f * a * (1.f + erff(a * 0.70710678118654752f)), 0.5f * b * (1.f + erff(b * 0.70710678118654752f)));
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------- self-attention, one CTA per (sample, head)
// qkv: bf16 [B*S][3*H*D] rows = (b, t): [q heads | k heads | v heads]; ctx: bf16 [B*S][H*D]; P: fp32 [B*H][S][S] (saved)
constexpr int AT_MAXS = 64, AT_D = 64;

__global__ void __launch_bounds__(128) bert_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx,
                                                            float* __restrict__ P, int S, int H)
{
    extern __shared__ float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    float* q = sm;                          // [S][D+1]
    float* k = q + S * (AT_D + 1);          // [S][D+1]
    float* v = k + S * (AT_D + 1);          // [S][D]
    float* p = v + S * AT_D;                // [S][S+1]
    const int ld = 3 * H * AT_D;
    for (int i = threadIdx.x; i < S * (AT_D / 2); i += blockDim.x) {
        const int t = i / (AT_D / 2), c = (i % (AT_D / 2)) * 2;
        const __nv_bfloat16* row = qkv + (size_t)(b * S + t) * ld + h * AT_D + c;
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(row);
        const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(row + H * AT_D);
        const __nv_bfloat162 cc = *reinterpret_cast<const __nv_bfloat162*>(row + 2 * H * AT_D);
        q[t * (AT_D + 1) + c] = __low2float(a); q[t * (AT_D + 1) + c + 1] = __high2float(a);
        k[t * (AT_D + 1) + c] = __low2float(bb); k[t * (AT_D + 1) + c + 1] = __high2float(bb);
        v[t * AT_D + c] = __low2float(cc); v[t * AT_D + c + 1] = __high2float(cc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int r = i / S, c = i % S;
        float acc = 0.f;
#pragma unroll 16
        for (int e = 0; e < AT_D; ++e) acc = fmaf(q[r * (AT_D + 1) + e], k[c * (AT_D + 1) + e], acc);
        p[r * (S + 1) + c] = acc * 0.125f;                      // 1 / sqrt(64)
    }
    __syncthreads();
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {   // one warp per row
        const int lane = threadIdx.x & 31;
        float m = -INFINITY;
        for (int c = lane; c < S; c += 32) m = fmaxf(m, p[r * (S + 1) + c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < S; c += 32) { const float e = __expf(p[r * (S + 1) + c] - m); p[r * (S + 1) + c] = e; s += e; }
        s = 1.f / warp_sum(s);
        for (int c = lane; c < S; c += 32) {
            const float pv = p[r * (S + 1) + c] * s;
            p[r * (S + 1) + c] = pv;
            if (P) P[((size_t)blockIdx.x * S + r) * S + c] = pv;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * (AT_D / 2); i += blockDim.x) {
        const int r = i / (AT_D / 2), c = (i % (AT_D / 2)) * 2;
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < S; ++j) {
            const float pv = p[r * (S + 1) + j];
            a0 = fmaf(pv, v[j * AT_D + c], a0); a1 = fmaf(pv, v[j * AT_D + c + 1], a1);
        }
        *reinterpret_cast<uint32_t*>(ctx + (size_t)(b * S + r) * (H * AT_D) + h * AT_D + c) = bf2(a0, a1);
    }
}

// dqkv from dctx: dV = P^T dO, dP = dO V^T, dS = P * (dP - rowsum(dP * P)), dQ = dS K / 8, dK = dS^T Q / 8
__global__ void __launch_bounds__(128) bert_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dctx,
                                                            const float* __restrict__ P, __nv_bfloat16* __restrict__ dqkv, int S, int H)
{
    extern __shared__ float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    float* q = sm;                          // [S][D+1]
    float* k = q + S * (AT_D + 1);
    float* v = k + S * (AT_D + 1);          // [S][D+1]
    float* d = v + S * (AT_D + 1);          // dO [S][D+1]
    float* p = d + S * (AT_D + 1);          // P  [S][S+1]
    float* ds = p + S * (S + 1);            // dS [S][S+1]
    const int ld = 3 * H * AT_D;
    for (int i = threadIdx.x; i < S * (AT_D / 2); i += blockDim.x) {
        const int t = i / (AT_D / 2), c = (i % (AT_D / 2)) * 2;
        const __nv_bfloat16* row = qkv + (size_t)(b * S + t) * ld + h * AT_D + c;
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(row);
        const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(row + H * AT_D);
        const __nv_bfloat162 cc = *reinterpret_cast<const __nv_bfloat162*>(row + 2 * H * AT_D);
        const __nv_bfloat162 dd = *reinterpret_cast<const __nv_bfloat162*>(dctx + (size_t)(b * S + t) * (H * AT_D) + h * AT_D + c);
        q[t * (AT_D + 1) + c] = __low2float(a); q[t * (AT_D + 1) + c + 1] = __high2float(a);
        k[t * (AT_D + 1) + c] = __low2float(bb); k[t * (AT_D + 1) + c + 1] = __high2float(bb);
        v[t * (AT_D + 1) + c] = __low2float(cc); v[t * (AT_D + 1) + c + 1] = __high2float(cc);
        d[t * (AT_D + 1) + c] = __low2float(dd); d[t * (AT_D + 1) + c + 1] = __high2float(dd);
    }
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) p[(i / S) * (S + 1) + i % S] = P[(size_t)blockIdx.x * S * S + i];
    __syncthreads();
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {                 // dP = dO V^T
        const int r = i / S, c = i % S;
        float acc = 0.f;
#pragma unroll 16
        for (int e = 0; e < AT_D; ++e) acc = fmaf(d[r * (AT_D + 1) + e], v[c * (AT_D + 1) + e], acc);
        ds[r * (S + 1) + c] = acc;
    }
    __syncthreads();
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {           // dS = P * (dP - sum(dP * P)) / 8
        const int lane = threadIdx.x & 31;
        float s = 0.f;
        for (int c = lane; c < S; c += 32) s += ds[r * (S + 1) + c] * p[r * (S + 1) + c];
        s = warp_sum(s);
        for (int c = lane; c < S; c += 32) ds[r * (S + 1) + c] = p[r * (S + 1) + c] * (ds[r * (S + 1) + c] - s) * 0.125f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * (AT_D / 2); i += blockDim.x) {
        const int r = i / (AT_D / 2), c = (i % (AT_D / 2)) * 2;
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int j = 0; j < S; ++j) {
            const float s_rj = ds[r * (S + 1) + j], s_jr = ds[j * (S + 1) + r], p_jr = p[j * (S + 1) + r];
            q0 = fmaf(s_rj, k[j * (AT_D + 1) + c], q0); q1 = fmaf(s_rj, k[j * (AT_D + 1) + c + 1], q1);
            k0 = fmaf(s_jr, q[j * (AT_D + 1) + c], k0); k1 = fmaf(s_jr, q[j * (AT_D + 1) + c + 1], k1);
            v0 = fmaf(p_jr, d[j * (AT_D + 1) + c], v0); v1 = fmaf(p_jr, d[j * (AT_D + 1) + c + 1], v1);
        }
        __nv_bfloat16* row = dqkv + (size_t)(b * S + r) * ld + h * AT_D + c;
        *reinterpret_cast<uint32_t*>(row) = bf2(q0, q1);
        *reinterpret_cast<uint32_t*>(row + H * AT_D) = bf2(k0, k1);
        *reinterpret_cast<uint32_t*>(row + 2 * H * AT_D) = bf2(v0, v1);
    }
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
    const size_t smem = ((size_t)2 * S * (AT_D + 1) + (size_t)S * AT_D + (size_t)S * (S + 1)) * sizeof(float);
    HOPK_CUDA(configure_smem_once((const void*)bert_attn_fwd_kernel, 96 * 1024));
    bert_attn_fwd_kernel<<<B * H, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)ctx, P, S, H);
    HOPK_LAUNCH_CHECK("bert_attn_fwd");
    return 0;
}

extern "C" int hopk_bert_attn_bwd(const void* qkv, const void* dctx, const float* P, void* dqkv, int B, int S, int H, int D, void* stream)
{
    HOPK_REQUIRE(B > 0 && H > 0 && S >= 1 && S <= AT_MAXS && D == AT_D && P, "bert attention backward: head dim 64, sequence <= 64, saved P");
    const size_t smem = ((size_t)4 * S * (AT_D + 1) + (size_t)2 * S * (S + 1)) * sizeof(float);
    HOPK_CUDA(configure_smem_once((const void*)bert_attn_bwd_kernel, 128 * 1024));
    bert_attn_bwd_kernel<<<B * H, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dctx, P, (__nv_bfloat16*)dqkv, S, H);
    HOPK_LAUNCH_CHECK("bert_attn_bwd");
    return 0;
}
