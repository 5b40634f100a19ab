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
// Every product is register-tiled (2 rows x 2 columns of the S x S matrices, 2 rows x 4 columns of the S x D ones) with
// 128-bit shared-memory reads along the contraction index: one LDS.128 per 4-8 FMAs instead of two LDS per FMA (the first
// version was shared-memory-bandwidth bound: 78 us per call at B = 128).
constexpr int AT_MAXS = 64, AT_D = 64, AT_LD = AT_D + 4;      // fp32 row pitch 68: 16-byte aligned rows

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// out[r][c] = sum_e X[r][e] Y[c][e] for r, c < S (X, Y: [S][AT_LD]); 2 x 2 register tiles; out pitch ldo; scaled
__device__ __forceinline__ void tile_xyT(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ out, int ldo, int S, float scale)
{
    const int TS = (S + 1) >> 1;
    for (int i = threadIdx.x; i < TS * TS; i += blockDim.x) {
        const int r0 = (i / TS) * 2, c0 = (i % TS) * 2;
        const int r1 = min(r0 + 1, S - 1), c1 = min(c0 + 1, S - 1);
        float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll 4
        for (int e = 0; e < AT_D; e += 4) {
            const float4 x0 = *reinterpret_cast<const float4*>(X + r0 * AT_LD + e), x1 = *reinterpret_cast<const float4*>(X + r1 * AT_LD + e);
            const float4 y0 = *reinterpret_cast<const float4*>(Y + c0 * AT_LD + e), y1 = *reinterpret_cast<const float4*>(Y + c1 * AT_LD + e);
            a00 += dot4(x0, y0); a01 += dot4(x0, y1); a10 += dot4(x1, y0); a11 += dot4(x1, y1);
        }
        out[r0 * ldo + c0] = a00 * scale;
        if (c0 + 1 < S) out[r0 * ldo + c0 + 1] = a01 * scale;
        if (r0 + 1 < S) {
            out[(r0 + 1) * ldo + c0] = a10 * scale;
            if (c0 + 1 < S) out[(r0 + 1) * ldo + c0 + 1] = a11 * scale;
        }
    }
}

// acc[2][4] = sum_j W(r, j) * Z[j][c..c+3] for rows r0, r0 + 1: W(r, j) = Wm[r * ldw + j] or (transposed) Wm[j * ldw + r]
template <bool TRANS>
__device__ __forceinline__ void tile_wz(const float* __restrict__ Wm, int ldw, const float* __restrict__ Z, int S, int r0, int r1, int c,
                                        float (&acc)[2][4])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[0][q] = 0.f; acc[1][q] = 0.f; }
    for (int j = 0; j < S; ++j) {
        const float w0 = TRANS ? Wm[j * ldw + r0] : Wm[r0 * ldw + j], w1 = TRANS ? Wm[j * ldw + r1] : Wm[r1 * ldw + j];
        const float4 z = *reinterpret_cast<const float4*>(Z + j * AT_LD + c);
        acc[0][0] = fmaf(w0, z.x, acc[0][0]); acc[0][1] = fmaf(w0, z.y, acc[0][1]); acc[0][2] = fmaf(w0, z.z, acc[0][2]); acc[0][3] = fmaf(w0, z.w, acc[0][3]);
        acc[1][0] = fmaf(w1, z.x, acc[1][0]); acc[1][1] = fmaf(w1, z.y, acc[1][1]); acc[1][2] = fmaf(w1, z.z, acc[1][2]); acc[1][3] = fmaf(w1, z.w, acc[1][3]);
    }
}

__device__ __forceinline__ void load_head(float* dst, const __nv_bfloat16* __restrict__ src, size_t ld, int S)
{
    for (int i = threadIdx.x; i < S * (AT_D / 8); i += blockDim.x) {          // 8 bf16 (16 bytes) per thread
        const int t = i / (AT_D / 8), c = (i % (AT_D / 8)) * 8;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (size_t)t * ld + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        float* d = dst + t * AT_LD + c;
        *reinterpret_cast<float4*>(d) = make_float4(__low2float(h[0]), __high2float(h[0]), __low2float(h[1]), __high2float(h[1]));
        *reinterpret_cast<float4*>(d + 4) = make_float4(__low2float(h[2]), __high2float(h[2]), __low2float(h[3]), __high2float(h[3]));
    }
}

__global__ void __launch_bounds__(128) bert_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx,
                                                            float* __restrict__ P, int S, int H)
{
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    float* q = sm;                          // [S][AT_LD]
    float* k = q + S * AT_LD;
    float* v = k + S * AT_LD;
    float* p = v + S * AT_LD;               // [S][S+1]
    const size_t ld = (size_t)3 * H * AT_D;
    const __nv_bfloat16* base = qkv + (size_t)b * S * ld + h * AT_D;
    load_head(q, base, ld, S);
    load_head(k, base + H * AT_D, ld, S);
    load_head(v, base + 2 * H * AT_D, ld, S);
    __syncthreads();
    tile_xyT(q, k, p, S + 1, S, 0.125f);                        // scores / sqrt(64)
    __syncthreads();
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {   // softmax: one warp per row
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
    const int TR = (S + 1) >> 1;
    for (int i = threadIdx.x; i < TR * (AT_D / 4); i += blockDim.x) {            // ctx = P V, 2 rows x 4 columns per thread
        const int r0 = (i / (AT_D / 4)) * 2, c = (i % (AT_D / 4)) * 4, r1 = min(r0 + 1, S - 1);
        float acc[2][4];
        tile_wz<false>(p, S + 1, v, S, r0, r1, c, acc);
        __nv_bfloat16* o = ctx + (size_t)(b * S + r0) * (H * AT_D) + h * AT_D + c;
        *reinterpret_cast<uint2*>(o) = make_uint2(bf2(acc[0][0], acc[0][1]), bf2(acc[0][2], acc[0][3]));
        if (r0 + 1 < S) *reinterpret_cast<uint2*>(o + H * AT_D) = make_uint2(bf2(acc[1][0], acc[1][1]), bf2(acc[1][2], acc[1][3]));
    }
}

// dqkv from dctx: dV = P^T dO, dP = dO V^T, dS = P * (dP - rowsum(dP * P)) / 8, dQ = dS K, dK = dS^T Q
__global__ void __launch_bounds__(128) bert_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dctx,
                                                            const float* __restrict__ P, __nv_bfloat16* __restrict__ dqkv, int S, int H)
{
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    float* q = sm;                          // [S][AT_LD]
    float* k = q + S * AT_LD;
    float* v = k + S * AT_LD;
    float* d = v + S * AT_LD;               // dO
    float* p = d + S * AT_LD;               // P  [S][S+1]
    float* ds = p + S * (S + 1);            // dS [S][S+1]
    const size_t ld = (size_t)3 * H * AT_D;
    const __nv_bfloat16* base = qkv + (size_t)b * S * ld + h * AT_D;
    load_head(q, base, ld, S);
    load_head(k, base + H * AT_D, ld, S);
    load_head(v, base + 2 * H * AT_D, ld, S);
    load_head(d, dctx + (size_t)b * S * (H * AT_D) + h * AT_D, (size_t)H * AT_D, S);
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) p[(i / S) * (S + 1) + i % S] = __ldg(P + (size_t)blockIdx.x * S * S + i);
    __syncthreads();
    tile_xyT(d, v, ds, S + 1, S, 1.f);                          // dP = dO V^T
    __syncthreads();
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {
        const int lane = threadIdx.x & 31;
        float s = 0.f;
        for (int c = lane; c < S; c += 32) s += ds[r * (S + 1) + c] * p[r * (S + 1) + c];
        s = warp_sum(s);
        for (int c = lane; c < S; c += 32) ds[r * (S + 1) + c] = p[r * (S + 1) + c] * (ds[r * (S + 1) + c] - s) * 0.125f;
    }
    __syncthreads();
    const int TR = (S + 1) >> 1;
    for (int i = threadIdx.x; i < TR * (AT_D / 4); i += blockDim.x) {
        const int r0 = (i / (AT_D / 4)) * 2, c = (i % (AT_D / 4)) * 4, r1 = min(r0 + 1, S - 1);
        float aq[2][4], ak[2][4], av[2][4];
        tile_wz<false>(ds, S + 1, k, S, r0, r1, c, aq);         // dQ = dS K
        tile_wz<true>(ds, S + 1, q, S, r0, r1, c, ak);          // dK = dS^T Q
        tile_wz<true>(p, S + 1, d, S, r0, r1, c, av);           // dV = P^T dO
        __nv_bfloat16* o = dqkv + (size_t)(b * S + r0) * ld + h * AT_D + c;
        *reinterpret_cast<uint2*>(o) = make_uint2(bf2(aq[0][0], aq[0][1]), bf2(aq[0][2], aq[0][3]));
        *reinterpret_cast<uint2*>(o + H * AT_D) = make_uint2(bf2(ak[0][0], ak[0][1]), bf2(ak[0][2], ak[0][3]));
        *reinterpret_cast<uint2*>(o + 2 * H * AT_D) = make_uint2(bf2(av[0][0], av[0][1]), bf2(av[0][2], av[0][3]));
        if (r0 + 1 < S) {
            o += ld;
            *reinterpret_cast<uint2*>(o) = make_uint2(bf2(aq[1][0], aq[1][1]), bf2(aq[1][2], aq[1][3]));
            *reinterpret_cast<uint2*>(o + H * AT_D) = make_uint2(bf2(ak[1][0], ak[1][1]), bf2(ak[1][2], ak[1][3]));
            *reinterpret_cast<uint2*>(o + 2 * H * AT_D) = make_uint2(bf2(av[1][0], av[1][1]), bf2(av[1][2], av[1][3]));
        }
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
    const size_t smem = ((size_t)3 * S * AT_LD + (size_t)S * (S + 1)) * sizeof(float) + 16;
    HOPK_CUDA(configure_smem_once((const void*)bert_attn_fwd_kernel, 96 * 1024));
    bert_attn_fwd_kernel<<<B * H, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)ctx, P, S, H);
    HOPK_LAUNCH_CHECK("bert_attn_fwd");
    return 0;
}

extern "C" int hopk_bert_attn_bwd(const void* qkv, const void* dctx, const float* P, void* dqkv, int B, int S, int H, int D, void* stream)
{
    HOPK_REQUIRE(B > 0 && H > 0 && S >= 1 && S <= AT_MAXS && D == AT_D && P, "bert attention backward: head dim 64, sequence <= 64, saved P");
    const size_t smem = ((size_t)4 * S * AT_LD + (size_t)2 * S * (S + 1)) * sizeof(float) + 16;
    HOPK_CUDA(configure_smem_once((const void*)bert_attn_bwd_kernel, 128 * 1024));
    bert_attn_bwd_kernel<<<B * H, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dctx, P, (__nv_bfloat16*)dqkv, S, H);
    HOPK_LAUNCH_CHECK("bert_attn_bwd");
    return 0;
}
