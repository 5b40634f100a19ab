// xattn.cu -- fused scaled-dot-product cross-attention of HOP's ReprogrammingLayer, fp32-exact path.
//
// Replaces reference model/HOP.py:289-299 (einsum "blhe,she->bhls" -> softmax -> dropout ->
// einsum "bhls,she->blhe") and its autograd backward.  The (B, H, L, S) score tensor that the
// reference materialises (209 MB at B=128) never leaves the SM: queries of one head are tiled
// 64 rows at a time over the flattened (b, l) axis, the S = 1500 text prototypes are streamed in
// tiles of 64 with an online softmax; K/V are batch-independent so every CTA of a head re-reads
// the same (L2-resident) tiles.  Backward recomputes the probabilities from the saved
// log-sum-exp in two passes (dQ: loop over S; dK/dV: loop over rows) -- no atomics, deterministic.
//
// Dropout mask: counter-based hash shared bit-for-bit with oracle/reprog_np.py::dropout_keep.
#include "gemm_core.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

constexpr int XT = 64;            // tile edge (rows and S)
constexpr int XLDP = XT + 4;      // leading dim of the 64 x 64 probability tiles


__device__ __forceinline__ uint32_t lowbias32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
// one hash per element pair (idx >> 1); element uses its low / high 16 bits; keep iff field >= round(p * 65536)
// (oracle/reprog_np.py::dropout_keep)
__device__ __forceinline__ bool keep_mask(uint64_t seed, uint64_t idx, uint32_t thr)
{
    uint32_t s_lo = (uint32_t)seed, s_hi = (uint32_t)(seed >> 32);
    uint64_t q = idx >> 1;
    uint32_t q_lo = (uint32_t)q, q_hi = (uint32_t)(q >> 32);
    uint32_t h = lowbias32((q_lo + lowbias32(q_hi ^ s_hi)) ^ s_lo);
    uint32_t field = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
    return field >= thr;
}

// acc[i][j] += sum_k A[r0+i][k] * B[c0 + 16 j][k]        (both row-major with k contiguous)
__device__ __forceinline__ void nt_tile(const float* __restrict__ As, int lda, int r0, const float* __restrict__ Bs, int ldb,
                                        int c0, int K, float (&acc)[4][4])
{
    for (int k = 0; k < K; k += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (r0 + i) * lda + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (c0 + 16 * j) * ldb + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
}
// acc[i][4g+j] += sum_k A[r0+i][k] * B[k][c0 + 64 g + j]   (A row-major k contiguous, B k-major); groups with c >= ncols skipped
__device__ __forceinline__ void nn_tile(const float* __restrict__ As, int lda, int r0, const float* __restrict__ Bs, int ldb,
                                        int c0, int ncols, int K, float (&acc)[4][8])
{
    const bool g0 = c0 < ncols, g1 = c0 + 64 < ncols;
    if (!g0) return;
    for (int k = 0; k < K; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (r0 + i) * lda + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 b0 = *reinterpret_cast<const float4*>(Bs + (k + kk) * ldb + c0);
            float4 b1 = g1 ? *reinterpret_cast<const float4*>(Bs + (k + kk) * ldb + c0 + 64) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
                acc[i][0] = fmaf(av, b0.x, acc[i][0]); acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av, b0.z, acc[i][2]); acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                acc[i][4] = fmaf(av, b1.x, acc[i][4]); acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                acc[i][6] = fmaf(av, b1.z, acc[i][6]); acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
        }
    }
}
// acc[i][4g+j] += sum_k A[k][r0+i] * B[k][c0 + 64 g + j]   (both k-major)
__device__ __forceinline__ void tn_tile(const float* __restrict__ As, int lda, int r0, const float* __restrict__ Bs, int ldb,
                                        int c0, int ncols, int K, float (&acc)[4][8])
{
    const bool g0 = c0 < ncols, g1 = c0 + 64 < ncols;
    if (!g0) return;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float4 a = *reinterpret_cast<const float4*>(As + k * lda + r0);
        float4 b0 = *reinterpret_cast<const float4*>(Bs + k * ldb + c0);
        float4 b1 = g1 ? *reinterpret_cast<const float4*>(Bs + k * ldb + c0 + 64) : make_float4(0, 0, 0, 0);
        float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(av[i], b0.x, acc[i][0]); acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b0.z, acc[i][2]); acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
            acc[i][4] = fmaf(av[i], b1.x, acc[i][4]); acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
            acc[i][6] = fmaf(av[i], b1.z, acc[i][6]); acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
        }
    }
}

// load a 64 x E tile of a (rows, H, E) tensor for head h into smem (row-major, ld = E + 4), zero-filling rows >= nrows
__device__ __forceinline__ void load_tile(float* dst, int ld, const float* __restrict__ src, int row0, int nrows, int H, int h,
                                          int E)
{
    const int E4 = E >> 2;
    for (int i = threadIdx.x; i < XT * E4; i += blockDim.x) {
        int r = i / E4, e4 = i - r * E4;
        float4 v = make_float4(0, 0, 0, 0);
        if (row0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(row0 + r) * H + h) * E) + e4);
        *reinterpret_cast<float4*>(dst + r * ld + 4 * e4) = v;
    }
}

__device__ __forceinline__ float group16_max(float v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group16_sum(float v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256)
xattn_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, float* __restrict__ O,
                 float* __restrict__ LSE, int M, int L, int H, int E, int S, float scale, float inv_keep, uint32_t thr,
                 uint64_t seed, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ __align__(16) float sm[];
    const int ld = E + 4;
    float* Qs = sm; float* Ks = Qs + XT * ld; float* Vs = Ks + XT * ld; float* Ps = Vs + XT * ld;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int h = blockIdx.y, m0 = blockIdx.x * XT;
    const int r0 = ty * 4, c0 = tx * 4;

    load_tile(Qs, ld, Q, m0, M, H, h, E);
    float o[4][8];
    float mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY; lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
    }
    for (int s0 = 0; s0 < S; s0 += XT) {
        __syncthreads();                       // previous tile fully consumed
        load_tile(Ks, ld, K, s0, S, H, h, E);
        load_tile(Vs, ld, V, s0, S, H, h, E);
        __syncthreads();
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
        nt_tile(Qs, ld, r0, Ks, ld, tx, E, sc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int s = s0 + tx + 16 * j;
                sc[i][j] = s < S ? sc[i][j] * scale : -INFINITY;
                mx = fmaxf(mx, sc[i][j]);
            }
            mx = group16_max(mx);
            float mnew = fmaxf(mrow[i], mx);
            float corr = expf(mrow[i] - mnew);
            float rs = 0.f;
            int m = m0 + r0 + i;
            uint64_t base = ((uint64_t)(m / L) * H + h) * (uint64_t)L + (uint64_t)(m % L);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int s = s0 + tx + 16 * j;
                float p = expf(sc[i][j] - mnew);
                rs += p;
                if (thr) p = keep_mask(seed, base * (uint64_t)S + (uint64_t)s, thr) ? p * inv_keep : 0.f;
                Ps[(r0 + i) * XLDP + tx + 16 * j] = p;
            }
            rs = group16_sum(rs);
            lrow[i] = lrow[i] * corr + rs;
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[i][j] *= corr;
        }
        __syncthreads();
        nn_tile(Ps, XLDP, r0, Vs, ld, c0, E, XT, o);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + r0 + i;
        if (m >= M) continue;
        float inv = 1.f / lrow[i];
        float* orow = O + ((size_t)m * H + h) * E;
        if (c0 < E) *reinterpret_cast<float4*>(orow + c0) = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
        if (c0 + 64 < E) *reinterpret_cast<float4*>(orow + c0 + 64) = make_float4(o[i][4] * inv, o[i][5] * inv, o[i][6] * inv, o[i][7] * inv);
        if (tx == 0) LSE[((size_t)(m / L) * H + h) * L + (m % L)] = mrow[i] + logf(lrow[i]);
    }
}

// ------------------------------------------------------------------------------------------------ backward: dQ (+ delta)
__global__ void __launch_bounds__(256)
xattn_bwd_dq_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                    const float* __restrict__ O, const float* __restrict__ LSE, const float* __restrict__ dO,
                    float* __restrict__ dQ, float* __restrict__ delta, int M, int L, int H, int E, int S, float scale,
                    float inv_keep, uint32_t thr, uint64_t seed, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ __align__(16) float sm[];
    const int ld = E + 4;
    float* Qs = sm; float* dOs = Qs + XT * ld; float* Ks = dOs + XT * ld; float* Vs = Ks + XT * ld; float* Ss = Vs + XT * ld;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int h = blockIdx.y, m0 = blockIdx.x * XT;
    const int r0 = ty * 4, c0 = tx * 4;
    load_tile(Qs, ld, Q, m0, M, H, h, E);
    load_tile(dOs, ld, dO, m0, M, H, h, E);
    __syncthreads();
    float lse[4], dl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + r0 + i;
        float part = 0.f;
        if (m < M) {
            const float* orow = O + ((size_t)m * H + h) * E;
            for (int e = c0; e < E; e += 64) {
                float4 ov = __ldg(reinterpret_cast<const float4*>(orow + e));
                float4 dv = *reinterpret_cast<const float4*>(dOs + (r0 + i) * ld + e);
                part += ov.x * dv.x + ov.y * dv.y + ov.z * dv.z + ov.w * dv.w;
            }
        }
        dl[i] = group16_sum(part);
        size_t li = m < M ? ((size_t)(m / L) * H + h) * L + (m % L) : 0;
        lse[i] = m < M ? __ldg(LSE + li) : 0.f;
        if (m < M && tx == 0) delta[li] = dl[i];
    }
    float dq[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dq[i][j] = 0.f;
    for (int s0 = 0; s0 < S; s0 += XT) {
        __syncthreads();
        load_tile(Ks, ld, K, s0, S, H, h, E);
        load_tile(Vs, ld, V, s0, S, H, h, E);
        __syncthreads();
        float sc[4][4], dp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { sc[i][j] = 0.f; dp[i][j] = 0.f; }
        nt_tile(Qs, ld, r0, Ks, ld, tx, E, sc);
        nt_tile(dOs, ld, r0, Vs, ld, tx, E, dp);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + r0 + i;
            uint64_t base = ((uint64_t)(m / L) * H + h) * (uint64_t)L + (uint64_t)(m % L);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int s = s0 + tx + 16 * j;
                float ds = 0.f;
                if (s < S && m < M) {
                    float p = expf(sc[i][j] * scale - lse[i]);
                    float d = dp[i][j];
                    if (thr) d = keep_mask(seed, base * (uint64_t)S + (uint64_t)s, thr) ? d * inv_keep : 0.f;
                    ds = p * (d - dl[i]) * scale;
                }
                Ss[(r0 + i) * XLDP + tx + 16 * j] = ds;
            }
        }
        __syncthreads();
        nn_tile(Ss, XLDP, r0, Ks, ld, c0, E, XT, dq);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + r0 + i;
        if (m >= M) continue;
        float* row = dQ + ((size_t)m * H + h) * E;
        if (c0 < E) *reinterpret_cast<float4*>(row + c0) = make_float4(dq[i][0], dq[i][1], dq[i][2], dq[i][3]);
        if (c0 + 64 < E) *reinterpret_cast<float4*>(row + c0 + 64) = make_float4(dq[i][4], dq[i][5], dq[i][6], dq[i][7]);
    }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
__global__ void __launch_bounds__(256)
xattn_bwd_dkv_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                     const float* __restrict__ LSE, const float* __restrict__ delta, const float* __restrict__ dO,
                     float* __restrict__ dK, float* __restrict__ dV, int M, int L, int H, int E, int S, float scale,
                     float inv_keep, uint32_t thr, uint64_t seed, const unsigned long long* __restrict__ epoch)
{
    seed += *epoch;                                     // dropout epoch (hopk_dropout_epoch_advance): lets a captured CUDA graph draw a fresh mask per replay
    extern __shared__ __align__(16) float sm[];
    const int ld = E + 4;
    float* Ks = sm; float* Vs = Ks + XT * ld; float* Qs = Vs + XT * ld; float* dOs = Qs + XT * ld;
    float* Ps = dOs + XT * ld; float* Ss = Ps + XT * XLDP;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int h = blockIdx.y, s0 = blockIdx.x * XT;
    const int r0 = ty * 4, c0 = tx * 4;
    load_tile(Ks, ld, K, s0, S, H, h, E);
    load_tile(Vs, ld, V, s0, S, H, h, E);
    float dk[4][8], dv[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { dk[i][j] = 0.f; dv[i][j] = 0.f; }
    for (int m0 = 0; m0 < M; m0 += XT) {
        __syncthreads();
        load_tile(Qs, ld, Q, m0, M, H, h, E);
        load_tile(dOs, ld, dO, m0, M, H, h, E);
        __syncthreads();
        float sc[4][4], dp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { sc[i][j] = 0.f; dp[i][j] = 0.f; }
        nt_tile(Qs, ld, r0, Ks, ld, tx, E, sc);
        nt_tile(dOs, ld, r0, Vs, ld, tx, E, dp);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + r0 + i;
            size_t li = m < M ? ((size_t)(m / L) * H + h) * L + (m % L) : 0;
            float lse = m < M ? __ldg(LSE + li) : 0.f;
            float dl = m < M ? __ldg(delta + li) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int s = s0 + tx + 16 * j;
                float pt = 0.f, ds = 0.f;
                if (s < S && m < M) {
                    float p = expf(sc[i][j] * scale - lse);
                    float d = dp[i][j];
                    pt = p;
                    if (thr) {
                        bool kp = keep_mask(seed, (uint64_t)li * (uint64_t)S + (uint64_t)s, thr);
                        pt = kp ? p * inv_keep : 0.f;
                        d = kp ? d * inv_keep : 0.f;
                    }
                    ds = p * (d - dl) * scale;
                }
                Ps[(r0 + i) * XLDP + tx + 16 * j] = pt;
                Ss[(r0 + i) * XLDP + tx + 16 * j] = ds;
            }
        }
        __syncthreads();
        tn_tile(Ps, XLDP, r0, dOs, ld, c0, E, XT, dv);     // dV[s][e] += sum_rows P~[row][s] dO[row][e]
        tn_tile(Ss, XLDP, r0, Qs, ld, c0, E, XT, dk);      // dK[s][e] += sum_rows dS[row][s] Q[row][e]
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int s = s0 + r0 + i;
        if (s >= S) continue;
        float* kr = dK + ((size_t)s * H + h) * E;
        float* vr = dV + ((size_t)s * H + h) * E;
        if (c0 < E) {
            *reinterpret_cast<float4*>(kr + c0) = make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
            *reinterpret_cast<float4*>(vr + c0) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
        }
        if (c0 + 64 < E) {
            *reinterpret_cast<float4*>(kr + c0 + 64) = make_float4(dk[i][4], dk[i][5], dk[i][6], dk[i][7]);
            *reinterpret_cast<float4*>(vr + c0 + 64) = make_float4(dv[i][4], dv[i][5], dv[i][6], dv[i][7]);
        }
    }
}

}  // namespace hopk
using namespace hopk;

static int xattn_check(int B, int L, int H, int E, int S, float p)
{
    HOPK_REQUIRE(B > 0 && L > 0 && H > 0 && S > 0, "xattn sizes");
    HOPK_REQUIRE(E >= 4 && E <= 128 && E % 4 == 0, "head dim must be a multiple of 4, <= 128");
    HOPK_REQUIRE(p >= 0.f && p < 1.f, "dropout p in [0,1)");
    return 0;
}

extern "C" int hopk_xattn_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int B, int L, int H, int E,
                              int S, float p_drop, uint64_t seed, void* stream)
{
    if (int rc = xattn_check(B, L, H, E, S, p_drop)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int M = B * L;
    size_t smem = ((size_t)3 * XT * (E + 4) + XT * XLDP) * sizeof(float);
    HOPK_CUDA(cudaFuncSetAttribute(xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t thr = (uint32_t)lrintf(p_drop * 65536.f);
    float inv_keep = 65536.f / (65536.f - (float)thr);
    dim3 grid(cdiv(M, XT), H);
    const unsigned long long* epoch = drop_epoch_ptr();
    HOPK_REQUIRE(epoch != nullptr, "dropout epoch symbol");
    xattn_fwd_kernel<<<grid, 256, smem, st>>>(q, k, v, o, lse, M, L, H, E, S, 1.f / sqrtf((float)E), inv_keep, thr, seed, epoch);
    HOPK_LAUNCH_CHECK("xattn_fwd");
    return 0;
}

extern "C" int hopk_xattn_bwd(const float* q, const float* k, const float* v, const float* o, const float* lse,
                              const float* dout, float* dq, float* dk, float* dv, float* delta, int B, int L, int H, int E,
                              int S, float p_drop, uint64_t seed, void* stream)
{
    if (int rc = xattn_check(B, L, H, E, S, p_drop)) return rc;
    HOPK_REQUIRE(delta != nullptr, "delta scratch (B*H*L floats) required");
    cudaStream_t st = (cudaStream_t)stream;
    int M = B * L;
    uint32_t thr = (uint32_t)lrintf(p_drop * 65536.f);
    float inv_keep = 65536.f / (65536.f - (float)thr);
    float scale = 1.f / sqrtf((float)E);
    size_t smem1 = ((size_t)4 * XT * (E + 4) + XT * XLDP) * sizeof(float);
    HOPK_CUDA(cudaFuncSetAttribute(xattn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    const unsigned long long* epoch = drop_epoch_ptr();
    HOPK_REQUIRE(epoch != nullptr, "dropout epoch symbol");
    xattn_bwd_dq_kernel<<<dim3(cdiv(M, XT), H), 256, smem1, st>>>(q, k, v, o, lse, dout, dq, delta, M, L, H, E, S, scale,
                                                                  inv_keep, thr, seed, epoch);
    HOPK_LAUNCH_CHECK("xattn_bwd_dq");
    size_t smem2 = ((size_t)4 * XT * (E + 4) + 2 * XT * XLDP) * sizeof(float);
    HOPK_CUDA(cudaFuncSetAttribute(xattn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    xattn_bwd_dkv_kernel<<<dim3(cdiv(S, XT), H), 256, smem2, st>>>(q, k, v, lse, delta, dout, dk, dv, M, L, H, E, S, scale,
                                                                   inv_keep, thr, seed, epoch);
    HOPK_LAUNCH_CHECK("xattn_bwd_dkv");
    return 0;
}
