// gwnet.cu -- Graph-WaveNet block of HOP, forward and backward, fp32-exact path (dtype 0).
//
// Replaces the ATen/cuDNN/cuBLAS sequence behind gwnet.forward (reference model/gwnet.py:143-249)
// and its autograd backward.  Maths: SURVEY.md Appendix A.  All activations live in the "rows"
// layout (see functors.cuh); per layer i with dilation d, T_out = T_in - d:
//
//   forward   gate GEMM      [x_t | x_{t+d}] (K = 2C)  ->  tanh(f) * sigmoid(g)      (BN_{i-1} folded into the load,
//                                                                                     skip slice written to ycat)
//             node mix       x1 = A^T y, x2 = (A^2)^T y                               (V x V support in smem)
//             mlp GEMM       [y | x1 | x2] (K = 3C) + bias + residual x_{t+d} -> u    (+ per-channel sum / sum^2)
//             bn finalize    mean/rstd, scale/shift for the next load, running stats
//   head      one concat GEMM over the last T_l steps of all layers' y -> relu -> end1 -> relu -> end2
//
//   backward  per layer, reverse: BN-backward elementwise (du), node mix with A, A^2 (P1, P2),
//             mlp weight-grad GEMM (+bias column), dA Gram accumulation (M1, M2), dy GEMM fused with the
//             gate derivative, gate weight-grad GEMM (+bias columns), dx GEMM (4C contraction over taps x {f,g})
//             fused with the residual gradient and the BatchNorm-backward sums of the previous layer.
//   dA = M1 + A^T M2 + M2 A^T is finished once, then pushed through softmax(relu(E1 E2)).
#include <stdlib.h>
#include <mutex>
#include "functors.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

// ============================================================================ layouts
struct GwLayout {
    int Tlen[HOPK_MAX_LAYERS + 1];
    int Tp, pad, Tl;
    size_t x0, u[HOPK_MAX_LAYERS], tf[HOPK_MAX_LAYERS], sg[HOPK_MAX_LAYERS], y[HOPK_MAX_LAYERS],
        x1[HOPK_MAX_LAYERS], x2[HOPK_MAX_LAYERS];
    size_t ycat, r0, r1, orow, ss, mr, A, A2, At, A2t, Z, stats, pack, ticket, total;
    // scratch (backward)
    // du / G / df / dg are kept per layer so that the weight-gradient and dA kernels can run on side streams while the
    // main stream walks down the layers
    size_t s_du[HOPK_MAX_LAYERS], s_g[HOPK_MAX_LAYERS], s_df[HOPK_MAX_LAYERS], s_dg[HOPK_MAX_LAYERS];
    size_t s_p1[HOPK_MAX_LAYERS], s_p2[HOPK_MAX_LAYERS], s_dxa, s_dxb, s_dorow, s_de1, s_dskip, s_dycat, s_bnsum, s_m12, s_dA, s_total;
};

static int receptive_field(const HopkGwnetShape* s)
{
    int rf = 1;
    for (int i = 0; i < s->L; ++i) rf += s->dil[i];
    return rf;
}

static size_t bump(size_t& cur, size_t bytes)
{
    size_t at = cur;
    cur += (bytes + 255) & ~size_t(255);
    return at;
}

static GwLayout make_layout(const HopkGwnetShape* s)
{
    GwLayout g;
    memset(&g, 0, sizeof(g));
    int rf = receptive_field(s);
    g.Tp = s->T < rf ? rf : s->T;
    g.pad = g.Tp - s->T;
    g.Tlen[0] = g.Tp;
    for (int i = 0; i < s->L; ++i) g.Tlen[i + 1] = g.Tlen[i] - s->dil[i];
    g.Tl = g.Tlen[s->L];
    const size_t f = sizeof(float);
    const size_t BV = (size_t)s->B * s->V;
    size_t cur = 0;
    g.x0 = bump(cur, BV * g.Tp * s->C * f);
    for (int i = 0; i < s->L; ++i) {
        size_t n = BV * g.Tlen[i + 1] * s->C * f;
        g.u[i] = bump(cur, n); g.tf[i] = bump(cur, n); g.sg[i] = bump(cur, n);
        g.y[i] = bump(cur, n); g.x1[i] = bump(cur, n); g.x2[i] = bump(cur, n);
    }
    g.ycat = bump(cur, BV * g.Tl * s->L * s->C * f);
    g.r0 = bump(cur, BV * g.Tl * s->S * f);
    g.r1 = bump(cur, BV * g.Tl * s->E * f);
    g.orow = bump(cur, BV * g.Tl * s->out_dim * f);
    g.ss = bump(cur, (size_t)(s->L + 1) * 2 * s->C * f);
    g.mr = bump(cur, (size_t)s->L * 2 * s->C * f);
    size_t vv = (size_t)s->V * s->V * f;
    g.A = bump(cur, vv); g.A2 = bump(cur, vv); g.At = bump(cur, vv); g.A2t = bump(cur, vv); g.Z = bump(cur, vv);
    g.stats = bump(cur, (size_t)s->L * 2 * s->C * sizeof(double));
    g.pack = bump(cur, (size_t)s->L * 57344 + 32768 + 1024);  // packed bf16 weight slabs + diffusion operator of the fused layer kernel
    g.ticket = bump(cur, (size_t)HOPK_MAX_LAYERS * sizeof(unsigned int));
    g.total = cur;

    cur = 0;
    size_t nmax = BV * g.Tlen[0] * s->C * f;
    for (int i = 0; i < s->L; ++i) {
        size_t n = BV * g.Tlen[i + 1] * s->C * f;
        g.s_p1[i] = bump(cur, n); g.s_p2[i] = bump(cur, n);
        g.s_du[i] = bump(cur, n); g.s_g[i] = bump(cur, 2 * n); g.s_df[i] = bump(cur, n); g.s_dg[i] = bump(cur, n);
    }
    g.s_dxa = bump(cur, nmax); g.s_dxb = bump(cur, nmax);
    g.s_dorow = bump(cur, BV * g.Tl * s->out_dim * f);
    g.s_de1 = bump(cur, BV * g.Tl * s->E * f);
    g.s_dskip = bump(cur, BV * g.Tl * s->S * f);
    g.s_dycat = bump(cur, BV * g.Tl * s->L * s->C * f);
    g.s_bnsum = bump(cur, (size_t)(s->L + 1) * 2 * s->C * sizeof(double));   // + one slot: column sums of the start-conv output gradient
    g.s_m12 = bump(cur, 2 * vv);
    g.s_dA = bump(cur, vv);
    g.s_total = cur;
    return g;
}

// ============================================================================ small kernels
// A = softmax(relu(E1 E2), dim=1); also A^2, transposes and Z = E1 E2 (gwnet.py:163)
__global__ void adp_fwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int V, int R,
                               float* A, float* A2, float* At, float* A2t, float* Z)
{
    extern __shared__ float sm[];
    float* sA = sm;                       // V*V
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
        int v = idx / V, w = idx % V;
        float z = 0.f;
        for (int r = 0; r < R; ++r) z = fmaf(e1[v * R + r], e2[r * V + w], z);
        Z[idx] = z;
        sA[idx] = fmaxf(z, 0.f);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float mx = -INFINITY;
        for (int w = 0; w < V; ++w) mx = fmaxf(mx, sA[v * V + w]);
        float sum = 0.f;
        for (int w = 0; w < V; ++w) { float e = expf(sA[v * V + w] - mx); sA[v * V + w] = e; sum += e; }
        float inv = 1.f / sum;
        for (int w = 0; w < V; ++w) sA[v * V + w] *= inv;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
        int v = idx / V, w = idx % V;
        float a2 = 0.f;
        for (int k = 0; k < V; ++k) a2 = fmaf(sA[v * V + k], sA[k * V + w], a2);
        A[idx] = sA[idx];
        At[w * V + v] = sA[idx];
        A2[idx] = a2;
        A2t[w * V + v] = a2;
    }
}

// dA = M1 + A^T M2 + M2 A^T ; dR = A*(dA - rowsum(dA*A)) ; dZ = dR*[Z>0] ; dE1 = dZ E2^T ; dE2 = E1^T dZ
// If M2 == nullptr, M1 already holds dA (standalone nconv use).
__global__ void adp_bwd_kernel(const float* __restrict__ e1, const float* __restrict__ e2, const float* __restrict__ A,
                               const float* __restrict__ Z, const float* __restrict__ M1, const float* __restrict__ M2,
                               int V, int R, float* de1, float* de2)
{
    extern __shared__ float sm[];
    float* dA = sm;               // V*V, becomes dZ
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
        int v = idx / V, w = idx % V;
        float acc = M1[idx];
        if (M2) {
            for (int k = 0; k < V; ++k) acc += A[k * V + v] * M2[k * V + w] + M2[v * V + k] * A[w * V + k];
        }
        dA[idx] = acc;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float dot = 0.f;
        for (int w = 0; w < V; ++w) dot += dA[v * V + w] * A[v * V + w];
        for (int w = 0; w < V; ++w) {
            float dr = A[v * V + w] * (dA[v * V + w] - dot);
            dA[v * V + w] = Z[v * V + w] > 0.f ? dr : 0.f;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < V * R; idx += blockDim.x) {
        int v = idx / R, r = idx % R;
        float a = 0.f;
        for (int w = 0; w < V; ++w) a += dA[v * V + w] * e2[r * V + w];
        de1[idx] = a;                              // (V, R)
        int r2 = idx / V, w2 = idx % V;            // reuse the same index space for (R, V)
        float b = 0.f;
        for (int k = 0; k < V; ++k) b += e1[k * R + r2] * dA[k * V + w2];
        de2[idx] = b;
    }
}

__global__ void sums_to_float_kernel(const double* __restrict__ sums, float* __restrict__ out, int C)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) out[c] = (float)sums[c];
}

__global__ void fill_identity_kernel(float* ss, int C)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { ss[c] = 1.f; ss[C + c] = 0.f; }
}

// BatchNorm2d finalize (gwnet.py:120,237): statistics -> mean/rstd, folded scale/shift, running stats
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* rmean, float* rvar, long long* nbt,
                                   float* mr, float* ss, int C, int training, float BN_MOM, float BN_EPS)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, rstd;
    if (training) {
        double m = stats[c] / count;
        double var = stats[C + c] / count - m * m;
        if (var < 0) var = 0;
        mean = (float)m;
        rstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
        double unb = count > 1 ? var * count / (count - 1) : var;
        rmean[c] = (1.f - BN_MOM) * rmean[c] + BN_MOM * mean;
        rvar[c] = (1.f - BN_MOM) * rvar[c] + BN_MOM * (float)unb;
        if (c == 0 && nbt) *nbt += 1;
    } else {
        mean = rmean[c];
        rstd = 1.f / sqrtf(rvar[c] + BN_EPS);
    }
    mr[c] = mean; mr[C + c] = rstd;
    float sc = gamma[c] * rstd;
    ss[c] = sc; ss[C + c] = beta[c] - mean * sc;
}

// node mix: out1[(g,w)] = sum_v M1[v][w] in[(g,v)], out2 likewise with M2 (gwnet.py:12-14 applied on rows layout)
__global__ void node_mix_kernel(const float* __restrict__ in, const float* __restrict__ M1, const float* __restrict__ M2,
                                float* __restrict__ out1, float* __restrict__ out2, int groups, int V, int C, int gpb)
{
    extern __shared__ float sm[];
    const int vv4 = (V * V + 3) & ~3;                                      // keep sx 16-byte aligned
    float* s1 = sm; float* s2 = s1 + vv4; float* sx = s2 + vv4;            // sx: gpb*V*C
    for (int i = threadIdx.x; i < V * V; i += blockDim.x) { s1[i] = M1[i]; s2[i] = M2[i]; }
    int g0 = blockIdx.x * gpb;
    int ng = min(gpb, groups - g0);
    int C4 = C >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in + (size_t)g0 * V * C);
    float4* sx4 = reinterpret_cast<float4*>(sx);
    for (int i = threadIdx.x; i < ng * V * C4; i += blockDim.x) sx4[i] = in4[i];
    __syncthreads();
    float4* o1 = reinterpret_cast<float4*>(out1 + (size_t)g0 * V * C);
    float4* o2 = reinterpret_cast<float4*>(out2 + (size_t)g0 * V * C);
    for (int i = threadIdx.x; i < ng * V * C4; i += blockDim.x) {
        int c4 = i % C4; int gw = i / C4; int w = gw % V; int g = gw / V;
        float4 a = make_float4(0, 0, 0, 0), b = a;
        const float4* row = sx4 + (size_t)g * V * C4 + c4;
        for (int v = 0; v < V; ++v) {
            float4 x = row[(size_t)v * C4];
            float m1 = s1[v * V + w], m2 = s2[v * V + w];
            a.x = fmaf(m1, x.x, a.x); a.y = fmaf(m1, x.y, a.y); a.z = fmaf(m1, x.z, a.z); a.w = fmaf(m1, x.w, a.w);
            b.x = fmaf(m2, x.x, b.x); b.y = fmaf(m2, x.y, b.y); b.z = fmaf(m2, x.z, b.z); b.w = fmaf(m2, x.w, b.w);
        }
        o1[i] = a; o2[i] = b;
    }
}

// dA Gram accumulation: M1[v][w] += sum_{g,c} Y[(g,v)][c] G[(g,w)][c], M2 with G[..][C + c]   (G has ld 2C)
// Persistent CTAs stride over chunks of GRAM_GPI node groups; thread = (pair slot, channel slice); partial sums stay in
// registers across the whole loop, are combined in shared memory and leave the CTA as one atomic per matrix entry.
constexpr int GRAM_GPI = 4;                  // node groups staged per iteration (fewer when V * C is large: smem budget)
constexpr int GRAM_MAXP = 8;                 // pairs per thread: V*V <= 8*256  (V <= 45)
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ TF, const float* __restrict__ SG, const float* __restrict__ G,
                                                   float* __restrict__ M12, int groups, int V, int C, int GPI)
{
    extern __shared__ float sm[];
    const int VV = V * V;
    const int ldy = C + 1, ldg = 2 * C + 1;
    float* sy = sm;                                   // GPI * V * ldy
    float* sg = sy + GPI * V * ldy;              // GPI * V * ldg
    float* acc = sg + GPI * V * ldg;             // 2 * VV
    const int NS = VV >= 256 ? 1 : 256 / VV;          // channel slices per pair
    const int slice = VV >= 256 ? 0 : (int)threadIdx.x / VV;
    const int p0 = VV >= 256 ? (int)threadIdx.x : (int)threadIdx.x % VV;
    const bool active = slice < NS;
    float a1[GRAM_MAXP], a2[GRAM_MAXP];
#pragma unroll
    for (int p = 0; p < GRAM_MAXP; ++p) { a1[p] = 0.f; a2[p] = 0.f; }
    for (int i = threadIdx.x; i < 2 * VV; i += blockDim.x) acc[i] = 0.f;
    for (int g0 = blockIdx.x * GPI; g0 < groups; g0 += gridDim.x * GPI) {
        const int ng = min(GPI, groups - g0);
        __syncthreads();
        const float4* yp = reinterpret_cast<const float4*>(TF + (size_t)g0 * V * C);      // y = tanh(f) * sigmoid(g), recomputed
        const float4* zp = reinterpret_cast<const float4*>(SG + (size_t)g0 * V * C);
        const float4* gp = reinterpret_cast<const float4*>(G + (size_t)g0 * V * 2 * C);
        for (int i = threadIdx.x; i < ng * V * C / 4; i += blockDim.x) {
            float4 x = __ldg(yp + i); const float4 z = SG ? __ldg(zp + i) : make_float4(1.f, 1.f, 1.f, 1.f);
            x.x *= z.x; x.y *= z.y; x.z *= z.z; x.w *= z.w;
            int e = i * 4; int r = e / C, c = e - r * C;
            float* d = sy + r * ldy + c; d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
        }
        for (int i = threadIdx.x; i < ng * V * 2 * C / 4; i += blockDim.x) {
            float4 x = __ldg(gp + i);
            int e = i * 4; int r = e / (2 * C), c = e - r * 2 * C;
            float* d = sg + r * ldg + c; d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int p = 0; p < GRAM_MAXP; ++p) {
                int pair = p0 + p * 256;
                if (pair < VV && (p == 0 || VV >= 256)) {
                    int v = pair / V, w = pair - v * V;
                    float s1 = 0.f, s2 = 0.f;
                    for (int g = 0; g < ng; ++g) {
                        const float* yr = sy + (g * V + v) * ldy;
                        const float* gr = sg + (g * V + w) * ldg;
                        for (int c = slice; c < C; c += NS) { s1 = fmaf(yr[c], gr[c], s1); s2 = fmaf(yr[c], gr[C + c], s2); }
                    }
                    a1[p] += s1; a2[p] += s2;
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int p = 0; p < GRAM_MAXP; ++p) {
            int pair = p0 + p * 256;
            if (pair < VV && (p == 0 || VV >= 256)) { atomicAdd(acc + pair, a1[p]); atomicAdd(acc + VV + pair, a2[p]); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * VV; i += blockDim.x) atomicAdd(M12 + i, acc[i]);
}

// BatchNorm backward, elementwise part: du = gamma*rstd*(dxn - mean(dxn) - xhat*mean(dxn*xhat))
__global__ void bn_bwd_kernel(const float* __restrict__ dxn, const float* __restrict__ u, const float* __restrict__ mr,
                              const float* __restrict__ gamma, const double* __restrict__ sums, double count,
                              float* __restrict__ du, float* dgamma, float* dbeta, size_t rows, int C, int training)
{
    extern __shared__ float sm[];
    float* k1 = sm; float* m1 = k1 + C; float* m2 = m1 + C; float* mu = m2 + C; float* rs = mu + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float rstd = mr[C + c];
        k1[c] = gamma[c] * rstd; mu[c] = mr[c]; rs[c] = rstd;
        m1[c] = training ? (float)(sums[c] / count) : 0.f;
        m2[c] = training ? (float)(sums[C + c] / count) : 0.f;
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[c] = (float)sums[c];
            if (dgamma) dgamma[c] = (float)sums[C + c];
        }
    }
    __syncthreads();
    size_t n4 = rows * C / 4;
    const float4* d4 = reinterpret_cast<const float4*>(dxn);
    const float4* u4 = reinterpret_cast<const float4*>(u);
    float4* o4 = reinterpret_cast<float4*>(du);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        int c = (int)((i * 4) % C);
        float4 d = d4[i], x = u4[i], o;
        o.x = k1[c] * (d.x - m1[c] - (x.x - mu[c]) * rs[c] * m2[c]);
        o.y = k1[c + 1] * (d.y - m1[c + 1] - (x.y - mu[c + 1]) * rs[c + 1] * m2[c + 1]);
        o.z = k1[c + 2] * (d.z - m1[c + 2] - (x.z - mu[c + 2]) * rs[c + 2] * m2[c + 2]);
        o.w = k1[c + 3] * (d.w - m1[c + 3] - (x.w - mu[c + 3]) * rs[c + 3] * m2[c + 3]);
        o4[i] = o;
    }
}

// (B, O, V, T) contiguous <-> rows layout (B, T, V, O); tiny tensors (head in/out)
__global__ void nchw_to_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int O, int V, int T)
{
    size_t n = (size_t)B * O * V * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int o = (int)(i % O); size_t r = i / O; int v = (int)(r % V); size_t bt = r / V; int t = (int)(bt % T); size_t b = bt / T;
        dst[i] = src[((b * O + o) * V + v) * T + t];
    }
}
__global__ void rows_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int O, int V, int T)
{
    size_t n = (size_t)B * O * V * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(i % T); size_t r = i / T; int v = (int)(r % V); size_t bo = r / V; int o = (int)(bo % O); size_t b = bo / O;
        dst[i] = src[((b * T + t) * V + v) * O + o];
    }
}

// ============================================================================ gwnet-specific functors
struct LayerGeom {           // rows bookkeeping of one layer
    int V, C, Ti, To, d;
    __device__ __forceinline__ long in_row(int m_out) const {      // out row -> row of the layer input at the same (b,t,v)
        int per = To * V; int b = m_out / per;
        return (long)m_out + (long)b * d * V;
    }
};

struct StartA {              // A(m, k) = x[b, k, v, t - pad] (0 in the left padding), arbitrary strides
    static constexpr bool kFast = true;
    const float* x; int Tp, V, pad; long sB, sC, sV, sT;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int v = m % V; int bt = m / V; int t = bt % Tp - pad; int b = bt / Tp;
        if (t < 0) return 0.f;
        return __ldg(x + b * sB + k * sC + v * sV + t * sT);
    }
};
struct StartAT {             // B'(kout, m) for the weight gradient, ones column at kout == K
    static constexpr bool kFast = false;
    const float* x; int Tp, V, pad, K; long sB, sC, sV, sT;
    __device__ __forceinline__ float operator()(int kout, int m) const {
        if (kout == K) return 1.f;
        int v = m % V; int bt = m / V; int t = bt % Tp - pad; int b = bt / Tp;
        if (t < 0) return 0.f;
        return __ldg(x + b * sB + kout * sC + v * sV + t * sT);
    }
};
struct StartDxEpi {          // dIn rows layout over the un-padded T
    float* dx; int Tp, T, V, pad, K;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[8]) {
        int vv = m % V; int bt = m / V; int t = bt % Tp - pad; int b = bt / Tp;
        if (t < 0) return;
        float* row = dx + ((size_t)(b * T + t) * V + vv) * K;
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) { int n = nb + 64 * g + j; if (n < K) row[n] = v[4 * g + j]; }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        if (!valid) return;
        int vv = m % V; int bt = m / V; int t = bt % Tp - pad; int b = bt / Tp;
        if (t < 0) return;
        float* row = dx + ((size_t)(b * T + t) * V + vv) * K;
#pragma unroll
        for (int j = 0; j < 32; ++j) { int n = n0 + j; if (n < K) row[n] = v[j]; }
    }
    __device__ __forceinline__ void finish(float*) {}
};

struct GateA {               // k = tap*C + c ; value = BN_{i-1}(u_prev)[b, t + tap*d, v, c]
    static constexpr bool kFast = true;
    const float* up; const float* ss; LayerGeom g;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int tap = k >= g.C; int c = k - tap * g.C;
        long r = g.in_row(m) + (long)tap * g.d * g.V;
        return fmaf(__ldg(up + r * g.C + c), __ldg(ss + c), __ldg(ss + g.C + c));
    }
};
struct GateB {               // logical column n: block of 128 = [64 filter channels | the same 64 gate channels]
    static constexpr bool kFast = true;
    const float* wf; const float* wg; int C; int tcmap;     // tcmap: blocks of 32 = [16 filter | the same 16 gate]
    __device__ __forceinline__ float operator()(int n, int k) const {
        int fg, o;
        if (tcmap) { fg = (n >> 4) & 1; o = (n >> 5) * 16 + (n & 15); }
        else { int blk = n >> 7, w = n & 127; fg = w >> 6; o = blk * 64 + (w & 63); }
        if (o >= C) return 0.f;
        int tap = k >= C; int c = k - tap * C;
        return __ldg((fg ? wg : wf) + (size_t)o * 2 * C + 2 * c + tap);
    }
};
struct GateEpi {             // tanh(f)*sigmoid(g) (gwnet.py:186-200); keeps tf, sg, y; writes the skip slice
    const float* bf; const float* bg; float* TF; float* SG; float* Y; float* ycat;
    LayerGeom g; int layer, L, Tl;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[8]) {
        int c0 = (nb >> 7) * 64 + (nb & 127);
        if (c0 >= g.C) return;
        float tf[4], sg[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tf[j] = tanhf(v[j] + __ldg(bf + c0 + j));
            sg[j] = sigmoidf_acc(v[4 + j] + __ldg(bg + c0 + j));
            y[j] = tf[j] * sg[j];
        }
        size_t o = (size_t)m * g.C + c0;
        *reinterpret_cast<float4*>(TF + o) = make_float4(tf[0], tf[1], tf[2], tf[3]);
        *reinterpret_cast<float4*>(SG + o) = make_float4(sg[0], sg[1], sg[2], sg[3]);
        *reinterpret_cast<float4*>(Y + o) = make_float4(y[0], y[1], y[2], y[3]);
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.To; int b = bt / g.To;
        int tt = t - (g.To - Tl);
        if (tt >= 0) {
            size_t q = ((size_t)(b * Tl + tt) * g.V + vv) * ((size_t)L * g.C) + (size_t)layer * g.C + c0;
            *reinterpret_cast<float4*>(ycat + q) = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
    __device__ __forceinline__ void flush(int) {}
    // tensor-core column map: v[0..15] = filter channels c0.., v[16..31] = gate channels c0..
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        int c0 = (n0 >> 5) * 16;
        if (!valid || c0 >= g.C) return;
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.To; int b = bt / g.To;
        int tt = t - (g.To - Tl);
        size_t o = (size_t)m * g.C + c0;
        size_t q = tt >= 0 ? ((size_t)(b * Tl + tt) * g.V + vv) * ((size_t)L * g.C) + (size_t)layer * g.C + c0 : 0;
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
            if (c0 + j4 >= g.C) break;
            float tf[4], sg[4], y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tf[j] = tanhf(v[j4 + j] + __ldg(bf + c0 + j4 + j));
                sg[j] = sigmoidf_acc(v[16 + j4 + j] + __ldg(bg + c0 + j4 + j));
                y[j] = tf[j] * sg[j];
            }
            *reinterpret_cast<float4*>(TF + o + j4) = make_float4(tf[0], tf[1], tf[2], tf[3]);
            *reinterpret_cast<float4*>(SG + o + j4) = make_float4(sg[0], sg[1], sg[2], sg[3]);
            *reinterpret_cast<float4*>(Y + o + j4) = make_float4(y[0], y[1], y[2], y[3]);
            if (tt >= 0) *reinterpret_cast<float4*>(ycat + q + j4) = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

struct Seg3A {               // A(m, k): k = seg*C + c over three row-layout sources
    static constexpr bool kFast = true;
    const float* p0; const float* p1; const float* p2; int C;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int seg = k / C; int c = k - seg * C;
        const float* p = seg == 0 ? p0 : (seg == 1 ? p1 : p2);
        return __ldg(p + (size_t)m * C + c);
    }
    __device__ __forceinline__ void ld8(int m, int k, int kmax, float (&f)[8]) const {      // C % 8 == 0
        if (k >= kmax) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int seg = k / C; int c = k - seg * C;
        const float* p = (seg == 0 ? p0 : (seg == 1 ? p1 : p2)) + (size_t)m * C + c;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
struct Seg3AT {              // B'(kout, m): same sources, transposed role, ones column at 3C
    static constexpr bool kFast = false;
    const float* p0; const float* p1; const float* p2; int C;
    __device__ __forceinline__ float operator()(int kout, int m) const {
        if (kout == 3 * C) return 1.f;
        int seg = kout / C; int c = kout - seg * C;
        const float* p = seg == 0 ? p0 : (seg == 1 ? p1 : p2);
        return __ldg(p + (size_t)m * C + c);
    }
};
template <int NG>
struct MlpEpi {              // u = h + bias + residual (gwnet.py:43-45, 233) and BatchNorm statistics
    const float* bm; const float* up; const float* ss; float* U; double* stats; LayerGeom g; int tc_bn;
    float s1[4 * NG], s2[4 * NG];
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float* red) {
        float sq[32];
        long rr = valid ? g.in_row(m) + (long)g.d * g.V : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            float u = 0.f;
            if (valid && n < g.C) {
                float x = fmaf(__ldg(up + rr * g.C + n), __ldg(ss + n), __ldg(ss + g.C + n));
                u = v[j] + __ldg(bm + n) + x;
                U[(size_t)m * g.C + n] = u;
            }
            v[j] = u; sq[j] = u * u;
        }
        float a = warp_transpose_sum(v), b = warp_transpose_sum(sq);
        int slot = (int)(threadIdx.x >> 5) * 256 + ((n0 + (threadIdx.x & 31)) & 127);      // this warp's own slots: no atomics
        red[slot] = a; red[128 + slot] = b;
    }
    __device__ __forceinline__ void finish(float* red) {
        int n = blockIdx.x * tc_bn + threadIdx.x;
        if ((int)threadIdx.x < tc_bn && n < g.C && stats) {
            const int q = n & 127;
            atomicAdd(stats + n, (double)((red[q] + red[256 + q]) + (red[512 + q] + red[768 + q])));
            atomicAdd(stats + g.C + n, (double)((red[128 + q] + red[384 + q]) + (red[640 + q] + red[896 + q])));
        }
    }
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
        long rr = g.in_row(m) + (long)g.d * g.V;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * gg + j;
                if (n < g.C) {
                    float x = fmaf(__ldg(up + rr * g.C + n), __ldg(ss + n), __ldg(ss + g.C + n));
                    float u = v[4 * gg + j] + __ldg(bm + n) + x;
                    U[(size_t)m * g.C + n] = u;
                    s1[4 * gg + j] += u; s2[4 * gg + j] += u * u;
                }
            }
    }
    __device__ __forceinline__ void flush(int nb) {
#pragma unroll
        for (int q = 0; q < 4 * NG; ++q) {
            float a = s1[q] + __shfl_xor_sync(0xffffffffu, s1[q], 16);
            float b = s2[q] + __shfl_xor_sync(0xffffffffu, s2[q], 16);
            int n = nb + 64 * (q >> 2) + (q & 3);
            if ((threadIdx.x & 31) < 16 && n < g.C && stats) {
                atomicAdd(stats + n, (double)a);
                atomicAdd(stats + g.C + n, (double)b);
            }
        }
    }
};

struct SkipW {               // B(n, k): concatenated skip weights, k = layer*C + c  
    static constexpr bool kFast = true;
    const float* w[HOPK_MAX_LAYERS]; int C;
    __device__ __forceinline__ float operator()(int n, int k) const {
        int l = k / C; int c = k - l * C;
        return __ldg(w[l] + (size_t)n * C + c);
    }
    __device__ __forceinline__ void ld8(int n, int k, int kmax, float (&f)[8]) const {      // C % 8 == 0
        if (k >= kmax) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int l = k / C; int c = k - l * C;
        const float* p = w[l] + (size_t)n * C + c;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
struct SkipWT {              // B(n = layer*C + c, k = s) = w[layer][s][c]   (dgrad through the concat GEMM)
    static constexpr bool kFast = false;
    const float* w[HOPK_MAX_LAYERS]; int C;
    __device__ __forceinline__ float operator()(int n, int k) const {
        int l = n / C; int c = n - l * C;
        return __ldg(w[l] + (size_t)k * C + c);
    }
    __device__ __forceinline__ void ld8mn(int k, int n, int nmax, float (&f)[8]) const {    // C % 8 == 0
        if (n >= nmax) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int l = n / C; int c = n - l * C;
        const float* p = w[l] + (size_t)k * C + c;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
template <int NG>
struct SkipEpi {             // relu(sum + sum_l bias_l)
    float* out; const float* b[HOPK_MAX_LAYERS]; int L, N;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * g + j;
                if (n < N) {
                    float x = v[4 * g + j];
                    for (int l = 0; l < L; ++l) x += __ldg(b[l] + n);
                    out[(size_t)m * N + n] = fmaxf(x, 0.f);
                }
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        if (!valid) return;
        if (n0 + 32 <= N && (N & 3) == 0) {
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 bb = __ldg(reinterpret_cast<const float4*>(b[l] + n0) + q);
                    v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            float* o = out + (size_t)m * N + n0;
            if ((N & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) tc::stg256(o + j, v + j);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            if (n < N) {
                float x = v[j];
                for (int l = 0; l < L; ++l) x += __ldg(b[l] + n);
                out[(size_t)m * N + n] = fmaxf(x, 0.f);
            }
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};
template <int NG>
struct SkipWgradEpi {        // out row n = s, col k = layer*C + c (+ bias column at L*C, written to every layer's bias)
    float* dw[HOPK_MAX_LAYERS]; float* db[HOPK_MAX_LAYERS]; int C, L;
    __device__ __forceinline__ void operator()(int n, int kb, const float (&v)[4 * NG]) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int k = kb + 64 * g + j;
                if (k < L * C) { int l = k / C; atomicAdd(dw[l] + (size_t)n * C + (k - l * C), v[4 * g + j]); }
                else if (k == L * C) { for (int l = 0; l < L; ++l) atomicAdd(db[l] + n, v[4 * g + j]); }
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int n, bool valid, int k0, float (&v)[32], float*) {
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int k = k0 + j;
            if (k < L * C) { int l = k / C; atomicAdd(dw[l] + (size_t)n * C + (k - l * C), v[j]); }
            else if (k == L * C) { for (int l = 0; l < L; ++l) atomicAdd(db[l] + n, v[j]); }
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

struct DyB {                 // B(n = c, k = seg*C + o) = Wm[o][seg*C + c]
    static constexpr bool kFast = false;
    const float* wm; int C;
    __device__ __forceinline__ float operator()(int n, int k) const {
        int seg = k / C; int o = k - seg * C;
        return __ldg(wm + (size_t)o * 3 * C + seg * C + n);
    }
    __device__ __forceinline__ void ld8mn(int k, int n, int nmax, float (&f)[8]) const {    // C % 8 == 0
        if (n >= nmax) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int seg = k / C; int o = k - seg * C;
        const float* p = wm + (size_t)o * 3 * C + seg * C + n;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
template <int NG>
struct DyEpi {               // dy (+ skip-path gradient) -> df, dg   (Appendix A, gated TCN backward)
    const float* TF; const float* SG; const float* dycat; float* DF; float* DG; LayerGeom g; int layer, L, Tl;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.To; int b = bt / g.To;
        int tt = t - (g.To - Tl);
        const float* dyc = tt >= 0 ? dycat + ((size_t)(b * Tl + tt) * g.V + vv) * ((size_t)L * g.C) + (size_t)layer * g.C : nullptr;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * gg + j;
                if (n < g.C) {
                    float dy = v[4 * gg + j] + (dyc ? __ldg(dyc + n) : 0.f);
                    size_t o = (size_t)m * g.C + n;
                    float tf = __ldg(TF + o), sg = __ldg(SG + o);
                    DF[o] = dy * sg * (1.f - tf * tf);
                    DG[o] = dy * tf * sg * (1.f - sg);
                }
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        if (!valid) return;
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.To; int b = bt / g.To;
        int tt = t - (g.To - Tl);
        const float* dyc = tt >= 0 ? dycat + ((size_t)(b * Tl + tt) * g.V + vv) * ((size_t)L * g.C) + (size_t)layer * g.C : nullptr;
        if (n0 + 32 <= g.C && (g.C & 7) == 0) {              // 128-bit path, loads batched ahead of the stores
            const size_t o = (size_t)m * g.C + n0;
            float4 tf[8], sg[8], dc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                tf[q] = __ldg(reinterpret_cast<const float4*>(TF + o) + q);
                sg[q] = __ldg(reinterpret_cast<const float4*>(SG + o) + q);
                dc[q] = dyc ? __ldg(reinterpret_cast<const float4*>(dyc + n0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 dy = make_float4(v[4 * q] + dc[q].x, v[4 * q + 1] + dc[q].y, v[4 * q + 2] + dc[q].z, v[4 * q + 3] + dc[q].w);
                float4 df, dg;
                df.x = dy.x * sg[q].x * (1.f - tf[q].x * tf[q].x); dg.x = dy.x * tf[q].x * sg[q].x * (1.f - sg[q].x);
                df.y = dy.y * sg[q].y * (1.f - tf[q].y * tf[q].y); dg.y = dy.y * tf[q].y * sg[q].y * (1.f - sg[q].y);
                df.z = dy.z * sg[q].z * (1.f - tf[q].z * tf[q].z); dg.z = dy.z * tf[q].z * sg[q].z * (1.f - sg[q].z);
                df.w = dy.w * sg[q].w * (1.f - tf[q].w * tf[q].w); dg.w = dy.w * tf[q].w * sg[q].w * (1.f - sg[q].w);
                v[4 * q] = df.x; v[4 * q + 1] = df.y; v[4 * q + 2] = df.z; v[4 * q + 3] = df.w;
                reinterpret_cast<float*>(&tf[q])[0] = dg.x; reinterpret_cast<float*>(&tf[q])[1] = dg.y;
                reinterpret_cast<float*>(&tf[q])[2] = dg.z; reinterpret_cast<float*>(&tf[q])[3] = dg.w;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {                    // full 32-byte sectors
                tc::stg256(DF + o + j, v + j);
                tc::stg256(DG + o + j, tf[j / 4].x, tf[j / 4].y, tf[j / 4].z, tf[j / 4].w, tf[j / 4 + 1].x, tf[j / 4 + 1].y, tf[j / 4 + 1].z, tf[j / 4 + 1].w);
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            if (n < g.C) {
                float dy = v[j] + (dyc ? __ldg(dyc + n) : 0.f);
                size_t o = (size_t)m * g.C + n;
                float tf = __ldg(TF + o), sg = __ldg(SG + o);
                DF[o] = dy * sg * (1.f - tf * tf);
                DG[o] = dy * tf * sg * (1.f - sg);
            }
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

struct GateWgA {             // A'(n = fg*C + o, m) = (fg ? DG : DF)[m][o]
    static constexpr bool kFast = false;
    const float* DF; const float* DG; int C;
    __device__ __forceinline__ float operator()(int n, int m) const {
        int fg = n >= C; int o = n - fg * C;
        return __ldg((fg ? DG : DF) + (size_t)m * C + o);
    }
};
struct GateWgB {             // B'(kout = 2c + tap, m) = x[in_row(m) + tap*d*V][c] ; ones column at 2C
    static constexpr bool kFast = false;
    const float* up; const float* ss; LayerGeom g;
    __device__ __forceinline__ float operator()(int kout, int m) const {
        if (kout == 2 * g.C) return 1.f;
        int c = kout >> 1, tap = kout & 1;
        long r = g.in_row(m) + (long)tap * g.d * g.V;
        return fmaf(__ldg(up + r * g.C + c), __ldg(ss + c), __ldg(ss + g.C + c));
    }
};
template <int NG>
struct GateWgEpi {
    float* dwf; float* dwg; float* dbf; float* dbg; int C;
    __device__ __forceinline__ void operator()(int n, int kb, const float (&v)[4 * NG]) {
        int fg = n >= C; int o = n - fg * C;
        float* dw = fg ? dwg : dwf; float* db = fg ? dbg : dbf;
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int k = kb + 64 * g + j;
                if (k < 2 * C) atomicAdd(dw + (size_t)o * 2 * C + k, v[4 * g + j]);
                else if (k == 2 * C) atomicAdd(db + o, v[4 * g + j]);
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int n, bool valid, int k0, float (&v)[32], float*) {
        if (!valid) return;
        int fg = n >= C; int o = n - fg * C;
        float* dw = fg ? dwg : dwf; float* db = fg ? dbg : dbf;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int k = k0 + j;
            if (k < 2 * C) atomicAdd(dw + (size_t)o * 2 * C + k, v[j]);
            else if (k == 2 * C) atomicAdd(db + o, v[j]);
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

// ---- operands / epilogue of the MN-major weight-gradient skeleton (gemm_tc_wgrad.cuh) ----
struct W8Seg3 {              // [p0 | p1 | p2], C columns each (C % 8 == 0)
    const float* p0; const float* p1; const float* p2; int C; int ncols;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        if (c0 >= ncols) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int seg = c0 / C; int c = c0 - seg * C;
        const float* p = (seg == 0 ? p0 : (seg == 1 ? p1 : p2)) + (size_t)r * C + c;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
struct W8GateX {             // columns j = tap*C + c : BN_{i-1}-folded layer input at (b, t + tap*d, v)
    const float* up; const float* ss; LayerGeom g; int ncols;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        if (c0 >= ncols) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int tap = c0 >= g.C; int c = c0 - tap * g.C;
        const float* p = up + (size_t)(g.in_row(r) + (long)tap * g.d * g.V) * g.C + c;
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        float4 s0 = __ldg(reinterpret_cast<const float4*>(ss + c)), s1 = __ldg(reinterpret_cast<const float4*>(ss + c + 4));
        float4 h0 = __ldg(reinterpret_cast<const float4*>(ss + g.C + c)), h1 = __ldg(reinterpret_cast<const float4*>(ss + g.C + c + 4));
        f[0] = fmaf(a.x, s0.x, h0.x); f[1] = fmaf(a.y, s0.y, h0.y); f[2] = fmaf(a.z, s0.z, h0.z); f[3] = fmaf(a.w, s0.w, h0.w);
        f[4] = fmaf(b.x, s1.x, h1.x); f[5] = fmaf(b.y, s1.y, h1.y); f[6] = fmaf(b.z, s1.z, h1.z); f[7] = fmaf(b.w, s1.w, h1.w);
    }
};
struct W8Prod {              // y = tanh(f) * sigmoid(g) recomputed from the two saved factors (C % 8 == 0)
    const float* p; const float* q; int C; int ncols;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        if (c0 >= ncols) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = 0.f;
            return;
        }
        float g[8];
        tc::ldg256(p + (size_t)r * C + c0, f);
        tc::ldg256(q + (size_t)r * C + c0, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] *= g[k];
    }
};
struct MlpWgEpi2 {           // out[i = seg*C + o][j = c] = sum_r [du|p1|p2][r][i] y[r][c]  ->  dWm[o][seg*C + c]; bias = column sums of du
    float* dw; float* db; int C;
    __device__ __forceinline__ void row32(int i, bool valid, int j0, float (&v)[32], float*) {
        if (!valid) return;
        const int seg = i / C, o = i - seg * C;
        float* row = dw + (size_t)o * 3 * C + seg * C + j0;
        if (j0 + 32 <= C && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                atomicAdd(reinterpret_cast<float4*>(row) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            return;
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
            if (j0 + jj < C) atomicAdd(row + jj, v[jj]);
    }
    __device__ __forceinline__ void finish(float*) {}
    __device__ __forceinline__ void bias(int i, float v) { if (i < C && db) atomicAdd(db + i, v); }
};
struct W8GateXI {            // columns j = 2c + tap (the weight's own [c][tap] order): BN-folded layer input at (b, t + tap*d, v)
    const float* up; const float* ss; LayerGeom g; int ncols;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        if (c0 >= ncols) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        const int c = c0 >> 1;                                    // 4 channels x 2 taps
        const float* p0 = up + (size_t)g.in_row(r) * g.C + c;
        const float* p1 = p0 + (size_t)g.d * g.V * g.C;
        float4 a = __ldg(reinterpret_cast<const float4*>(p0)), b = __ldg(reinterpret_cast<const float4*>(p1));
        float4 sc = __ldg(reinterpret_cast<const float4*>(ss + c)), sh = __ldg(reinterpret_cast<const float4*>(ss + g.C + c));
        f[0] = fmaf(a.x, sc.x, sh.x); f[1] = fmaf(b.x, sc.x, sh.x); f[2] = fmaf(a.y, sc.y, sh.y); f[3] = fmaf(b.y, sc.y, sh.y);
        f[4] = fmaf(a.z, sc.z, sh.z); f[5] = fmaf(b.z, sc.z, sh.z); f[6] = fmaf(a.w, sc.w, sh.w); f[7] = fmaf(b.w, sc.w, sh.w);
    }
};
struct GateWgEpi2 {          // out[i = fg*C + o][j = 2c + tap] -> d(filter|gate)_w[o][c][tap]; bias = column sums of [DF|DG]
    float* dwf; float* dwg; float* dbf; float* dbg; int C;
    __device__ __forceinline__ void row32(int i, bool valid, int j0, float (&v)[32], float*) {
        if (!valid) return;
        int fg = i >= C; int o = i - fg * C;
        float* dw = (fg ? dwg : dwf) + (size_t)o * 2 * C + j0;
        if (j0 + 32 <= 2 * C && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                atomicAdd(reinterpret_cast<float4*>(dw) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            return;
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
            if (j0 + jj < 2 * C) atomicAdd(dw + jj, v[jj]);
    }
    __device__ __forceinline__ void finish(float*) {}
    __device__ __forceinline__ void bias(int i, float v) { int fg = i >= C; atomicAdd((fg ? dbg : dbf) + (i - fg * C), v); }
};

// ---- dA Gram products as ONE tensor-core weight-gradient GEMM (used for large graphs, V*V >= 256):
//   M1[v][w] = sum_{g,c} y[(g,v)][c] G[(g,w)][c],  M2[v][w] = sum_{g,c} y[(g,v)][c] G[(g,w)][C + c]
// is out[i][j] = sum_r A(r, i) B(r, j) with contraction rows r = (group g, channel c), i = v and j = w | V + w.
struct W8GramY {             // A(r = g*C + c, i = v) = tanh f * sigmoid g at row (g, v), channel c
    const float* tf; const float* sg; int V, C;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        const int g = r / C, c = r - g * C;
        const size_t base = (size_t)g * V * C + c;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int v = c0 + q;
            f[q] = v < V ? __ldg(tf + base + (size_t)v * C) * __ldg(sg + base + (size_t)v * C) : 0.f;
        }
    }
};
struct W8GramG {             // B(r = g*C + c, j) = G[(g, j)][c] for j < V, G[(g, j - V)][C + c] for V <= j < 2V   (G has ld 2C)
    const float* G; int V, C;
    __device__ __forceinline__ void ld8(int r, int c0, float (&f)[8]) const {
        const int g = r / C, c = r - g * C;
        const size_t base = (size_t)g * V * 2 * C + c;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = c0 + q;
            float x = 0.f;
            if (j < V) x = __ldg(G + base + (size_t)j * 2 * C);
            else if (j < 2 * V) x = __ldg(G + base + (size_t)(j - V) * 2 * C + C);
            f[q] = x;
        }
    }
};
struct GramEpi {             // out[v][j] -> M12[v*V + w] (j = w < V) or M12[V*V + v*V + w] (j = V + w)
    float* M12; int V;
    __device__ __forceinline__ void row32(int i, bool valid, int j0, float (&v)[32], float*) {
        if (!valid || i >= V) return;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = j0 + jj;
            if (j < V) atomicAdd(M12 + i * V + j, v[jj]);
            else if (j < 2 * V) atomicAdd(M12 + V * V + i * V + (j - V), v[jj]);
        }
    }
    __device__ __forceinline__ void finish(float*) {}
    __device__ __forceinline__ void bias(int, float) {}
};

struct DxA {                 // A(m_in, k = (tap*2 + fg)*C + o) = (fg?DG:DF)[(b, t - tap*d, v)][o], 0 outside [0, To)
    static constexpr bool kFast = true;
    const float* DF; const float* DG; LayerGeom g;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int q = k / g.C; int o = k - q * g.C; int tap = q >> 1, fg = q & 1;
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.Ti - tap * g.d; int b = bt / g.Ti;
        if (t < 0 || t >= g.To) return 0.f;
        return __ldg((fg ? DG : DF) + ((size_t)(b * g.To + t) * g.V + vv) * g.C + o);
    }
    __device__ __forceinline__ void ld8(int m, int k, int kmax, float (&f)[8]) const {      // C % 8 == 0
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = 0.f;
        if (k >= kmax) return;
        int q4 = k / g.C; int o = k - q4 * g.C; int tap = q4 >> 1, fg = q4 & 1;
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.Ti - tap * g.d; int b = bt / g.Ti;
        if (t < 0 || t >= g.To) return;
        const float* p = (fg ? DG : DF) + ((size_t)(b * g.To + t) * g.V + vv) * g.C + o;
        float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p + 4));
        f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w; f[4] = y.x; f[5] = y.y; f[6] = y.z; f[7] = y.w;
    }
};
struct DxB {                 // B(n = c, k) = W_fg[o][c][tap]
    static constexpr bool kFast = false;
    const float* wf; const float* wg; int C;
    __device__ __forceinline__ float operator()(int n, int k) const {
        int q = k / C; int o = k - q * C; int tap = q >> 1, fg = q & 1;
        return __ldg((fg ? wg : wf) + (size_t)o * 2 * C + 2 * n + tap);
    }
    __device__ __forceinline__ void ld8mn(int k, int n, int nmax, float (&f)[8]) const {    // C % 8 == 0
        if (n >= nmax) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = 0.f;
            return;
        }
        int q4 = k / C; int o = k - q4 * C; int tap = q4 >> 1, fg = q4 & 1;
        const float4* p = reinterpret_cast<const float4*>((fg ? wg : wf) + (size_t)o * 2 * C + 2 * n);
        float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
        if (tap) { f[0] = a.y; f[1] = a.w; f[2] = b.y; f[3] = b.w; f[4] = c.y; f[5] = c.w; f[6] = d.y; f[7] = d.w; }
        else     { f[0] = a.x; f[1] = a.z; f[2] = b.x; f[3] = b.z; f[4] = c.x; f[5] = c.z; f[6] = d.x; f[7] = d.z; }
    }
};
template <int NG>
struct DxEpi {               // + residual gradient du[t-d]; BatchNorm-backward sums for the previous layer
    const float* DU; float* DX; const float* uprev; const float* mr_prev; double* sums_prev; LayerGeom g; int tc_bn;
    float s1[4 * NG], s2[4 * NG];
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float* red) {
        float w[32];
        int vv = 0, t = 0, b = 0;
        if (valid) { vv = m % g.V; int bt = m / g.V; t = bt % g.Ti; b = bt / g.Ti; }
        const float* du = (valid && DU && t >= g.d) ? DU + ((size_t)(b * g.To + t - g.d) * g.V + vv) * g.C : nullptr;
        if (n0 + 32 <= g.C && (g.C & 7) == 0) {              // 128-bit path
            float4 dd[8], uu[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                dd[q] = du ? __ldg(reinterpret_cast<const float4*>(du + n0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                uu[q] = (valid && sums_prev) ? __ldg(reinterpret_cast<const float4*>(uprev + (size_t)m * g.C + n0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
                if (sums_prev) {
                    mu = __ldg(reinterpret_cast<const float4*>(mr_prev + n0) + q);
                    rs = __ldg(reinterpret_cast<const float4*>(mr_prev + g.C + n0) + q);
                }
                float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    dx = make_float4(v[4 * q] + dd[q].x, v[4 * q + 1] + dd[q].y, v[4 * q + 2] + dd[q].z, v[4 * q + 3] + dd[q].w);
                }
                v[4 * q] = dx.x; v[4 * q + 1] = dx.y; v[4 * q + 2] = dx.z; v[4 * q + 3] = dx.w;
                w[4 * q] = dx.x * (uu[q].x - mu.x) * rs.x; w[4 * q + 1] = dx.y * (uu[q].y - mu.y) * rs.y;
                w[4 * q + 2] = dx.z * (uu[q].z - mu.z) * rs.z; w[4 * q + 3] = dx.w * (uu[q].w - mu.w) * rs.w;
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) tc::stg256(DX + (size_t)m * g.C + n0 + j, v + j);      // full 32-byte sectors
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int n = n0 + j;
                float dx = 0.f, xh = 0.f;
                if (valid && n < g.C) {
                    dx = v[j] + (du ? __ldg(du + n) : 0.f);
                    DX[(size_t)m * g.C + n] = dx;
                    if (sums_prev) xh = (__ldg(uprev + (size_t)m * g.C + n) - __ldg(mr_prev + n)) * __ldg(mr_prev + g.C + n);
                }
                v[j] = dx; w[j] = dx * xh;
            }
        }
        if (sums_prev) {
            float a = warp_transpose_sum(v), c = warp_transpose_sum(w);
            int slot = (int)(threadIdx.x >> 5) * 256 + ((n0 + (threadIdx.x & 31)) & 127);  // this warp's own slots: no atomics
            red[slot] = a; red[128 + slot] = c;
        }
    }
    __device__ __forceinline__ void finish(float* red) {
        int n = blockIdx.x * tc_bn + threadIdx.x;
        if (sums_prev && (int)threadIdx.x < tc_bn && n < g.C) {
            const int q = n & 127;
            atomicAdd(sums_prev + n, (double)((red[q] + red[256 + q]) + (red[512 + q] + red[768 + q])));
            atomicAdd(sums_prev + g.C + n, (double)((red[128 + q] + red[384 + q]) + (red[640 + q] + red[896 + q])));
        }
    }
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
        int vv = m % g.V; int bt = m / g.V; int t = bt % g.Ti; int b = bt / g.Ti;
        const float* du = (DU && t >= g.d) ? DU + ((size_t)(b * g.To + t - g.d) * g.V + vv) * g.C : nullptr;
#pragma unroll
        for (int gg = 0; gg < NG; ++gg)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * gg + j;
                if (n < g.C) {
                    float dx = v[4 * gg + j] + (du ? __ldg(du + n) : 0.f);
                    DX[(size_t)m * g.C + n] = dx;
                    if (sums_prev) {
                        float xh = (__ldg(uprev + (size_t)m * g.C + n) - __ldg(mr_prev + n)) * __ldg(mr_prev + g.C + n);
                        s1[4 * gg + j] += dx; s2[4 * gg + j] += dx * xh;
                    }
                }
            }
    }
    __device__ __forceinline__ void flush(int nb) {
        if (!sums_prev) return;
#pragma unroll
        for (int q = 0; q < 4 * NG; ++q) {
            float a = s1[q] + __shfl_xor_sync(0xffffffffu, s1[q], 16);
            float b = s2[q] + __shfl_xor_sync(0xffffffffu, s2[q], 16);
            int n = nb + 64 * (q >> 2) + (q & 3);
            if ((threadIdx.x & 31) < 16 && n < g.C) {
                atomicAdd(sums_prev + n, (double)a);
                atomicAdd(sums_prev + g.C + n, (double)b);
            }
        }
    }
};

}  // namespace hopk
#include "gwnet_fused.cuh"
namespace hopk {

// ============================================================================ launch helpers
template <int MG, int NG, class AL, class BL, class EP>
static void launch_gemm(bool tc, int M, int N, int K, int splits, AL a, BL b, EP e, cudaStream_t st)
{
    if (tc) {                     // dtype 1: bf16 operands on tcgen05, fp32 accumulate (gemm_tc.cuh)
        if (N <= 64) launch_gemm_tc<64>(M, N, K, splits, a, b, e, st);
        else launch_gemm_tc<128>(M, N, K, splits, a, b, e, st);
        return;
    }
    int kper = K;
    if (splits > 1) { kper = ((cdiv(K, splits) + GEMM_BK - 1) / GEMM_BK) * GEMM_BK; splits = cdiv(K, kper); }
    if (splits < 1) splits = 1;
    dim3 grid(cdiv(N, 64 * NG), cdiv(M, 64 * MG), splits);
    gemm_kernel<MG, NG, AL, BL, EP><<<grid, GEMM_THREADS, 0, st>>>(M, N, K, kper, a, b, e);
}

static int pick_splits(bool tc, int M, int N, int K, int mg, int ng)
{
    if (tc) { mg = 2; ng = N <= 64 ? 1 : 2; }
    long tiles = (long)cdiv(M, 64 * mg) * cdiv(N, 64 * ng);
    long want = (2 * 148 + tiles - 1) / tiles;
    long maxs = K / (tc ? 512 : 256) > 0 ? K / (tc ? 512 : 256) : 1;
    return (int)(want < maxs ? want : maxs);
}

static size_t node_mix_smem(int V, int C, int gpb) { return (2 * (size_t)((V * V + 3) & ~3) + (size_t)gpb * V * C) * sizeof(float); }

static int launch_node_mix(const float* in, const float* M1, const float* M2, float* o1, float* o2, int groups, int V, int C,
                           cudaStream_t st)
{
    int gpb = 48 / V > 0 ? 48 / V : 1;
    size_t smem = node_mix_smem(V, C, gpb);
    if (smem > 48 * 1024) HOPK_CUDA(configure_smem_once((const void*)node_mix_kernel, 200 * 1024));
    node_mix_kernel<<<cdiv(groups, gpb), 256, smem, st>>>(in, M1, M2, o1, o2, groups, V, C, gpb);
    HOPK_LAUNCH_CHECK("node_mix");
    return 0;
}

}  // namespace hopk

using namespace hopk;

// ============================================================================ C ABI
// ============================================================================ backward plumbing
struct SideStreams {
    cudaStream_t s[3];
    cudaEvent_t ev_du[HOPK_MAX_LAYERS], ev_dfg[HOPK_MAX_LAYERS], ev_join[3], ev_head[3], ev_tail;
};
// created once per device (keyed by the current device, so a process that drives several GPUs gets streams and events
// on the right one); non-blocking so they never serialise against stream 0.  Calls on one device are expected from one
// host thread at a time (the autograd thread of that device's rank): events are reused from call to call.
// Joins the side streams into `st` when a backward entry point leaves early with an error, so that a caller that is
// capturing a CUDA graph can still end the capture cleanly (errors of the join itself are ignored: the call already failed).
struct SideJoinOnError {
    SideStreams* sd; cudaStream_t st; bool armed;
    ~SideJoinOnError()
    {
        if (!armed) return;
        for (int k = 0; k < 3; ++k)
            if (cudaEventRecord(sd->ev_join[k], sd->s[k]) == cudaSuccess) cudaStreamWaitEvent(st, sd->ev_join[k], 0);
        cudaGetLastError();
    }
};

static SideStreams* side_streams()
{
    constexpr int MAXDEV = 64;
    static SideStreams table[MAXDEV];
    static int states[MAXDEV] = {0};            // 0 = not created, 1 = ok, -1 = failed
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SideStreams& sd = table[dev];
    int& state = states[dev];
    if (state == 0) {
        state = 1;
        for (int k = 0; k < 3; ++k) {
            if (cudaStreamCreateWithFlags(&sd.s[k], cudaStreamNonBlocking) != cudaSuccess) state = -1;
            if (cudaEventCreateWithFlags(&sd.ev_join[k], cudaEventDisableTiming) != cudaSuccess) state = -1;
            if (cudaEventCreateWithFlags(&sd.ev_head[k], cudaEventDisableTiming) != cudaSuccess) state = -1;
        }
        if (cudaEventCreateWithFlags(&sd.ev_tail, cudaEventDisableTiming) != cudaSuccess) state = -1;
        for (int l = 0; l < HOPK_MAX_LAYERS; ++l) {
            if (cudaEventCreateWithFlags(&sd.ev_du[l], cudaEventDisableTiming) != cudaSuccess) state = -1;
            if (cudaEventCreateWithFlags(&sd.ev_dfg[l], cudaEventDisableTiming) != cudaSuccess) state = -1;
        }
    }
    return state == 1 ? &sd : nullptr;
}

static int zero_param_grads(const HopkGwnetShape* s, const HopkGwnetGrads* gr, cudaStream_t st)
{
    if (gr->flat && gr->flat_bytes) {
        HOPK_CUDA(cudaMemsetAsync(gr->flat, 0, gr->flat_bytes, st));
        return 0;
    }
    const size_t f = sizeof(float);
    const size_t C = s->C, Sk = s->S, E = s->E, O = s->out_dim;
    auto z = [&](float* p, size_t n) -> cudaError_t { return p ? cudaMemsetAsync(p, 0, n * f, st) : cudaSuccess; };
    HOPK_CUDA(z(gr->start_w, C * s->in_dim)); HOPK_CUDA(z(gr->start_b, C));
    HOPK_CUDA(z(gr->end1_w, E * Sk)); HOPK_CUDA(z(gr->end1_b, E));
    HOPK_CUDA(z(gr->end2_w, O * E)); HOPK_CUDA(z(gr->end2_b, O));
    for (int l = 0; l < s->L; ++l) {
        HOPK_CUDA(z(gr->filter_w[l], C * 2 * C)); HOPK_CUDA(z(gr->filter_b[l], C));
        HOPK_CUDA(z(gr->gate_w[l], C * 2 * C)); HOPK_CUDA(z(gr->gate_b[l], C));
        HOPK_CUDA(z(gr->skip_w[l], Sk * C)); HOPK_CUDA(z(gr->skip_b[l], Sk));
        HOPK_CUDA(z(gr->mlp_w[l], C * 3 * C)); HOPK_CUDA(z(gr->mlp_b[l], C));
    }
    return 0;
}

extern "C" size_t hopk_gwnet_workspace_bytes(const HopkGwnetShape* s) { return make_layout(s).total; }
extern "C" size_t hopk_gwnet_scratch_bytes(const HopkGwnetShape* s) { return make_layout(s).s_total; }
extern "C" int hopk_gwnet_out_steps(const HopkGwnetShape* s) { return make_layout(s).Tl; }

extern "C" int hopk_gwnet_ws_field(const HopkGwnetShape* s, const char* name, int layer, size_t* offset, size_t* bytes,
                                   int* elem_bytes)
{
    HOPK_REQUIRE(s && name && offset && bytes && elem_bytes, "null argument");
    HOPK_REQUIRE(s->L >= 1 && s->L <= HOPK_MAX_LAYERS && layer >= 0 && layer < s->L, "layer index");
    GwLayout g = make_layout(s);
    const size_t f = sizeof(float), BV = (size_t)s->B * s->V;
    const size_t nl = BV * g.Tlen[layer + 1] * s->C * f;
    *elem_bytes = 4;
    if (!strcmp(name, "x0")) { *offset = g.x0; *bytes = BV * g.Tp * s->C * f; }
    else if (!strcmp(name, "u")) { *offset = g.u[layer]; *bytes = nl; }
    else if (!strcmp(name, "tf")) { *offset = g.tf[layer]; *bytes = nl; }
    else if (!strcmp(name, "sg")) { *offset = g.sg[layer]; *bytes = nl; }
    else if (!strcmp(name, "ycat")) { *offset = g.ycat; *bytes = BV * g.Tl * s->L * s->C * f; }
    else if (!strcmp(name, "r0")) { *offset = g.r0; *bytes = BV * g.Tl * s->S * f; }
    else if (!strcmp(name, "r1")) { *offset = g.r1; *bytes = BV * g.Tl * s->E * f; }
    else if (!strcmp(name, "mr")) { *offset = g.mr + (size_t)layer * 2 * s->C * f; *bytes = 2 * s->C * f; }
    else if (!strcmp(name, "A")) { *offset = g.A; *bytes = (size_t)s->V * s->V * f; }
    else return fail(2, "bad argument: unknown workspace field", name);
    return 0;
}

static int check_shape(const HopkGwnetShape* s)
{
    HOPK_REQUIRE(s->bn_momentum >= 0.f && s->bn_momentum <= 1.f && s->bn_eps > 0.f, "BatchNorm momentum in [0,1] and eps > 0");
    HOPK_REQUIRE(s->dtype == 0 || s->dtype == 1, "gwnet: dtype must be 0 (fp32 FFMA) or 1 (bf16 tensor-core math)");
    HOPK_REQUIRE(s->L >= 1 && s->L <= HOPK_MAX_LAYERS, "layer count");
    HOPK_REQUIRE(s->C % 4 == 0 && s->C >= 4 && s->C <= 256, "C must be a multiple of 4, <= 256");
    HOPK_REQUIRE(s->V >= 1 && s->V <= 45, "V must be <= 45");
    HOPK_REQUIRE(s->B >= 1 && s->T >= 1 && s->in_dim >= 1 && s->out_dim >= 1 && s->S >= 1 && s->E >= 1, "sizes");
    HOPK_REQUIRE((long)s->B * s->V * 64 * 256 < (1L << 31), "tensor too large for 32-bit row indices");
    return 0;
}

extern "C" int hopk_gwnet_forward(const HopkGwnetShape* s, const HopkGwnetParams* p, const float* x, const int64_t xs[4],
                                  float* out, void* ws_, void* stream)
{
    if (int rc = check_shape(s)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GwLayout g = make_layout(s);
    char* ws = (char*)ws_;
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    const int C = s->C, V = s->V, B = s->B, L = s->L;
    const bool tc = s->dtype == 1;
    const int tcbn = C <= 64 ? 64 : 128;
    double* stats = reinterpret_cast<double*>(ws + g.stats);
    HOPK_CUDA(cudaMemsetAsync(stats, 0, (size_t)L * 2 * C * sizeof(double), st));

    adp_fwd_kernel<<<1, 256, V * V * sizeof(float), st>>>(p->nodevec1, p->nodevec2, V, s->rank, F(g.A), F(g.A2), F(g.At),
                                                           F(g.A2t), F(g.Z));
    HOPK_LAUNCH_CHECK("adp_fwd");
    fill_identity_kernel<<<cdiv(C, 128), 128, 0, st>>>(F(g.ss), C);

    // start conv (gwnet.py:144-149)
    {
        int M = B * g.Tp * V;
        StartA a{x, g.Tp, V, g.pad, (long)xs[0], (long)xs[1], (long)xs[2], (long)xs[3]};
        Ld2D<true, 0> b{p->start_w, nullptr, s->in_dim};
        EpiStore<1> e{F(g.x0), C, p->start_b, nullptr, C, 0};
        if (C <= 64) launch_gemm<2, 1>(tc, M, C, s->in_dim, 1, a, b, e, st);
        else {
            EpiStore<2> e2{F(g.x0), C, p->start_b, nullptr, C, 0};
            launch_gemm<2, 2>(tc, M, C, s->in_dim, 1, a, b, e2, st);
        }
        HOPK_LAUNCH_CHECK("start_conv");
    }

    const float* uprev = F(g.x0);
    const bool fused = tc && C == FZ_C && V <= 64;                   // one kernel per layer (gwnet_fused.cuh)
    uint8_t* pack = reinterpret_cast<uint8_t*>(((uintptr_t)(ws + g.pack) + 1023) & ~uintptr_t(1023));
    uint8_t* bd = pack + (size_t)L * FZ_PACK_BYTES;
    unsigned int* tickets = reinterpret_cast<unsigned int*>(ws + g.ticket);
    if (fused) {
        HOPK_CUDA(configure_smem_once((const void*)fz_layer_fwd_kernel, fz_smem_bytes()));
        HOPK_CUDA(cudaMemsetAsync(tickets, 0, HOPK_MAX_LAYERS * sizeof(unsigned int), st));
        fz_pack_weights_kernel<<<L, 256, 0, st>>>(*p, L, pack);
        HOPK_LAUNCH_CHECK("fz_pack");
        fz_pack_bd_kernel<<<16, 256, 0, st>>>(F(g.A), V, bd);
        HOPK_LAUNCH_CHECK("fz_pack_bd");
    }
    // The whole stack as one persistent cooperative kernel when the problem is latency-bound: every tile of the largest layer
    // has its own resident CTA (TED at batch 128: 146 tiles).  Larger problems (Expressive, batch 1024) keep one launch per
    // layer: with several tiles per CTA the hardware block scheduler overlaps them better than the in-kernel tile loop
    // (measured: 0.68 vs 0.76 ms at V = 42, 1.05 vs 1.11 ms at B = 1024).
    bool net = false;
    FzNetArgs na;
    int net_grid = 0;
    if (fused && L <= FZ_NET_LAYERS) {
        int dev = 0, sms = 0, per_sm = 0, coop = 0;
        HOPK_CUDA(cudaGetDevice(&dev));
        HOPK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        HOPK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        HOPK_CUDA(configure_smem_once((const void*)fz_net_fwd_kernel, fz_smem_bytes()));
        HOPK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fz_net_fwd_kernel, 256, fz_smem_bytes()));
        net_grid = sms * per_sm;
        net = coop && per_sm > 0 && cdiv(B * g.Tlen[1], 128 / V) <= net_grid;      // layer 0 has the most tiles
        na.L = L;
    }
    int max_tiles = 0;
    for (int i = 0; i < L; ++i) {
        LayerGeom lg{V, C, g.Tlen[i], g.Tlen[i + 1], s->dil[i]};
        int M = B * lg.To * V;
        const float* ss = F(g.ss) + (size_t)i * 2 * C;
        if (fused) {
            FzArgs fa;
            fa.up = uprev; fa.ss = ss; fa.pack = pack + (size_t)i * FZ_PACK_BYTES; fa.bd = bd;
            fa.bf = p->filter_b[i]; fa.bg = p->gate_b[i]; fa.bm = p->mlp_b[i];
            fa.TF = F(g.tf[i]); fa.SG = F(g.sg[i]); fa.U = F(g.u[i]);
            fa.Y = nullptr; fa.X1 = nullptr; fa.X2 = nullptr;        // backward recomputes y = tf * sg and never needs x1, x2
            fa.ycat = F(g.ycat); fa.stats = stats + (size_t)i * 2 * C;
            fa.g = lg; fa.layer = i; fa.L = L; fa.Tl = g.Tl; fa.groups = B * lg.To; fa.gpt = 128 / V;
            fa.ticket = tickets + i; fa.count = (double)M; fa.gamma = p->bn_w[i]; fa.beta = p->bn_b[i];
            fa.rmean = p->bn_mean[i]; fa.rvar = p->bn_var[i]; fa.nbt = (long long*)p->bn_nbt[i];
            fa.mr = F(g.mr) + (size_t)i * 2 * C; fa.ss_next = F(g.ss) + (size_t)(i + 1) * 2 * C; fa.training = s->training;
            fa.bn_momentum = s->bn_momentum; fa.bn_eps = s->bn_eps;
#ifdef HOPK_DEBUG
            { static const char* e = getenv("HOPK_FZ_STOP"); fa.stop = e ? atoi(e) : 0; }
#endif
            if (net) {
                na.layer[i] = fa;
                max_tiles = max_tiles > cdiv(fa.groups, fa.gpt) ? max_tiles : cdiv(fa.groups, fa.gpt);
            } else {
                fz_layer_fwd_kernel<<<cdiv(fa.groups, fa.gpt), 256, fz_smem_bytes(), st>>>(fa);
                HOPK_LAUNCH_CHECK("fz_layer_fwd");
            }
            uprev = F(g.u[i]);
            continue;
        }
        // gated dilated conv (gwnet.py:186-200) + skip slice
        {
            GateA a{uprev, ss, lg};
            GateB b{p->filter_w[i], p->gate_w[i], C, tc ? 1 : 0};
            GateEpi e{p->filter_b[i], p->gate_b[i], F(g.tf[i]), F(g.sg[i]), F(g.y[i]), F(g.ycat), lg, i, L, g.Tl};
            int Nlog = tc ? cdiv(C, 16) * 32 : cdiv(C, 64) * 128;
            launch_gemm<2, 2>(tc, M, Nlog, 2 * C, 1, a, b, e, st);
            HOPK_LAUNCH_CHECK("gate");
        }
        // diffusion (gwnet.py:12-14, 35-41)
        if (int rc = launch_node_mix(F(g.y[i]), F(g.A), F(g.A2), F(g.x1[i]), F(g.x2[i]), B * lg.To, V, C, st)) return rc;
        // gcn mlp + residual + BN statistics (gwnet.py:43-45, 233, 237)
        {
            Seg3A a{F(g.y[i]), F(g.x1[i]), F(g.x2[i]), C};
            Ld2D<true, 0> b{p->mlp_w[i], nullptr, 3 * C};
            double* st_i = stats + (size_t)i * 2 * C;
            if (C <= 64) {
                MlpEpi<1> e; memset(&e, 0, sizeof(e));
                e.bm = p->mlp_b[i]; e.up = uprev; e.ss = ss; e.U = F(g.u[i]); e.stats = st_i; e.g = lg; e.tc_bn = tcbn;
                launch_gemm<2, 1>(tc, M, C, 3 * C, 1, a, b, e, st);
            } else {
                MlpEpi<2> e; memset(&e, 0, sizeof(e));
                e.bm = p->mlp_b[i]; e.up = uprev; e.ss = ss; e.U = F(g.u[i]); e.stats = st_i; e.g = lg; e.tc_bn = tcbn;
                launch_gemm<2, 2>(tc, M, C, 3 * C, 1, a, b, e, st);
            }
            HOPK_LAUNCH_CHECK("mlp");
        }
        bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats + (size_t)i * 2 * C, (double)M, p->bn_w[i], p->bn_b[i],
                                                          p->bn_mean[i], p->bn_var[i], (long long*)p->bn_nbt[i],
                                                          F(g.mr) + (size_t)i * 2 * C, F(g.ss) + (size_t)(i + 1) * 2 * C, C,
                                                          s->training, s->bn_momentum, s->bn_eps);
        HOPK_LAUNCH_CHECK("bn_finalize");
        uprev = F(g.u[i]);
    }

    if (net) {
        void* kargs[] = {(void*)&na};
        const int grid = max_tiles < net_grid ? max_tiles : net_grid;
        HOPK_CUDA(cudaLaunchCooperativeKernel((const void*)fz_net_fwd_kernel, dim3(grid), dim3(256), kargs, fz_smem_bytes(), st));
        HOPK_LAUNCH_CHECK("fz_net_fwd");
    }
    // head (gwnet.py:240-246) on the last Tl steps only
    {
        int M = B * g.Tl * V;
        Ld2D<true, 0> a{F(g.ycat), nullptr, (long)L * C};
        SkipW b; b.C = C; for (int l = 0; l < L; ++l) b.w[l] = p->skip_w[l];
        SkipEpi<2> e; e.out = F(g.r0); e.L = L; e.N = s->S; for (int l = 0; l < L; ++l) e.b[l] = p->skip_b[l];
        launch_gemm<1, 2>(tc, M, s->S, L * C, 1, a, b, e, st);
        HOPK_LAUNCH_CHECK("skip");
        Ld2D<true, 0> a1{F(g.r0), nullptr, s->S};
        Ld2D<true, 0> b1{p->end1_w, nullptr, s->S};
        EpiStore<2> e1{F(g.r1), s->E, p->end1_b, nullptr, s->E, 1};
        launch_gemm<1, 2>(tc, M, s->E, s->S, 1, a1, b1, e1, st);
        HOPK_LAUNCH_CHECK("end1");
        Ld2D<true, 0> a2{F(g.r1), nullptr, s->E};
        Ld2D<true, 0> b2{p->end2_w, nullptr, s->E};
        RowMap rm{g.Tl, V, (long)s->out_dim * V * g.Tl, 1, g.Tl, (long)V * g.Tl};
        EpiStoreStrided<2> e2{out, rm, p->end2_b, s->out_dim};
        launch_gemm<1, 2>(tc, M, s->out_dim, s->E, 1, a2, b2, e2, st);
        HOPK_LAUNCH_CHECK("end2");
    }
    return 0;
}

extern "C" int hopk_gwnet_backward(const HopkGwnetShape* s, const HopkGwnetParams* p, const float* x, const int64_t xs[4],
                                   const float* dout, void* ws_, void* scratch_, const HopkGwnetGrads* gr, float* dx,
                                   void* stream)
{
    if (int rc = check_shape(s)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GwLayout g = make_layout(s);
    char* ws = (char*)ws_; char* sc = (char*)scratch_;
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    auto S = [&](size_t off) { return reinterpret_cast<float*>(sc + off); };
    const int C = s->C, V = s->V, B = s->B, L = s->L, Sk = s->S, E = s->E, O = s->out_dim;
    const bool tc = s->dtype == 1;
    const int tcbn = C <= 64 ? 64 : 128;
    double* bnsum = reinterpret_cast<double*>(sc + g.s_bnsum);
    HOPK_CUDA(cudaMemsetAsync(bnsum, 0, (size_t)(L + 1) * 2 * C * sizeof(double), st));
    HOPK_CUDA(cudaMemsetAsync(S(g.s_m12), 0, 2 * (size_t)V * V * sizeof(float), st));
    if (int rc = zero_param_grads(s, gr, st)) return rc;      // split-K / atomic epilogues accumulate into zeroed buffers

    SideStreams* sd = side_streams();
    if (!sd) return fail(3, "side streams", "cudaStreamCreate failed");
    SideJoinOnError join_guard{sd, st, true};
#ifdef HOPK_DEBUG
    static const bool skip_side = getenv("HOPK_BWD_SKIP_SIDE") != nullptr;     // timing experiments only: wrong gradients
#else
    constexpr bool skip_side = false;
#endif

    // ---- head backward: the dgrad GEMMs form the dependent chain, the three weight-gradient GEMMs go to the side streams
    const int M4 = B * g.Tl * V;
    {
        nchw_to_rows_kernel<<<cdiv((long)M4 * O, 256), 256, 0, st>>>(dout, S(g.s_dorow), B, O, V, g.Tl);
        HOPK_LAUNCH_CHECK("dout_rows");
        // end_conv_2
        {
            HOPK_CUDA(cudaEventRecord(sd->ev_head[0], st));
            HOPK_CUDA(cudaStreamWaitEvent(sd->s[0], sd->ev_head[0], 0));
            Ld2D<false, 0> a{S(g.s_dorow), nullptr, O};
            Ld2DOnes<false> b{F(g.r1), E, E};
            EpiWgrad<2> e{gr->end2_w, E, gr->end2_b, E, O};
            launch_gemm<1, 2>(tc, O, E + 1, M4, pick_splits(tc, O, E + 1, M4, 1, 2), a, b, e, sd->s[0]);
            HOPK_LAUNCH_CHECK("end2_wgrad");
            Ld2D<true, 0> a2{S(g.s_dorow), nullptr, O};
            Ld2D<false, 0> b2{p->end2_w, nullptr, E};
            EpiStore<2> e2{S(g.s_de1), E, nullptr, F(g.r1), E, 4};
            launch_gemm<1, 2>(tc, M4, E, O, 1, a2, b2, e2, st);
            HOPK_LAUNCH_CHECK("end2_dgrad");
        }
        // end_conv_1
        {
            HOPK_CUDA(cudaEventRecord(sd->ev_head[1], st));
            HOPK_CUDA(cudaStreamWaitEvent(sd->s[1], sd->ev_head[1], 0));
            Ld2D<false, 0> a{S(g.s_de1), nullptr, E};
            Ld2DOnes<false> b{F(g.r0), Sk, Sk};
            EpiWgrad<2> e{gr->end1_w, Sk, gr->end1_b, Sk, E};
            launch_gemm<1, 2>(tc, E, Sk + 1, M4, pick_splits(tc, E, Sk + 1, M4, 1, 2), a, b, e, sd->s[1]);
            HOPK_LAUNCH_CHECK("end1_wgrad");
            Ld2D<true, 0> a2{S(g.s_de1), nullptr, E};
            Ld2D<false, 0> b2{p->end1_w, nullptr, Sk};
            EpiStore<2> e2{S(g.s_dskip), Sk, nullptr, F(g.r0), Sk, 4};
            launch_gemm<1, 2>(tc, M4, Sk, E, 1, a2, b2, e2, st);
            HOPK_LAUNCH_CHECK("end1_dgrad");
        }
        // skip convs: one concat GEMM each way
        {
            HOPK_CUDA(cudaEventRecord(sd->ev_head[2], st));
            HOPK_CUDA(cudaStreamWaitEvent(sd->s[2], sd->ev_head[2], 0));
            Ld2D<false, 0> a{S(g.s_dskip), nullptr, Sk};
            Ld2DOnes<false> b{F(g.ycat), (long)L * C, L * C};
            SkipWgradEpi<2> e; e.C = C; e.L = L;
            for (int l = 0; l < L; ++l) { e.dw[l] = gr->skip_w[l]; e.db[l] = gr->skip_b[l]; }
            launch_gemm<1, 2>(tc, Sk, L * C + 1, M4, pick_splits(tc, Sk, L * C + 1, M4, 1, 2), a, b, e, sd->s[2]);
            HOPK_LAUNCH_CHECK("skip_wgrad");
            Ld2D<true, 0> a2{S(g.s_dskip), nullptr, Sk};
            SkipWT b2; b2.C = C; for (int l = 0; l < L; ++l) b2.w[l] = p->skip_w[l];
            EpiStore<2> e2{S(g.s_dycat), (long)L * C, nullptr, nullptr, L * C, 0};
            launch_gemm<1, 2>(tc, M4, L * C, Sk, 1, a2, b2, e2, st);
            HOPK_LAUNCH_CHECK("skip_dgrad");
        }
    }

    // ---- layers, last to first.  The caller's stream carries the dependent chain (BN backward -> node mix -> dy -> dx);
    // the weight-gradient GEMMs and the dA Gram products only feed parameter gradients, so they run on three side
    // streams forked from / joined back to the caller's stream with events (du, G, df, dg are kept per layer).
    float* dxn = nullptr;                 // gradient w.r.t. BN_i output (= layer i+1 input)
    float* dx_buf[2] = {S(g.s_dxa), S(g.s_dxb)};
    int flip = 0;
    for (int i = L - 1; i >= 0; --i) {
        LayerGeom lg{V, C, g.Tlen[i], g.Tlen[i + 1], s->dil[i]};
        const int M = B * lg.To * V;
        const int Min = B * lg.Ti * V;
        const float* uprev = i == 0 ? F(g.x0) : F(g.u[i - 1]);
        const float* ss = F(g.ss) + (size_t)i * 2 * C;
        const bool has_du = dxn != nullptr;
        float* DU = S(g.s_du[i]);
        float* DF = S(g.s_df[i]);
        float* DG = S(g.s_dg[i]);
        if (has_du) {
            size_t smem = 5 * C * sizeof(float);
            long n4 = (long)M * C / 4;
            int blocks = (int)((n4 + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
            bn_bwd_kernel<<<blocks, 256, smem, st>>>(dxn, F(g.u[i]), F(g.mr) + (size_t)i * 2 * C, p->bn_w[i],
                                                     bnsum + (size_t)i * 2 * C, (double)M, DU, gr->bn_w[i], gr->bn_b[i],
                                                     (size_t)M, C, s->training);
            HOPK_LAUNCH_CHECK("bn_bwd");
            int groups = B * lg.To;
            // P1 = A du, P2 = A^2 du  (node mix with the transposed supports)
            if (int rc = launch_node_mix(DU, F(g.At), F(g.A2t), S(g.s_p1[i]), S(g.s_p2[i]), groups, V, C, st)) return rc;
            HOPK_CUDA(cudaEventRecord(sd->ev_du[i], st));
            // side stream 0: mlp weight + bias gradient
            if (!skip_side) {
                cudaStream_t s0 = sd->s[0];
                HOPK_CUDA(cudaStreamWaitEvent(s0, sd->ev_du[i], 0));
                if (tc && C % 8 == 0) {
                    // dWm[o][seg*C + c] = sum_r du[r][o] x_seg[r][c] with x1 = A^T y, x2 = (A^2)^T y inside every node group
                    //                   = sum_r [du | A du | A^2 du][r][seg*C + o] y[r][c]:  the forward never stores y, x1, x2
                    W8Seg3 a{DU, S(g.s_p1[i]), S(g.s_p2[i]), C, 3 * C};
                    W8Prod b{F(g.tf[i]), F(g.sg[i]), C, C};
                    MlpWgEpi2 e{gr->mlp_w[i], gr->mlp_b[i], C};
                    HOPK_CUDA(launch_gemm_tc_wgrad<64>(M, 3 * C, C, a, b, e, true, s0));
                    HOPK_LAUNCH_CHECK("mlp_wgrad_tc");
                } else {
                    Ld2D<false, 0> a{DU, nullptr, C};
                    Seg3AT b{F(g.y[i]), F(g.x1[i]), F(g.x2[i]), C};
                    EpiWgrad<2> e{gr->mlp_w[i], (long)3 * C, gr->mlp_b[i], 3 * C, C};
                    if (C <= 64) launch_gemm<1, 2>(tc, C, 3 * C + 1, M, pick_splits(tc, C, 3 * C + 1, M, 1, 2), a, b, e, s0);
                    else launch_gemm<2, 2>(tc, C, 3 * C + 1, M, pick_splits(tc, C, 3 * C + 1, M, 2, 2), a, b, e, s0);
                    HOPK_LAUNCH_CHECK("mlp_wgrad");
                }
            }
            // side stream 1: G = du [Wm1 | Wm2]  and the Gram products for dA
            if (!skip_side) {
                cudaStream_t s1 = sd->s[1];
                HOPK_CUDA(cudaStreamWaitEvent(s1, sd->ev_du[i], 0));
                Ld2D<true, 0> a{DU, nullptr, C};
                Ld2D<false, 0> b{p->mlp_w[i] + C, nullptr, (long)3 * C};
                EpiStore<2> e{S(g.s_g[i]), (long)2 * C, nullptr, nullptr, 2 * C, 0};
                launch_gemm<2, 2>(tc, M, 2 * C, C, 1, a, b, e, s1);
                HOPK_LAUNCH_CHECK("g_gemm");
                if (tc && C % 8 == 0 && V * V >= 256 && 2 * V <= 128) {      // large graph: the Gram products as one UMMA GEMM
                    W8GramY ga{F(g.tf[i]), F(g.sg[i]), V, C};
                    W8GramG gb{S(g.s_g[i]), V, C};
                    GramEpi ge{S(g.s_m12), V};
                    HOPK_CUDA(launch_gemm_tc_wgrad<128>(groups * C, V, 2 * V, ga, gb, ge, false, s1));
                    HOPK_LAUNCH_CHECK("gram_tc");
                } else {
                int gpi = GRAM_GPI;                              // node groups staged per iteration, within the smem budget
                auto gram_smem = [&](int n) { return ((size_t)n * V * (3 * C + 2) + 2 * (size_t)V * V) * sizeof(float); };
                while (gpi > 1 && gram_smem(gpi) > 200 * 1024) --gpi;
                size_t smem3 = gram_smem(gpi);
                HOPK_REQUIRE(smem3 <= 200 * 1024, "gram kernel: V * C too large for shared memory");
                if (smem3 > 48 * 1024) HOPK_CUDA(configure_smem_once((const void*)gram_kernel, 200 * 1024));
                int gblocks = cdiv(groups, gpi); if (gblocks > 148) gblocks = 148;
                gram_kernel<<<gblocks, 256, smem3, s1>>>(F(g.tf[i]), F(g.sg[i]), S(g.s_g[i]), S(g.s_m12), groups, V, C, gpi);
                HOPK_LAUNCH_CHECK("gram");
                }
            }
        }
        // dy (+ skip path) -> df, dg
        {
            Seg3A a{DU, S(g.s_p1[i]), S(g.s_p2[i]), C};
            DyB b{p->mlp_w[i], C};
            int K = has_du ? 3 * C : 0;
            if (C <= 64) {
                DyEpi<1> e{F(g.tf[i]), F(g.sg[i]), S(g.s_dycat), DF, DG, lg, i, L, g.Tl};
                launch_gemm<2, 1>(tc, M, C, K, 1, a, b, e, st);
            } else {
                DyEpi<2> e{F(g.tf[i]), F(g.sg[i]), S(g.s_dycat), DF, DG, lg, i, L, g.Tl};
                launch_gemm<2, 2>(tc, M, C, K, 1, a, b, e, st);
            }
            HOPK_LAUNCH_CHECK("dy_gemm");
            HOPK_CUDA(cudaEventRecord(sd->ev_dfg[i], st));
        }
        // side stream 2: gate conv weight + bias gradients
        if (!skip_side) {
            cudaStream_t s2 = sd->s[2];
            HOPK_CUDA(cudaStreamWaitEvent(s2, sd->ev_dfg[i], 0));
            if (tc && C % 8 == 0) {
                W8Seg3 a{DF, DG, DG, C, 2 * C};
                W8GateXI b{uprev, ss, lg, 2 * C};
                GateWgEpi2 e{gr->filter_w[i], gr->gate_w[i], gr->filter_b[i], gr->gate_b[i], C};
                HOPK_CUDA(launch_gemm_tc_wgrad<128>(M, 2 * C, 2 * C, a, b, e, true, s2));
                HOPK_LAUNCH_CHECK("gate_wgrad_tc");
            } else {
                GateWgA a{DF, DG, C};
                GateWgB b{uprev, ss, lg};
                GateWgEpi<2> e{gr->filter_w[i], gr->gate_w[i], gr->filter_b[i], gr->gate_b[i], C};
                launch_gemm<2, 2>(tc, 2 * C, 2 * C + 1, M, pick_splits(tc, 2 * C, 2 * C + 1, M, 2, 2), a, b, e, s2);
                HOPK_LAUNCH_CHECK("gate_wgrad");
            }
        }
        // dx of the layer input (+ residual gradient) and BN-backward sums of layer i-1
        {
            DxA a{DF, DG, lg};
            DxB b{p->filter_w[i], p->gate_w[i], C};
            float* DX = dx_buf[flip];
            if (C <= 64) {
                DxEpi<1> e; memset(&e, 0, sizeof(e));
                e.DU = has_du ? DU : nullptr; e.DX = DX; e.uprev = uprev; e.g = lg; e.tc_bn = tcbn;
                // layer 0: the same sums (against the identity "statistics" mean 1 / rstd 0 of ss[0]) give the start-conv
                // bias gradient from the un-rounded fp32 gradient instead of a ones column of the bf16 weight-gradient GEMM
                e.mr_prev = i > 0 ? F(g.mr) + (size_t)(i - 1) * 2 * C : F(g.ss);
                e.sums_prev = i > 0 ? bnsum + (size_t)(i - 1) * 2 * C : bnsum + (size_t)L * 2 * C;
                launch_gemm<2, 1>(tc, Min, C, 4 * C, 1, a, b, e, st);
            } else {
                DxEpi<2> e; memset(&e, 0, sizeof(e));
                e.DU = has_du ? DU : nullptr; e.DX = DX; e.uprev = uprev; e.g = lg; e.tc_bn = tcbn;
                // layer 0: the same sums (against the identity "statistics" mean 1 / rstd 0 of ss[0]) give the start-conv
                // bias gradient from the un-rounded fp32 gradient instead of a ones column of the bf16 weight-gradient GEMM
                e.mr_prev = i > 0 ? F(g.mr) + (size_t)(i - 1) * 2 * C : F(g.ss);
                e.sums_prev = i > 0 ? bnsum + (size_t)(i - 1) * 2 * C : bnsum + (size_t)L * 2 * C;
                launch_gemm<2, 2>(tc, Min, C, 4 * C, 1, a, b, e, st);
            }
            HOPK_LAUNCH_CHECK("dx_gemm");
            dxn = DX; flip ^= 1;
        }
    }
    // ---- adaptive adjacency (side stream 1, behind the Gram kernels it consumes) and start conv
    HOPK_CUDA(cudaEventRecord(sd->ev_tail, st));                 // the gradient w.r.t. the start conv output is complete
    adp_bwd_kernel<<<1, 256, V * V * sizeof(float), sd->s[1]>>>(p->nodevec1, p->nodevec2, F(g.A), F(g.Z), S(g.s_m12),
                                                                 S(g.s_m12) + V * V, V, s->rank, gr->nodevec1, gr->nodevec2);
    HOPK_LAUNCH_CHECK("adp_bwd");
    {
        int M = B * g.Tp * V;
        HOPK_CUDA(cudaStreamWaitEvent(sd->s[0], sd->ev_tail, 0));
        EpiWgrad<2> e{gr->start_w, s->in_dim, nullptr, s->in_dim, C};
        const bool rows_input = g.pad == 0 && xs[1] == 1 && xs[3] == (long)V * xs[2] && xs[0] == (long)g.Tp * xs[3];
        if (tc && rows_input) {
            // the input is the (b, t, v)-rows matrix with contiguous channels (what HOP.forecast hands over): dW = dX0^T . X is a
            // plain rows-contraction for the weight-gradient skeleton (both operands MN-major as they lie in memory) instead of
            // the gathering generic loader -- this kernel is the tail of the backward (83 -> ~20 us)
            W8Plain a{dxn, nullptr, C, C, 0};
            W8Plain b{x, nullptr, (long)xs[2], s->in_dim, 0};
            HOPK_CUDA(launch_gemm_tc_wgrad<128>(M, C, s->in_dim, a, b, e, false, sd->s[0]));
        } else {
            Ld2D<false, 0> a{dxn, nullptr, C};
            StartAT b{x, g.Tp, V, g.pad, s->in_dim, (long)xs[0], (long)xs[1], (long)xs[2], (long)xs[3]};
            if (C <= 64) launch_gemm<1, 2>(tc, C, s->in_dim + 1, M, pick_splits(tc, C, s->in_dim + 1, M, 1, 2), a, b, e, sd->s[0]);
            else launch_gemm<2, 2>(tc, C, s->in_dim + 1, M, pick_splits(tc, C, s->in_dim + 1, M, 2, 2), a, b, e, sd->s[0]);
        }
        HOPK_LAUNCH_CHECK("start_wgrad");
        if (gr->start_b) {
            sums_to_float_kernel<<<cdiv(C, 128), 128, 0, st>>>(bnsum + (size_t)L * 2 * C, gr->start_b, C);
            HOPK_LAUNCH_CHECK("start_bias_grad");
        }
        if (dx) {
            Ld2D<true, 0> a2{dxn, nullptr, C};
            Ld2D<false, 0> b2{p->start_w, nullptr, s->in_dim};
            StartDxEpi e2{dx, g.Tp, s->T, V, g.pad, s->in_dim};
            launch_gemm<2, 2>(tc, M, s->in_dim, C, 1, a2, b2, e2, st);
            HOPK_LAUNCH_CHECK("start_dgrad");
        }
    }
    // join the side streams: everything the caller enqueues next sees the finished gradients
    for (int k = 0; k < 3; ++k) {
        HOPK_CUDA(cudaEventRecord(sd->ev_join[k], sd->s[k]));
        HOPK_CUDA(cudaStreamWaitEvent(st, sd->ev_join[k], 0));
    }
    join_guard.armed = false;
    return 0;
}


// ============================================================================ per-op entry points
// The same kernels that hopk_gwnet_forward / backward chain together (the generic, non-fused path), exposed one operator at
// a time so that each can be tested and timed in isolation through the C ABI.  All activations in the rows layout.
namespace hopk {
__global__ void gate_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ tf, const float* __restrict__ sg,
                                float* __restrict__ df, float* __restrict__ dg, size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 d = reinterpret_cast<const float4*>(dy)[i], t = reinterpret_cast<const float4*>(tf)[i], s = reinterpret_cast<const float4*>(sg)[i];
        reinterpret_cast<float4*>(df)[i] = make_float4(d.x * s.x * (1.f - t.x * t.x), d.y * s.y * (1.f - t.y * t.y), d.z * s.z * (1.f - t.z * t.z),
                                                       d.w * s.w * (1.f - t.w * t.w));
        reinterpret_cast<float4*>(dg)[i] = make_float4(d.x * t.x * s.x * (1.f - s.x), d.y * t.y * s.y * (1.f - s.y), d.z * t.z * s.z * (1.f - s.z),
                                                       d.w * t.w * s.w * (1.f - s.w));
    }
}
// dA = M1 + A^T M2 + M2 A^T  (two diffusion hops, Appendix A)
__global__ void da_combine_kernel(const float* __restrict__ A, const float* __restrict__ M1, const float* __restrict__ M2, float* __restrict__ dA, int V)
{
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
        const int v = idx / V, w = idx % V;
        float acc = M1[idx];
        for (int k = 0; k < V; ++k) acc += A[k * V + v] * M2[k * V + w] + M2[v * V + k] * A[w * V + k];
        dA[idx] = acc;
    }
}
}  // namespace hopk

static int op_check(int B, int V, int Ti, int d, int C, int dtype)
{
    HOPK_REQUIRE(B >= 1 && V >= 1 && V <= 45 && d >= 1 && Ti > d, "op sizes (V <= 45, Ti > d)");
    HOPK_REQUIRE(C % 4 == 0 && C >= 4 && C <= 256, "C must be a multiple of 4, <= 256");
    HOPK_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 or 1");
    HOPK_REQUIRE((long)B * V * Ti * 256 < (1L << 31), "tensor too large for 32-bit row indices");
    return 0;
}

extern "C" int hopk_adp_softmax_fwd(const float* e1, const float* e2, int V, int R, float* A5, void* stream)
{
    HOPK_REQUIRE(V >= 1 && V <= 45 && R >= 1, "adp_softmax sizes");
    const size_t vv = (size_t)V * V;
    adp_fwd_kernel<<<1, 256, V * V * sizeof(float), (cudaStream_t)stream>>>(e1, e2, V, R, A5, A5 + vv, A5 + 2 * vv, A5 + 3 * vv, A5 + 4 * vv);
    HOPK_LAUNCH_CHECK("adp_softmax_fwd");
    return 0;
}

extern "C" int hopk_adp_softmax_bwd(const float* e1, const float* e2, const float* A5, const float* dA, int V, int R, float* de1, float* de2,
                                    void* stream)
{
    HOPK_REQUIRE(V >= 1 && V <= 45 && R >= 1, "adp_softmax sizes");
    const size_t vv = (size_t)V * V;
    adp_bwd_kernel<<<1, 256, V * V * sizeof(float), (cudaStream_t)stream>>>(e1, e2, A5, A5 + 4 * vv, dA, nullptr, V, R, de1, de2);
    HOPK_LAUNCH_CHECK("adp_softmax_bwd");
    return 0;
}

extern "C" int hopk_gated_tcn_fwd(const float* x, const float* ss, const float* wf, const float* bf, const float* wg, const float* bg, int B,
                                  int V, int Ti, int d, int C, int dtype, float* tf, float* sg, float* y, void* stream)
{
    if (int rc = op_check(B, V, Ti, d, C, dtype)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc = dtype == 1;
    LayerGeom lg{V, C, Ti, Ti - d, d};
    const int M = B * lg.To * V;
    GateA a{x, ss, lg};
    GateB b{wf, wg, C, tc ? 1 : 0};
    GateEpi e{bf, bg, tf, sg, y, nullptr, lg, 0, 1, 0};          // Tl = 0: no skip slice
    const int Nlog = tc ? cdiv(C, 16) * 32 : cdiv(C, 64) * 128;
    launch_gemm<2, 2>(tc, M, Nlog, 2 * C, 1, a, b, e, st);
    HOPK_LAUNCH_CHECK("gated_tcn_fwd");
    return 0;
}

extern "C" size_t hopk_gated_tcn_bwd_scratch_bytes(int B, int V, int Ti, int d, int C) { return (size_t)2 * B * (Ti - d) * V * C * sizeof(float); }

extern "C" int hopk_gated_tcn_bwd(const float* x, const float* ss, const float* wf, const float* wg, const float* tf, const float* sg,
                                  const float* dy, int B, int V, int Ti, int d, int C, int dtype, void* scratch, float* dx, float* dwf,
                                  float* dbf, float* dwg, float* dbg, void* stream)
{
    if (int rc = op_check(B, V, Ti, d, C, dtype)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc = dtype == 1;
    LayerGeom lg{V, C, Ti, Ti - d, d};
    const int M = B * lg.To * V, Min = B * Ti * V;
    float* DF = (float*)scratch;
    float* DG = DF + (size_t)M * C;
    const size_t n4 = (size_t)M * C / 4;
    gate_bwd_kernel<<<cdiv((long)n4, 256) > 148 * 8 ? 148 * 8 : cdiv((long)n4, 256), 256, 0, st>>>(dy, tf, sg, DF, DG, n4);
    HOPK_LAUNCH_CHECK("gate_bwd");
    HOPK_CUDA(cudaMemsetAsync(dwf, 0, (size_t)C * 2 * C * sizeof(float), st));
    HOPK_CUDA(cudaMemsetAsync(dwg, 0, (size_t)C * 2 * C * sizeof(float), st));
    HOPK_CUDA(cudaMemsetAsync(dbf, 0, (size_t)C * sizeof(float), st));
    HOPK_CUDA(cudaMemsetAsync(dbg, 0, (size_t)C * sizeof(float), st));
    if (tc && C % 8 == 0) {
        W8Seg3 a{DF, DG, DG, C, 2 * C};
        W8GateXI b{x, ss, lg, 2 * C};
        GateWgEpi2 e{dwf, dwg, dbf, dbg, C};
        HOPK_CUDA(launch_gemm_tc_wgrad<128>(M, 2 * C, 2 * C, a, b, e, true, st));
    } else {
        GateWgA a{DF, DG, C};
        GateWgB b{x, ss, lg};
        GateWgEpi<2> e{dwf, dwg, dbf, dbg, C};
        launch_gemm<2, 2>(tc, 2 * C, 2 * C + 1, M, pick_splits(tc, 2 * C, 2 * C + 1, M, 2, 2), a, b, e, st);
    }
    HOPK_LAUNCH_CHECK("gated_tcn_wgrad");
    if (dx) {
        DxA a{DF, DG, lg};
        DxB b{wf, wg, C};
        if (C <= 64) {
            DxEpi<1> e; memset(&e, 0, sizeof(e));
            e.DX = dx; e.g = lg; e.tc_bn = 64;
            launch_gemm<2, 1>(tc, Min, C, 4 * C, 1, a, b, e, st);
        } else {
            DxEpi<2> e; memset(&e, 0, sizeof(e));
            e.DX = dx; e.g = lg; e.tc_bn = 128;
            launch_gemm<2, 2>(tc, Min, C, 4 * C, 1, a, b, e, st);
        }
        HOPK_LAUNCH_CHECK("gated_tcn_dx");
    }
    return 0;
}

extern "C" size_t hopk_gcn_scratch_bytes(int B, int V, int To, int C)
{
    return ((size_t)4 * B * To * V * C + (size_t)3 * V * V) * sizeof(float) + 1024;
}

/* u = Wm . [y | A^T y | (A^2)^T y] + bm + BN-folded residual x[t + d]; stats = per-channel sum / sum^2 of u (doubles, 2C) */
extern "C" int hopk_gcn_diffuse_mlp_res_bnstat_fwd(const float* y, const float* A5, const float* xres, const float* ss, const float* wm,
                                                   const float* bm, int B, int V, int Ti, int d, int C, int dtype, float* x1, float* x2,
                                                   float* u, double* stats, void* stream)
{
    if (int rc = op_check(B, V, Ti, d, C, dtype)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc = dtype == 1;
    LayerGeom lg{V, C, Ti, Ti - d, d};
    const int M = B * lg.To * V;
    const size_t vv = (size_t)V * V;
    HOPK_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * C * sizeof(double), st));
    if (int rc = launch_node_mix(y, A5, A5 + vv, x1, x2, B * lg.To, V, C, st)) return rc;
    Seg3A a{y, x1, x2, C};
    Ld2D<true, 0> b{wm, nullptr, 3 * C};
    if (C <= 64) {
        MlpEpi<1> e; memset(&e, 0, sizeof(e));
        e.bm = bm; e.up = xres; e.ss = ss; e.U = u; e.stats = stats; e.g = lg; e.tc_bn = 64;
        launch_gemm<2, 1>(tc, M, C, 3 * C, 1, a, b, e, st);
    } else {
        MlpEpi<2> e; memset(&e, 0, sizeof(e));
        e.bm = bm; e.up = xres; e.ss = ss; e.U = u; e.stats = stats; e.g = lg; e.tc_bn = 128;
        launch_gemm<2, 2>(tc, M, C, 3 * C, 1, a, b, e, st);
    }
    HOPK_LAUNCH_CHECK("gcn_mlp_fwd");
    return 0;
}

/* given du (gradient w.r.t. u; also the gradient of the residual branch): dy, dWm, dbm, dA */
extern "C" int hopk_gcn_diffuse_mlp_res_bnstat_bwd(const float* du, const float* y, const float* x1, const float* x2, const float* A5,
                                                   const float* wm, int B, int V, int To, int C, int dtype, void* scratch, float* dy,
                                                   float* dwm, float* dbm, float* dA, void* stream)
{
    if (int rc = op_check(B, V, To + 1, 1, C, dtype)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc = dtype == 1;
    const int M = B * To * V, groups = B * To;
    const size_t vv = (size_t)V * V;
    float* P1 = (float*)(((uintptr_t)scratch + 255) & ~uintptr_t(255));
    float* P2 = P1 + (size_t)M * C;
    float* G = P2 + (size_t)M * C;
    float* m12 = G + (size_t)2 * M * C;
    if (int rc = launch_node_mix(du, A5 + 2 * vv, A5 + 3 * vv, P1, P2, groups, V, C, st)) return rc;      // A du, A^2 du
    {
        Seg3A a{du, P1, P2, C};
        DyB b{wm, C};
        if (C <= 64) { EpiStore<1> e{dy, C, nullptr, nullptr, C, 0}; launch_gemm<2, 1>(tc, M, C, 3 * C, 1, a, b, e, st); }
        else { EpiStore<2> e{dy, C, nullptr, nullptr, C, 0}; launch_gemm<2, 2>(tc, M, C, 3 * C, 1, a, b, e, st); }
        HOPK_LAUNCH_CHECK("gcn_dy");
    }
    HOPK_CUDA(cudaMemsetAsync(dwm, 0, (size_t)C * 3 * C * sizeof(float), st));
    HOPK_CUDA(cudaMemsetAsync(dbm, 0, (size_t)C * sizeof(float), st));
    {
        Ld2D<false, 0> a{du, nullptr, C};
        Seg3AT b{y, x1, x2, C};
        EpiWgrad<2> e{dwm, (long)3 * C, dbm, 3 * C, C};
        if (C <= 64) launch_gemm<1, 2>(tc, C, 3 * C + 1, M, pick_splits(tc, C, 3 * C + 1, M, 1, 2), a, b, e, st);
        else launch_gemm<2, 2>(tc, C, 3 * C + 1, M, pick_splits(tc, C, 3 * C + 1, M, 2, 2), a, b, e, st);
        HOPK_LAUNCH_CHECK("gcn_wgrad");
    }
    {
        Ld2D<true, 0> a{du, nullptr, C};
        Ld2D<false, 0> b{wm + C, nullptr, (long)3 * C};
        EpiStore<2> e{G, (long)2 * C, nullptr, nullptr, 2 * C, 0};
        launch_gemm<2, 2>(tc, M, 2 * C, C, 1, a, b, e, st);
        HOPK_LAUNCH_CHECK("gcn_g");
        HOPK_CUDA(cudaMemsetAsync(m12, 0, 2 * vv * sizeof(float), st));
        int gpi = GRAM_GPI;
        auto gram_smem = [&](int n) { return ((size_t)n * V * (3 * C + 2) + 2 * (size_t)V * V) * sizeof(float); };
        while (gpi > 1 && gram_smem(gpi) > 200 * 1024) --gpi;
        const size_t smem3 = gram_smem(gpi);
        HOPK_REQUIRE(smem3 <= 200 * 1024, "gram kernel: V * C too large for shared memory");
        if (smem3 > 48 * 1024) HOPK_CUDA(configure_smem_once((const void*)gram_kernel, 200 * 1024));
        int gblocks = cdiv(groups, gpi); if (gblocks > 148) gblocks = 148;
        gram_kernel<<<gblocks, 256, smem3, st>>>(y, nullptr, G, m12, groups, V, C, gpi);
        HOPK_LAUNCH_CHECK("gcn_gram");
        da_combine_kernel<<<1, 256, 0, st>>>(A5, m12, m12 + vv, dA, V);
        HOPK_LAUNCH_CHECK("gcn_dA");
    }
    return 0;
}

extern "C" int hopk_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float* rmean, float* rvar,
                                int64_t* nbt, float* mean_rstd, float* scale_shift, int C, int training, float momentum, float eps, void* stream)
{
    HOPK_REQUIRE(C >= 1 && count >= 1 && eps > 0.f, "bn_finalize sizes");
    bn_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, rmean, rvar, (long long*)nbt, mean_rstd,
                                                                        scale_shift, C, training, momentum, eps);
    HOPK_LAUNCH_CHECK("bn_finalize");
    return 0;
}
