// glue.cu -- index-map kernels of HOP.Model.forecast around the Graph-WaveNet block (reference model/HOP.py:210-217).
//
// The reference runs the beat MLP on `in_audio.unfold(1, 3400, 2191).unsqueeze(1).repeat(1, J, 1, 1)` -- the same 16 windows
// J times -- and then REINTERPRETS (B, J, 16, 170) as (B, 16, J, 170) (HOP.py:210-212): the feature at [b, t, j] is
// beat(window[b, (t*J + j) % 16]).  Here the MLP runs once per window (GEMMs in gemm_tma.cu) and these kernels apply the
// index map while writing / reading Graph-WaveNet's rows buffer (B, 16, J, 3 + F) = [x, y, z of bone j at frame t | F features].
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

__global__ void beat_rows_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ seed, float* __restrict__ rows,
                                     int B, int J, int F, int NW)
{
    const int W = 3 + F;
    const size_t n = (size_t)B * NW * J * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % W);
        const size_t r = i / W;                            // (b*NW + t)*J + j
        const int j = (int)(r % J);
        const size_t bt = r / J;
        const int t = (int)(bt % NW);
        const size_t b = bt / NW;
        float v;
        if (c < 3) v = seed[(b * NW + t) * (size_t)(3 * J) + j * 3 + c];
        else v = feat[(b * NW + (size_t)((t * J + j) % NW)) * F + (c - 3)];
        rows[i] = v;
    }
}

// dfeat[b*NW + k][f] = sum over the J positions p = k, k + NW, ... (p = t*J + j) of drows[b, p / J, p % J, 3 + f]  -> bf16
__global__ void beat_rows_bwd_kernel(const float* __restrict__ drows, __nv_bfloat16* __restrict__ dfeat, int B, int J, int F, int NW, long ldd)
{
    const int W = 3 + F;
    const size_t n = (size_t)B * NW * ldd;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % ldd);
        const size_t bk = i / ldd;
        const int k = (int)(bk % NW);
        const size_t b = bk / NW;
        float acc = 0.f;
        if (f < F) {
            for (int q = 0; q < J; ++q) {
                const int p = k + q * NW;
                acc += drows[((b * NW + p / J) * J + p % J) * (size_t)W + 3 + f];
            }
        }
        dfeat[i] = __float2bfloat16_rn(acc);
    }
}


// ---------------------------------------------------------------- generator losses of the training step (train_llm.py:46-79)
//   huber   = mean smooth_l1(out / 0.1, tgt / 0.1) * 0.1
//   pose[b] = sum_{t,p} smooth_l1(out / 0.05, rand / 0.05) * 0.05,   zl1[b] = mean_z |zc - zr|
//   div_reg = mean_b clamp(-pose[b] / (zl1[b] + 1e-5), min = -1000)
//   kld     = -0.5 mean(1 + logvar - mu^2 - exp(logvar))
//   loss    = w_reg huber + w_div div_reg + w_kld kld
// One block per sample accumulates its three partial sums; the last block to finish combines them (fixed order over the
// samples: reproducible).  The reference spends about 65 elementwise / reduction launches on this, forward + backward.
__device__ __forceinline__ float sl1(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float sl1_grad(float d) { return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f); }

__device__ __forceinline__ float block_sum(float v, float* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;
}

// per[b] = {sum of the huber terms, pose[b], zl1[b], sum of the kld terms}; vals = {loss, huber, div_reg, kld}
__global__ void __launch_bounds__(256) step_losses_fwd_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                              const float* __restrict__ rnd, const float* __restrict__ zc,
                                                              const float* __restrict__ zr, const float* __restrict__ mu,
                                                              const float* __restrict__ logvar, int B, int TP, int Z, float w_reg,
                                                              float w_div, float w_kld, float* __restrict__ per,
                                                              float* __restrict__ vals, unsigned int* __restrict__ ticket)
{
    __shared__ float sh[8];
    __shared__ unsigned int last;
    const int b = blockIdx.x;
    float h = 0.f, ps = 0.f, zl = 0.f, kl = 0.f;
    for (int i = threadIdx.x; i < TP; i += blockDim.x) {
        const float o = out[(size_t)b * TP + i];
        h += sl1((o - tgt[(size_t)b * TP + i]) * 10.f);
        if (rnd) ps += sl1((o - rnd[(size_t)b * TP + i]) * 20.f);
    }
    for (int i = threadIdx.x; i < Z; i += blockDim.x) {
        if (zc) zl += fabsf(zc[(size_t)b * Z + i] - zr[(size_t)b * Z + i]);
        if (mu) { const float m = mu[(size_t)b * Z + i], lv = logvar[(size_t)b * Z + i]; kl += 1.f + lv - m * m - __expf(lv); }
    }
    h = block_sum(h, sh); ps = block_sum(ps, sh); zl = block_sum(zl, sh); kl = block_sum(kl, sh);
    if (threadIdx.x == 0) {
        per[4 * b] = h; per[4 * b + 1] = ps * 0.05f; per[4 * b + 2] = zl / (float)Z; per[4 * b + 3] = kl;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float sh_ = 0.f, sd = 0.f, sk = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        sh_ += __ldcg(per + 4 * i);
        if (rnd) sd += fmaxf(-__ldcg(per + 4 * i + 1) / (__ldcg(per + 4 * i + 2) + 1.0e-5f), -1000.f);
        sk += __ldcg(per + 4 * i + 3);
    }
    sh_ = block_sum(sh_, sh); sd = block_sum(sd, sh); sk = block_sum(sk, sh);
    if (threadIdx.x == 0) {
        const float huber = sh_ / ((float)B * (float)TP) * 0.1f;
        const float div = rnd ? sd / (float)B : 0.f;
        const float kld = mu ? -0.5f * sk / ((float)B * (float)Z) : 0.f;
        vals[0] = w_reg * huber + w_div * div + w_kld * kld;
        vals[1] = huber; vals[2] = div; vals[3] = kld;
        *ticket = 0u;                                              // ready for the next call (stream-ordered)
    }
}

// d loss / d out, d mu, d logvar, scaled by the upstream gradient *gl (a device scalar)
__global__ void __launch_bounds__(256) step_losses_bwd_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                              const float* __restrict__ rnd, const float* __restrict__ mu,
                                                              const float* __restrict__ logvar, const float* __restrict__ per,
                                                              const float* __restrict__ gl, int B, int TP, int Z, float w_reg,
                                                              float w_div, float w_kld, float* __restrict__ dout,
                                                              float* __restrict__ dmu, float* __restrict__ dlogvar)
{
    const int b = blockIdx.x;
    const float g = *gl;
    const float ch = g * w_reg / ((float)B * (float)TP);          // huber: 0.1 * (1 / 0.1) cancel
    float cd = 0.f;
    if (rnd) {
        const float zl = per[4 * b + 2] + 1.0e-5f;
        const bool active = -per[4 * b + 1] / zl > -1000.f;       // clamp(min = -1000) passes the gradient where it is not active
        cd = active ? -g * w_div / ((float)B * zl) : 0.f;        // d pose / d out = sl1'(.) : 0.05 * (1 / 0.05) cancel
    }
    for (int i = threadIdx.x; i < TP; i += blockDim.x) {
        const float o = out[(size_t)b * TP + i];
        float d = ch * sl1_grad((o - tgt[(size_t)b * TP + i]) * 10.f);
        if (rnd) d += cd * sl1_grad((o - rnd[(size_t)b * TP + i]) * 20.f);
        dout[(size_t)b * TP + i] = d;
    }
    if (mu) {
        const float ck = g * w_kld / ((float)B * (float)Z);
        for (int i = threadIdx.x; i < Z; i += blockDim.x) {
            dmu[(size_t)b * Z + i] = ck * mu[(size_t)b * Z + i];
            dlogvar[(size_t)b * Z + i] = -0.5f * ck * (1.f - __expf(logvar[(size_t)b * Z + i]));
        }
    }
}

static int grid_for(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : b); }

}  // namespace hopk
using namespace hopk;

extern "C" int hopk_beat_rows_fwd(const float* feat, const float* seed, float* rows, int B, int J, int F, void* stream)
{
    HOPK_REQUIRE(B > 0 && J > 0 && F > 0, "beat_rows sizes");
    const int NW = 16;
    beat_rows_fwd_kernel<<<grid_for((size_t)B * NW * J * (3 + F)), 256, 0, (cudaStream_t)stream>>>(feat, seed, rows, B, J, F, NW);
    HOPK_LAUNCH_CHECK("beat_rows_fwd");
    return 0;
}

extern "C" int hopk_beat_rows_bwd(const float* drows, void* dfeat_bf16, float* dbias, int B, int J, int F, long ldd, void* stream)
{
    HOPK_REQUIRE(B > 0 && J > 0 && F > 0 && ldd >= F && ldd % 8 == 0, "beat_rows_bwd sizes (ldd multiple of 8, >= F)");
    const int NW = 16;
    beat_rows_bwd_kernel<<<grid_for((size_t)B * NW * ldd), 256, 0, (cudaStream_t)stream>>>(drows, (__nv_bfloat16*)dfeat_bf16, B, J, F, NW, ldd);
    HOPK_LAUNCH_CHECK("beat_rows_bwd");
    if (dbias) return hopk_colsum(dfeat_bf16, dbias, (long)B * NW, F, ldd, 1, stream);
    return 0;
}

extern "C" int hopk_step_losses_fwd(const float* out, const float* tgt, const float* rnd, const float* zc, const float* zr, const float* mu,
                                    const float* logvar, int B, int TP, int Z, float w_reg, float w_div, float w_kld, float* per, float* vals,
                                    unsigned int* ticket, void* stream)
{
    HOPK_REQUIRE(B > 0 && TP > 0 && out && tgt && per && vals && ticket, "step losses: sizes / buffers");
    HOPK_REQUIRE((rnd == nullptr) == (zc == nullptr) && (zc == nullptr) == (zr == nullptr), "step losses: rnd, zc, zr come together");
    HOPK_REQUIRE((mu == nullptr) == (logvar == nullptr) && (Z > 0 || (!mu && !zc)), "step losses: mu / logvar come together");
    step_losses_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(out, tgt, rnd, zc, zr, mu, logvar, B, TP, Z, w_reg, w_div, w_kld, per, vals, ticket);
    HOPK_LAUNCH_CHECK("step_losses_fwd");
    return 0;
}

extern "C" int hopk_step_losses_bwd(const float* out, const float* tgt, const float* rnd, const float* mu, const float* logvar,
                                    const float* per, const float* gl, int B, int TP, int Z, float w_reg, float w_div, float w_kld,
                                    float* dout, float* dmu, float* dlogvar, void* stream)
{
    HOPK_REQUIRE(B > 0 && TP > 0 && out && tgt && per && gl && dout, "step losses backward: sizes / buffers");
    HOPK_REQUIRE(!mu || (logvar && dmu && dlogvar), "step losses backward: mu / logvar gradients");
    step_losses_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(out, tgt, rnd, mu, logvar, per, gl, B, TP, Z, w_reg, w_div, w_kld, dout, dmu, dlogvar);
    HOPK_LAUNCH_CHECK("step_losses_bwd");
    return 0;
}
