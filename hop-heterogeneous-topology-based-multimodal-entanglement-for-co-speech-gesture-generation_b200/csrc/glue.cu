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
