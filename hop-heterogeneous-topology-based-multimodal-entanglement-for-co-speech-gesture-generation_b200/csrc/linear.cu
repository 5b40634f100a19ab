// linear.cu -- dense layers of the hot path on the fp32-exact FFMA skeleton:
//   * hopk_linear_{fwd,bwd}: the reprogramming Q/K/V/O projections (reference model/HOP.py:262-265,276-285)
//   * hopk_conv1x1_nchw_{fwd,bwd}: gwnet's `linear` module used stand-alone (model/gwnet.py:16-22)
//   * hopk_nconv_{fwd,bwd}: gwnet's `nconv` module used stand-alone (model/gwnet.py:8-14)
#include <type_traits>
#include "functors.cuh"
#include "gemm_tc.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

template <class T, class = void> struct has_row32 : std::false_type {};
template <class T> struct has_row32<T, std::void_t<decltype(&T::row32)>> : std::true_type {};

static thread_local bool g_tc = false;     // set per call from flag 0x100: bf16 tensor-core math (tcgen05)

template <int MG, int NG, class AL, class BL, class EP>
static void launch_gemm2(int M, int N, int K, int splits, AL a, BL b, EP e, cudaStream_t st)
{
    if constexpr (has_row32<EP>::value) {
        if (g_tc) {
            if (N <= 64) launch_gemm_tc<64>(M, N, K, splits, a, b, e, st);
            else launch_gemm_tc<128>(M, N, K, splits, a, b, e, st);
            return;
        }
    }
    int kper = K;
    if (splits > 1) { kper = ((cdiv(K, splits) + GEMM_BK - 1) / GEMM_BK) * GEMM_BK; splits = cdiv(K, kper); }
    if (splits < 1) splits = 1;
    dim3 grid(cdiv(N, 64 * NG), cdiv(M, 64 * MG), splits);
    gemm_kernel<MG, NG, AL, BL, EP><<<grid, GEMM_THREADS, 0, st>>>(M, N, K, kper, a, b, e);
}

static int splits_for(int M, int N, int K, int mg, int ng)
{
    if (g_tc) { mg = 2; ng = N <= 64 ? 1 : 2; }
    long tiles = (long)cdiv(M, 64 * mg) * cdiv(N, 64 * ng);
    long want = (2 * 148 + tiles - 1) / tiles;
    long maxs = K / 256 > 0 ? K / 256 : 1;
    return (int)(want < maxs ? want : maxs);
}

// NCHW 1x1-conv operand views: row m = (b, v, t) flattened as b*V*T + vt ; element (m, c) at b*C*VT + c*VT + vt
struct NchwA {
    static constexpr bool kFast = false;         // rows (vt) are the contiguous index
    const float* x; int C, VT;
    __device__ __forceinline__ float operator()(int m, int k) const {
        int b = m / VT, vt = m - b * VT;
        return __ldg(x + ((size_t)b * C + k) * VT + vt);
    }
};
struct NchwAT {                                   // B'(kout, m) with ones column at kout == C
    static constexpr bool kFast = true;           // contraction index m is contiguous
    const float* x; int C, VT;
    __device__ __forceinline__ float operator()(int kout, int m) const {
        if (kout == C) return 1.f;
        int b = m / VT, vt = m - b * VT;
        return __ldg(x + ((size_t)b * C + kout) * VT + vt);
    }
};
struct NchwATplain {                              // A'(n, m) = dy[b, n, vt]
    static constexpr bool kFast = true;
    const float* x; int C, VT;
    __device__ __forceinline__ float operator()(int n, int m) const {
        int b = m / VT, vt = m - b * VT;
        return __ldg(x + ((size_t)b * C + n) * VT + vt);
    }
};
template <int NG>
struct NchwEpi {
    float* y; const float* bias; int N, VT;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
        int b = m / VT, vt = m - b * VT;
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * g + j;
                if (n < N) y[((size_t)b * N + n) * VT + vt] = v[4 * g + j] + (bias ? __ldg(bias + n) : 0.f);
            }
    }
    __device__ __forceinline__ void flush(int) {}
};

struct XRelu {                                    // B'(kout, m) = relu(x[m][kout]), ones column at kout == ones_at
    static constexpr bool kFast = false;
    const float* p; long ld; int ones_at;
    __device__ __forceinline__ float operator()(int i, int k) const {
        if (i == ones_at) return 1.f;
        return fmaxf(__ldg(p + (size_t)k * ld + i), 0.f);
    }
};

// stand-alone nconv on NCHW: out[n,c,w,l] = sum_v x[n,c,v,l] A[v,w]
__global__ void nconv_nchw_kernel(const float* __restrict__ x, const float* __restrict__ A, float* __restrict__ out,
                                  size_t NC, int V, int T, int transpose_a)
{
    extern __shared__ float sA[];
    for (int i = threadIdx.x; i < V * V; i += blockDim.x) {
        int v = i / V, w = i % V;
        sA[i] = transpose_a ? A[w * V + v] : A[i];
    }
    __syncthreads();
    size_t total = NC * V * T;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int l = (int)(i % T); size_t r = i / T; int w = (int)(r % V); size_t nc = r / V;
        const float* xp = x + nc * V * T + l;
        float acc = 0.f;
        for (int v = 0; v < V; ++v) acc = fmaf(__ldg(xp + (size_t)v * T), sA[v * V + w], acc);
        out[i] = acc;
    }
}
// dA[v][w] = sum_{n,c,l} x[n,c,v,l] dout[n,c,w,l]
__global__ void nconv_dA_kernel(const float* __restrict__ x, const float* __restrict__ dout, float* __restrict__ dA,
                                size_t NC, int V, int T)
{
    int pair = blockIdx.x;                  // one CTA per (v, w)
    int v = pair / V, w = pair % V;
    double acc = 0.0;
    size_t n = NC * T;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        size_t nc = i / T; int l = (int)(i % T);
        acc += (double)x[(nc * V + v) * T + l] * (double)dout[(nc * V + w) * T + l];
    }
    acc = warp_sum(acc);
    __shared__ double red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
        dA[pair] = (float)s;
    }
}

}  // namespace hopk
using namespace hopk;

extern "C" int hopk_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int flags,
                               void* stream)
{
    HOPK_REQUIRE(M > 0 && N > 0 && K > 0, "linear sizes");
    cudaStream_t st = (cudaStream_t)stream;
    g_tc = (flags & 0x100) != 0;
    Ld2D<true, 0> wl{w, nullptr, K};
    int oflags = (flags & 2) ? 1 : 0;
    if (flags & 1) {
        Ld2D<true, 1> a{x, nullptr, K};
        EpiStore<2> e{y, N, b, nullptr, N, oflags};
        if (M <= 64) launch_gemm2<1, 2>(M, N, K, 1, a, wl, e, st); else launch_gemm2<2, 2>(M, N, K, 1, a, wl, e, st);
    } else {
        Ld2D<true, 0> a{x, nullptr, K};
        EpiStore<2> e{y, N, b, nullptr, N, oflags};
        if (M <= 64) launch_gemm2<1, 2>(M, N, K, 1, a, wl, e, st); else launch_gemm2<2, 2>(M, N, K, 1, a, wl, e, st);
    }
    HOPK_LAUNCH_CHECK("linear_fwd");
    return 0;
}

extern "C" int hopk_linear_bwd(const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                               float* db, int M, int N, int K, int flags, void* stream)
{
    HOPK_REQUIRE(M > 0 && N > 0 && K > 0, "linear sizes");
    HOPK_REQUIRE(!(flags & 2) || y != nullptr, "relu-out backward needs y");
    cudaStream_t st = (cudaStream_t)stream;
    g_tc = (flags & 0x100) != 0;
    const bool relu_in = flags & 1, relu_out = flags & 2;
    // dw[n][k] = sum_m dy_eff[m][n] * x_eff[m][k]  (+ bias gradient as virtual column K)
    if (dw) {
        HOPK_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
        if (db) HOPK_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
        EpiWgrad<2> e{dw, K, db, K, N};
        int sp = splits_for(N, K + 1, M, 2, 2);
        if (g_tc) {                          // MN-major tensor-core weight gradient (gemm_tc_wgrad.cuh)
            W8Plain a{dy, y, N, N, relu_out ? 2 : 0};
            W8Plain bl{x, nullptr, K, K, relu_in ? 1 : 0};
            if (K <= 64) HOPK_CUDA(launch_gemm_tc_wgrad<64>(M, N, K, a, bl, e, db != nullptr, st));
            else HOPK_CUDA(launch_gemm_tc_wgrad<128>(M, N, K, a, bl, e, db != nullptr, st));
        } else if (relu_in) {
            XRelu bl{x, K, K};
            if (relu_out) { Ld2D<false, 2> a{dy, y, N}; launch_gemm2<2, 2>(N, K + 1, M, sp, a, bl, e, st); }
            else { Ld2D<false, 0> a{dy, nullptr, N}; launch_gemm2<2, 2>(N, K + 1, M, sp, a, bl, e, st); }
        } else {
            Ld2DOnes<false> bl{x, K, K};
            if (relu_out) { Ld2D<false, 2> a{dy, y, N}; launch_gemm2<2, 2>(N, K + 1, M, sp, a, bl, e, st); }
            else { Ld2D<false, 0> a{dy, nullptr, N}; launch_gemm2<2, 2>(N, K + 1, M, sp, a, bl, e, st); }
        }
        HOPK_LAUNCH_CHECK("linear_wgrad");
    }
    // dx[m][k] = sum_n dy_eff[m][n] * w[n][k]   (masked by x > 0 when the input was rectified)
    if (dx) {
        Ld2D<false, 0> bl{w, nullptr, K};
        EpiStore<2> e{dx, K, nullptr, x, K, relu_in ? 4 : 0};
        if (relu_out) { Ld2D<true, 2> a{dy, y, N}; launch_gemm2<2, 2>(M, K, N, 1, a, bl, e, st); }
        else { Ld2D<true, 0> a{dy, nullptr, N}; launch_gemm2<2, 2>(M, K, N, 1, a, bl, e, st); }
        HOPK_LAUNCH_CHECK("linear_dgrad");
    }
    return 0;
}

extern "C" int hopk_conv1x1_nchw_fwd(const float* x, const float* w, const float* b, float* y, int B, int K, int N, int V,
                                     int T, void* stream)
{
    HOPK_REQUIRE(B > 0 && K > 0 && N > 0 && V > 0 && T > 0, "conv1x1 sizes");
    cudaStream_t st = (cudaStream_t)stream;
    g_tc = false;
    int VT = V * T, M = B * VT;
    NchwA a{x, K, VT};
    Ld2D<true, 0> wl{w, nullptr, K};
    NchwEpi<2> e{y, b, N, VT};
    launch_gemm2<2, 2>(M, N, K, 1, a, wl, e, st);
    HOPK_LAUNCH_CHECK("conv1x1_fwd");
    return 0;
}

extern "C" int hopk_conv1x1_nchw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int B,
                                     int K, int N, int V, int T, void* stream)
{
    HOPK_REQUIRE(B > 0 && K > 0 && N > 0 && V > 0 && T > 0, "conv1x1 sizes");
    cudaStream_t st = (cudaStream_t)stream;
    g_tc = false;
    int VT = V * T, M = B * VT;
    if (dw) {
        HOPK_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
        if (db) HOPK_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
        NchwATplain a{dy, N, VT};
        NchwAT bl{x, K, VT};
        EpiWgrad<2> e{dw, K, db, K, N};
        launch_gemm2<2, 2>(N, K + 1, M, splits_for(N, K + 1, M, 2, 2), a, bl, e, st);
        HOPK_LAUNCH_CHECK("conv1x1_wgrad");
    }
    if (dx) {
        NchwA a{dy, N, VT};
        Ld2D<false, 0> bl{w, nullptr, K};
        NchwEpi<2> e{dx, nullptr, K, VT};
        launch_gemm2<2, 2>(M, K, N, 1, a, bl, e, st);
        HOPK_LAUNCH_CHECK("conv1x1_dgrad");
    }
    return 0;
}

extern "C" int hopk_nconv_fwd(const float* x, const float* A, float* out, int N, int C, int V, int T, void* stream)
{
    HOPK_REQUIRE(N > 0 && C > 0 && V > 0 && V <= 64 && T > 0, "nconv sizes");
    cudaStream_t st = (cudaStream_t)stream;
    size_t total = (size_t)N * C * V * T;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    nconv_nchw_kernel<<<blocks, 256, V * V * sizeof(float), st>>>(x, A, out, (size_t)N * C, V, T, 0);
    HOPK_LAUNCH_CHECK("nconv_fwd");
    return 0;
}

extern "C" int hopk_nconv_bwd(const float* x, const float* A, const float* dout, float* dx, float* dA, int N, int C, int V,
                              int T, void* stream)
{
    HOPK_REQUIRE(N > 0 && C > 0 && V > 0 && V <= 64 && T > 0, "nconv sizes");
    cudaStream_t st = (cudaStream_t)stream;
    size_t total = (size_t)N * C * V * T;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    if (dx) {
        // dx[n,c,v,l] = sum_w dout[n,c,w,l] A[v,w]  == nconv with A transposed
        nconv_nchw_kernel<<<blocks, 256, V * V * sizeof(float), st>>>(dout, A, dx, (size_t)N * C, V, T, 1);
        HOPK_LAUNCH_CHECK("nconv_dx");
    }
    if (dA) {
        nconv_dA_kernel<<<V * V, 256, 0, st>>>(x, dout, dA, (size_t)N * C, V, T);
        HOPK_LAUNCH_CHECK("nconv_dA");
    }
    return 0;
}
