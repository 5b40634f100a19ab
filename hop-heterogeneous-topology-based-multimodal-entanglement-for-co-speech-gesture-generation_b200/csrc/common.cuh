// common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace hopk {

char* err_buf();                    // thread-local message buffer (api.cu)
constexpr int ERR_BUF = 512;

inline int fail(int code, const char* what, const char* detail = "")
{
    snprintf(err_buf(), ERR_BUF, "hopk: %s %s", what, detail);
    return code;
}

#define HOPK_REQUIRE(cond, msg) \
    do { if (!(cond)) return hopk::fail(2, "bad argument:", msg " [" #cond "]"); } while (0)

#define HOPK_CUDA(expr) \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) return hopk::fail(3, #expr, cudaGetErrorString(_e)); } while (0)

long long& launch_counter();        // kernels launched by this library since load (api.cu)

#define HOPK_LAUNCH_CHECK(name) \
    do { hopk::launch_counter() += 1; cudaError_t _e = cudaGetLastError(); \
         if (_e != cudaSuccess) return hopk::fail(4, "launch failed: " name, cudaGetErrorString(_e)); } while (0)

// device address of the dropout epoch on the current device (api.cu); nullptr on failure
const unsigned long long* drop_epoch_ptr();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) (api.cu)
cudaError_t configure_smem_once(const void* func, size_t bytes);

// dense bf16 GEMM on tcgen05 + TMA (gemm_tma.cu); see hopk_gemm_bf16 in include/hopk.h for the operand conventions
int gemm_bf16_launch(const void* A, const void* B, void* C, const float* bias, const void* addend, int M, int N, int K, long lda,
                     long ldb, long ldc, int a_mn, int b_mn, int out_bf16, int accumulate, int act, float slope, int splits,
                     cudaStream_t st, const void* mask = nullptr, int a_kshift = 0, int b_kshift = 0, int bias_row = 0, int mask_bf16 = 0, int mask_gelu = 0);

// bf16 row-major matrix (rows x cols, pitch ld elements) -> CUtensorMap with {64 columns, box_rows rows} boxes, 128B swizzle
int make_tensor_map_bf16(void* tmap, const void* base, long rows, long cols, long ld, int box_rows);
int make_tensor_map_2d(void* tmap, const void* base, int elt_bytes, long rows, long cols, long ld, int box_cols, int box_rows, int swizzle_bytes);

#ifdef __CUDACC__
// exact-form (erf) GELU and its derivative with one MUFU.EX2 and one MUFU.RCP: erf by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7, the size of fp32 rounding), sharing exp(-x^2 / 2) between the normal CDF and its density.
//   cdf(x) = 0.5 (1 + erf(x / sqrt 2)),  gelu(x) = x cdf(x),  gelu'(x) = cdf(x) + x pdf(x)
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf)
{
    const float ax = fabsf(x);
    float e, t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));            // exp(-x^2 / 2)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.f)));
    const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
    const float half_erfc = 0.5f * poly * e;                                                         // 0.5 erfc(|x| / sqrt 2)
    cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
    pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_fast(float x)
{
    float c, p;
    gelu_cdf_pdf(x, c, p);
    return x * c;
}
__device__ __forceinline__ float gelu_grad_fast(float x)
{
    float c, p;
    gelu_cdf_pdf(x, c, p);
    return fmaf(x, p, c);
}
#endif

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace hopk
