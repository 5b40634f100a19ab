// functors.cuh -- tile loaders and epilogues plugged into hopk::gemm_kernel (see gemm_core.cuh).
//
// Activation layout used by every gwnet kernel ("rows" layout): a tensor that the reference keeps
// as NCHW (B, C, V, T) is stored as rows r = (b*T + t)*V + v with the C channels contiguous.
// Time-major rows make a dilated tap a constant row offset (d*V) and let HOP.Model's glue hand
// its (B, 16, V, 173) buffer to the kernels without a permute copy.
#pragma once
#include <type_traits>
#include "gemm_core.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc_wgrad.cuh"

namespace hopk {

// ---------------------------------------------------------------- generic 2-D operand loaders
// MODE 0: plain, 1: relu(value), 2: value masked by aux > 0
template <bool KFAST, int MODE>
struct Ld2D {
    static constexpr bool kFast = KFAST;
    const float* p; const float* aux; long ld;
    __device__ __forceinline__ float operator()(int i, int k) const {
        long idx = KFAST ? (long)i * ld + k : (long)k * ld + i;
        float v = __ldg(p + idx);
        if (MODE == 1) v = fmaxf(v, 0.f);
        if (MODE == 2) v = (__ldg(aux + idx) > 0.f) ? v : 0.f;
        return v;
    }
    // vector interfaces for the tensor-core skeleton (see gemm_tc.cuh): 8 elements along the contiguous index
    __device__ __forceinline__ void vec8(long base, int c, int cmax, float (&f)[8]) const {
        if (c + 8 <= cmax && ((base | ld | c) & 3) == 0) {
            float4 a = __ldg(reinterpret_cast<const float4*>(p + base + c)), b = __ldg(reinterpret_cast<const float4*>(p + base + c + 4));
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = c + q < cmax ? __ldg(p + base + c + q) : 0.f;
        }
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
        }
        if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = (c + q < cmax && __ldg(aux + base + c + q) > 0.f) ? f[q] : 0.f;
        }
    }
    template <bool KF = KFAST, typename std::enable_if<KF, int>::type = 0>
    __device__ __forceinline__ void ld8(int i, int k, int kmax, float (&f)[8]) const { vec8((long)i * ld, k, kmax, f); }
    template <bool KF = KFAST, typename std::enable_if<!KF, int>::type = 0>
    __device__ __forceinline__ void ld8mn(int k, int i, int imax, float (&f)[8]) const { vec8((long)k * ld, i, imax, f); }
};

// same, plus one virtual all-ones *output* column at index `ones_at`
// (used by weight-gradient GEMMs so the bias gradient falls out as an extra column for free).
template <bool KFAST>
struct Ld2DOnes {
    static constexpr bool kFast = KFAST;
    const float* p; long ld; int ones_at;
    __device__ __forceinline__ float operator()(int i, int k) const {
        if (i == ones_at) return 1.f;
        long idx = KFAST ? (long)i * ld + k : (long)k * ld + i;
        return __ldg(p + idx);
    }
};

// rows-layout <-> (b,t,v) decomposition with arbitrary element strides (NCHW views etc.)
struct RowMap {
    int T, V; long sB, sT, sV, sC;
    __device__ __forceinline__ long off(int m, int c) const {
        int v = m % V; int bt = m / V; int t = bt % T; int b = bt / T;
        return (long)b * sB + (long)t * sT + (long)v * sV + (long)c * sC;
    }
};
template <bool KFAST>
struct LdStrided {          // A(m, k) = x[rowmap(m), k]
    static constexpr bool kFast = KFAST;
    const float* p; RowMap rm;
    __device__ __forceinline__ float operator()(int m, int k) const { return __ldg(p + rm.off(m, k)); }
};
template <bool KFAST>
struct LdStridedT {         // B'(kout, m) = x[rowmap(m), kout], with a virtual ones column at kout == ones_at
    static constexpr bool kFast = KFAST;
    const float* p; RowMap rm; int ones_at;
    __device__ __forceinline__ float operator()(int kout, int m) const {
        if (kout == ones_at) return 1.f;
        return __ldg(p + rm.off(m, kout));
    }
};

// ---------------------------------------------------------------- generic epilogues
// flags: 1 = relu on output, 2 = atomicAdd (split-K partials into a zeroed buffer), 4 = mask by aux>0
template <int NG>
struct EpiStore {
    float* out; long ld; const float* bias; const float* aux; int N; int flags;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * g + j;
                if (n < N) {
                    float x = v[4 * g + j];
                    if (bias) x += __ldg(bias + n);
                    if (flags & 1) x = fmaxf(x, 0.f);
                    long idx = (long)m * ld + n;
                    if (flags & 4) x = (__ldg(aux + idx) > 0.f) ? x : 0.f;
                    if (flags & 2) atomicAdd(out + idx, x); else out[idx] = x;
                }
            }
    }
    __device__ __forceinline__ void flush(int) {}
    // tensor-core skeleton (gemm_tc.cuh): 32 consecutive columns of one row
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        if (!valid) return;
        const long base = (long)m * ld + n0;
        const bool fast = n0 + 32 <= N && (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias) |
                                                              reinterpret_cast<uintptr_t>(aux)) & 15) == 0;
        if (fast) {                                          // 128-bit loads / stores, loads batched ahead of the stores
            float4 bb[8], aa[8];
            if (bias) {
#pragma unroll
                for (int q = 0; q < 8; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(bias + n0) + q);
            }
            if (flags & 4) {
#pragma unroll
                for (int q = 0; q < 8; ++q) aa[q] = __ldg(reinterpret_cast<const float4*>(aux + base) + q);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 x = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                if (bias) { x.x += bb[q].x; x.y += bb[q].y; x.z += bb[q].z; x.w += bb[q].w; }
                if (flags & 1) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                if (flags & 4) {
                    x.x = aa[q].x > 0.f ? x.x : 0.f; x.y = aa[q].y > 0.f ? x.y : 0.f;
                    x.z = aa[q].z > 0.f ? x.z : 0.f; x.w = aa[q].w > 0.f ? x.w : 0.f;
                }
                if (flags & 2) { float* o = out + base + 4 * q; atomicAdd(o, x.x); atomicAdd(o + 1, x.y); atomicAdd(o + 2, x.z); atomicAdd(o + 3, x.w); }
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
            if (!(flags & 2)) {
                if ((reinterpret_cast<uintptr_t>(out + base) & 31) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) tc::stg256(out + base + j, v + j);      // full 32-byte sectors
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + base + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            if (n < N) {
                float x = v[j];
                if (bias) x += __ldg(bias + n);
                if (flags & 1) x = fmaxf(x, 0.f);
                long idx = (long)m * ld + n;
                if (flags & 4) x = (__ldg(aux + idx) > 0.f) ? x : 0.f;
                if (flags & 2) atomicAdd(out + idx, x); else out[idx] = x;
            }
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

template <int NG>
struct EpiStoreStrided {    // out[rowmap(m), n] = v + bias[n]
    float* out; RowMap rm; const float* bias; int N;
    __device__ __forceinline__ void operator()(int m, int nb, const float (&v)[4 * NG]) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = nb + 64 * g + j;
                if (n < N) out[rm.off(m, n)] = v[4 * g + j] + (bias ? __ldg(bias + n) : 0.f);
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int m, bool valid, int n0, float (&v)[32], float*) {
        if (!valid) return;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int n = n0 + j;
            if (n < N) out[rm.off(m, n)] = v[j] + (bias ? __ldg(bias + n) : 0.f);
        }
    }
    __device__ __forceinline__ void finish(float*) {}
};

// weight-gradient epilogue: rows n of dW (ld = K) plus the bias gradient in virtual column K.
template <int NG>
struct EpiWgrad {
    float* dw; long ld; float* db; int K; int Nrows;
    __device__ __forceinline__ void operator()(int n, int kb, const float (&v)[4 * NG]) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int k = kb + 64 * g + j;
                if (k < K) atomicAdd(dw + (long)n * ld + k, v[4 * g + j]);
                else if (k == K && db) atomicAdd(db + n, v[4 * g + j]);
            }
    }
    __device__ __forceinline__ void flush(int) {}
    __device__ __forceinline__ void row32(int n, bool valid, int k0, float (&v)[32], float*) {
        if (!valid) return;
        float* row = dw + (long)n * ld + k0;
        if (k0 + 32 <= K && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {          // 128-bit reductions (sm_90+)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                atomicAdd(reinterpret_cast<float4*>(row) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            int k = k0 + j;
            if (k < K) atomicAdd(dw + (long)n * ld + k, v[j]);
            else if (k == K && db) atomicAdd(db + n, v[j]);
        }
    }
    __device__ __forceinline__ void finish(float*) {}
    __device__ __forceinline__ void bias(int n, float v) { if (db) atomicAdd(db + n, v); }   // gemm_tc_wgrad.cuh
};

}  // namespace hopk
