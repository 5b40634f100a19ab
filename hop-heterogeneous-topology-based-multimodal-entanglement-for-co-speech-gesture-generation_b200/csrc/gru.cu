// gru.cu -- the decoder tail of HOP.Model: 4-layer bidirectional GRU (hidden 350), forward and backward (dtype 1).
//
// Replaces cuDNN's RNN behind `self.gru(dec_out)` (reference model/HOP.py:166-167, 248) and its autograd backward.
// PyTorch GRU cell (gate order r, z, n):
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
//
// Per layer:
//   1. input projections of ALL time steps and both directions as ONE dense GEMM on the TMA + tcgen05 kernel
//      (gemm_tma.cu):  Gi[T*B][2 x 1056] = X[T*B][I] . W_ih^T + b_ih (+ b_hr, b_hz folded in)
//   2. the recurrence as a persistent kernel: a thread-block cluster of 8 CTAs owns one (direction, 16-sample batch slice);
//      CTA c keeps the rows of W_hh of its 44 hidden units (r | z | n: 132 rows x 352, bf16, 102 KB) resident in shared
//      memory for all T steps.  Each step is a tcgen05 MMA  D[gate row][sample] = W_slice . h_{t-1}^T  (M 128, N 16, K 352;
//      accumulator in TMEM), the gate arithmetic on (sample, 4 units) threads with h carried in fp32 registers, and the
//      new h written as bf16 straight into the B-operand buffer of all 8 CTAs through distributed shared memory
//      (st.shared::cluster), ordered by one cluster barrier per step.  16 clusters x 8 CTAs = 128 SMs at batch 128.
// Backward (BPTT) mirrors it: W_hh^T slices resident, per step  dh_{t-1} = dh * z + dGh . W_hh  (M 128, N 16, K 1056) with
// the gate derivatives exchanged through DSMEM; afterwards dW_ih, dW_hh, dX are dense GEMMs over all time steps (the
// h_{t-1} operand of dW_hh is the stored output read through a TMA row shift, out-of-range rows = the zero initial state).
//
// Internal layouts are time-major (row = t*B + b) and padded: hidden 350 -> 352 per direction (704 per row), gates 3 x 352.
#include <mutex>
#include "tc_core.cuh"
#include "common.cuh"
#include "../../include/hopk.h"

namespace hopk {

constexpr int GRU_CL = 8;                      // CTAs per cluster
constexpr int GRU_UNITS = 44;                  // hidden units per CTA
constexpr int GRU_HP = GRU_CL * GRU_UNITS;     // 352: padded hidden size
constexpr int GRU_G = 3 * GRU_HP;              // 1056 gate columns per direction
constexpr int GRU_NB = 16;                     // samples per cluster
constexpr int GRU_THREADS = 288;               // 8 worker warps + 1 MMA-issue warp
constexpr int GRU_GATE_THREADS = GRU_NB * (GRU_UNITS / 4);      // 176: (sample, group of 4 units)

// forward A operand (the CTA's rows [r 44 | z 44 | n 44] of W_hh, K = 352) lives in TENSOR MEMORY for the whole kernel:
// lane = gate row, 32-bit column j = the bf16 pair k = 2j, 2j+1.  Block 0 = rows 0..127, block 1 = rows 128..131 (lanes 0..3).
// Streaming a 128-row A tile from shared memory costs 4 KB per K = 16 step; from TMEM the tensor core only reads the
// 512-byte B tile.
constexpr int GRU_FA_NSLAB = 6;                                             // K = 352 -> 5.5 slabs of 64 (B operand)
constexpr uint32_t GRU_FT_COLS = GRU_HP / 2;                                // 176 columns per block
constexpr uint32_t GRU_FT_A0 = 32, GRU_FT_A1 = GRU_FT_A0 + GRU_FT_COLS;     // D in columns [0, 32), blocks at 32 and 208
constexpr uint32_t GRU_FT_IMG = (128 + 32) * GRU_FT_COLS * 4;               // packed image per (dir, CTA): 160 rows x 176 u32
constexpr uint32_t GRU_B_SLAB = GRU_NB * 128;                               // [16 samples][64 bf16]
constexpr uint32_t GRU_FB_BYTES = GRU_FA_NSLAB * GRU_B_SLAB;                // 12288 per buffer
// backward A operand: rows = the CTA's 44 units k of W_hh^T, K = 1056 gate columns -> 17 slabs of [48 rows][64 bf16]
constexpr int GRU_BA_ROWS = 48;
constexpr uint32_t GRU_BA_SLAB = GRU_BA_ROWS * 128;
constexpr int GRU_BA_NSLAB = 17;
// ... of which K = 0..959 lives in TENSOR MEMORY (480 columns, lanes 0..63) and only the last two K-slabs (k = 960..1087) stay
// in shared memory: 480 + 16 accumulator columns fill the 512-column allocation
constexpr int GRU_BT_K = 960;
constexpr uint32_t GRU_BT_COLS = GRU_BT_K / 2;                              // 480
constexpr uint32_t GRU_BT_A0 = 16;                                          // D in columns [0, 16)
constexpr uint32_t GRU_BT_IMG = 64 * GRU_BT_COLS * 4;                       // 122880: image of 64 rows x 480 u32
constexpr int GRU_BS_SLABS = 2;
constexpr uint32_t GRU_BS_BYTES = 16384;                                    // 2 slabs of [48][64] (12288) + in-range room for the M = 128 read
constexpr uint32_t GRU_BB_BYTES = GRU_BA_NSLAB * GRU_B_SLAB;                // 34816 per buffer
constexpr uint32_t GRU_F_TX = GRU_NB * GRU_HP * 2;                           // bytes of h_t a CTA receives per step (11264)
constexpr uint32_t GRU_B_TX = GRU_NB * GRU_G * 2;                            // bytes of dGh a CTA receives per step (33792)
constexpr int GRU_XLD = GRU_NB + 1;                                         // fp32 pitch of the TMEM -> thread exchange tile


// ---------------------------------------------------------------- cluster primitives
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// asynchronous remote store: 8 bytes into a peer CTA's shared memory, completion counted (complete_tx) on THAT CTA's mbarrier.
// The sender neither fences nor waits; the receiver's MMA issuer waits on its own barrier for the expected byte count.
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t a, uint32_t b, uint32_t mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "r"(a), "r"(b), "r"(mbar) : "memory");
}
// generic-proxy writes (own and remote shared memory) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float tanh_approx(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// byte offset of element (sample row gb, contraction index k) inside a stack of [16][64] K-major slabs; k % 4 == 0
__device__ __forceinline__ uint32_t bop_off(int gb, int k)
{
    return (uint32_t)(k >> 6) * GRU_B_SLAB + (uint32_t)gb * 128u + (uint32_t)((((k & 63) >> 3) ^ (gb & 7)) << 4) + (uint32_t)(k & 7) * 2u;
}

// ---------------------------------------------------------------- warp roles and CTA-level synchronisation
// warps 0-5  gate warps (176 gate threads = (sample, group of 4 units); warps 0-3 also drain TMEM): on-chip work only
// warps 6-7  I/O warps: every global load / store of the step, through shared-memory staging, so that the threads that
//            execute the cluster-scope fences (MEMBAR.ALL.GPU in SASS) never have global traffic in flight
// warp 8     MMA issuer
constexpr int GRU_IO_T0 = 192, GRU_IO_THREADS = 64;
constexpr int GRU_BAR_IN = 1, GRU_BAR_OUT = 2;          // named barriers: step inputs staged / step results staged (256 threads)
constexpr int GRU_ROW4 = GRU_UNITS / 4;                 // 11 groups of 4 units per sample row
constexpr int GRU_TILE4 = GRU_NB * GRU_ROW4;            // 176 float4 per staged array

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_cluster() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// K-major descriptor of a slab at `addr`; advancing the start address by `bytes` adds bytes >> 4 to the low word
__device__ __forceinline__ uint64_t desc_base(uint32_t addr) { return tc::desc_kmajor(addr, 0); }
__device__ __forceinline__ uint64_t desc_adv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// Tensor-memory image of an A operand: [chunk of 16 columns][row][16 x u32], so that a warp's tcgen05.st of one chunk reads
// 32 rows x 64 contiguous bytes.  rows = TMEM lanes written (a multiple of 32), cols = 32-bit columns (a multiple of 16).
__device__ __forceinline__ void tmem_load_image(uint32_t tmem_dst, const uint32_t* __restrict__ img, int rows, int cols, int row)
{
#pragma unroll 1
    for (int c = 0; c < cols / 16; ++c) {
        const uint4* src = reinterpret_cast<const uint4*>(img + ((size_t)c * rows + row) * 16);
        uint32_t v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 u = __ldg(src + q);
            v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
        }
        tc::tmem_st16(tmem_dst + c * 16, v);
    }
}

// ---------------------------------------------------------------- forward recurrence
struct GruFwdArgs {
    const float* Gi;            // [T*B][2*1056]: input projections + b_ih (+ b_hr, b_hz)
    const uint8_t* whh;         // [2 dirs][8 CTAs][GRU_FT_IMG] packed W_hh slices (tensor-memory images)
    const float* bhn;           // [2][352]
    __nv_bfloat16* Y;           // [T*B][704] outputs h_t (bf16): operand of the next layer's GEMM / of backward
    float *R, *Z, *N, *HN;      // [T*B][704] saved gate values (nullable: inference / no-grad pass)
    float* out;                 // optional (B, T, 2H) batch-first fp32 output (last layer)
    int B, T, H;
};

constexpr size_t GRU_F_XCHG = 132 * GRU_XLD * 4;                           // 8976
constexpr size_t GRU_F_SIN = 3 * GRU_TILE4 * 16;                           // staged gi_r | gi_z | gi_n       8448
constexpr size_t GRU_F_SOUT = 5 * GRU_TILE4 * 16;                          // staged h | r | z | n | hn      14080
// one CTA per SM is REQUIRED (each CTA allocates all 512 tensor-memory columns; a second resident CTA of the same cluster would
// wait for them forever): the request is padded past half of the SM's shared memory
constexpr size_t GRU_SMEM_FLOOR = 120 * 1024;
constexpr size_t gru_fwd_smem() { return GRU_SMEM_FLOOR; }
static_assert(2 * GRU_FB_BYTES + GRU_F_XCHG + 16 + GRU_F_SIN + GRU_F_SOUT + 1024 <= GRU_SMEM_FLOOR, "forward smem");

__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(GRU_THREADS, 1) gru_fwd_kernel(const GruFwdArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ __align__(16) uint64_t hbar[2];        // operand buffer b complete: 8 CTAs x 176 threads x 8 bytes of st.async
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;
    float* xchg = reinterpret_cast<float*>(sB + 2 * GRU_FB_BYTES);
    float4* sin4 = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(xchg) + ((GRU_F_XCHG + 15) & ~size_t(15)));
    float4* sout4 = sin4 + 3 * GRU_TILE4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cl = blockIdx.x / GRU_CL;
    const int dir = cl & 1, b0 = (cl >> 1) * GRU_NB;
    const int B = a.B, T = a.T;

    if (tid == 0) {
        tc::mbar_init(&mbar, 1);
        tc::mbar_init(&hbar[0], 1);
        tc::mbar_init(&hbar[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&hbar[0], GRU_F_TX);       // armed for their first fill (steps 2 and 1)
        tc::mbar_expect_tx(&hbar[1], GRU_F_TX);
    }
    if (warp == 8) tc::tmem_alloc(&tmem_slot, 512);
    for (int i = tid; i < (int)(2 * GRU_FB_BYTES / 16); i += GRU_THREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);   // h_{-1} = 0
    fence_proxy_async_all();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp < 4) {
        // W_hh slice -> tensor memory: thread = one gate row (TMEM lane), 176 packed bf16 pairs per row
        const uint32_t* img = reinterpret_cast<const uint32_t*>(a.whh + (size_t)(dir * GRU_CL + rank) * GRU_FT_IMG);
        tmem_load_image(tmem + ((uint32_t)(warp * 32) << 16) + GRU_FT_A0, img, 128, GRU_FT_COLS, warp * 32 + lane);
        if (warp == 0) tmem_load_image(tmem + GRU_FT_A1, img + 128 * GRU_FT_COLS, 32, GRU_FT_COLS, lane);
        tc::tmem_wait_st();
        tc::fence_before_sync();
    }
    __syncthreads();
    tc::fence_after_sync();
    cluster_arrive();
    cluster_wait();                                   // every CTA's operand buffers are initialised before a peer writes into them

    if (warp == 8) {
        // ============================================================ MMA issuer
        const uint64_t dB0 = desc_base(tc::smem_u32(sB)), dB1 = desc_base(tc::smem_u32(sB) + GRU_FB_BYTES);
        constexpr uint32_t idesc = tc::idesc_bf16(128, GRU_NB, 0, 0);
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            if (step > 0) {                           // h_{t-1}: every CTA's slice has landed in buffer step & 1
                uint64_t* hb = &hbar[step & 1];
                tc::mbar_wait(hb, (uint32_t)((step - 1) >> 1) & 1u);
                if (lane == 0 && step + 2 < T) tc::mbar_expect_tx(hb, GRU_F_TX);      // re-arm for its next fill
                tc::fence_async_smem();               // the st.async data -> async proxy (operand reads of the MMAs below)
            }
            tc::fence_after_sync();
            if (elect_one()) {
                const uint64_t dB = (step & 1) ? dB1 : dB0;
#pragma unroll
                for (int kt = 0; kt < GRU_HP / 16; ++kt) {
                    const uint64_t db = desc_adv(dB, (kt >> 2) * GRU_B_SLAB + (kt & 3) * 32);
                    tc::mma_bf16_ts(tmem, tmem + GRU_FT_A0 + kt * 8, db, idesc, kt != 0);        // gate rows 0..127
                    tc::mma_bf16_ts(tmem + 16, tmem + GRU_FT_A1 + kt * 8, db, idesc, kt != 0);   // gate rows 128..131
                }
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ============================================================ I/O warps: global <-> staging
        const int it = tid - GRU_IO_T0;
        float4 g[3][3];                               // up to 3 float4 per thread per array (176 float4 over 64 threads)
        auto load_gi = [&](int t) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                const int gb = i / GRU_ROW4, grp = i - gb * GRU_ROW4;
                const bool ok = i < GRU_TILE4 && b0 + gb < B;
                const float* p = a.Gi + ((size_t)t * B + (b0 + gb)) * (2 * GRU_G) + dir * GRU_G + (int)rank * GRU_UNITS + grp * 4;
#pragma unroll
                for (int q = 0; q < 3; ++q) g[q][k] = ok ? __ldg(reinterpret_cast<const float4*>(p + q * GRU_HP)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto stage_gi = [&]() {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                if (i < GRU_TILE4) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) sin4[q * GRU_TILE4 + i] = g[q][k];
                }
            }
        };
        load_gi(dir ? T - 1 : 0);
        stage_gi();
        bar_arrive(GRU_BAR_IN, 256);
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            const int t = dir ? T - 1 - step : step;
            if (step + 1 < T) load_gi(dir ? t - 1 : t + 1);                  // in flight while the gate warps work
            bar_sync(GRU_BAR_OUT, 256);                                       // results of this step staged
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                const int gb = i / GRU_ROW4, grp = i - gb * GRU_ROW4;
                if (i < GRU_TILE4 && b0 + gb < B) {
                    const int b = b0 + gb, ju = (int)rank * GRU_UNITS + grp * 4;
                    const size_t o = ((size_t)t * B + b) * (2 * GRU_HP) + dir * GRU_HP + ju;
                    const float4 h = sout4[i];
                    uint2 hb;
                    hb.x = tc::pack_bf16x2(h.x, h.y); hb.y = tc::pack_bf16x2(h.z, h.w);
                    *reinterpret_cast<uint2*>(a.Y + o) = hb;
                    if (a.R) {
                        *reinterpret_cast<float4*>(a.R + o) = sout4[GRU_TILE4 + i];
                        *reinterpret_cast<float4*>(a.Z + o) = sout4[2 * GRU_TILE4 + i];
                        *reinterpret_cast<float4*>(a.N + o) = sout4[3 * GRU_TILE4 + i];
                        *reinterpret_cast<float4*>(a.HN + o) = sout4[4 * GRU_TILE4 + i];
                    }
                    if (a.out) {
                        float* po = a.out + ((size_t)b * T + t) * (2 * a.H) + dir * a.H + ju;
                        const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) if (ju + e < a.H) po[e] = hv[e];
                    }
                }
            }
            if (step + 1 < T) stage_gi();
            bar_arrive(GRU_BAR_IN, 256);                                      // inputs of step + 1 staged (and the results tile is free)
        }
    } else {
        // ============================================================ gate warps
        const bool gate_thr = tid < GRU_GATE_THREADS;
        const int gb = tid & 15, grp = tid >> 4;
        const bool act = gate_thr && b0 + gb < B;
        const int ju = (int)rank * GRU_UNITS + grp * 4;   // first of this thread's 4 hidden units (padded index)
        const int si = gb * GRU_ROW4 + grp;               // this thread's float4 slot in a staged array
        float h[4] = {0.f, 0.f, 0.f, 0.f};
        float4 bhn = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t raddr[GRU_CL], rbar[GRU_CL];
        if (gate_thr) {
            const uint32_t local = tc::smem_u32(sB) + bop_off(gb, ju), lbar = tc::smem_u32(&hbar[0]);
#pragma unroll
            for (int r = 0; r < GRU_CL; ++r) { raddr[r] = mapa(local, (uint32_t)r); rbar[r] = mapa(lbar, (uint32_t)r); }
            bhn = __ldg(reinterpret_cast<const float4*>(a.bhn + dir * GRU_HP + ju));
        }
        uint32_t mphase = 0;
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            const int cur = step & 1;
            tc::mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc::fence_after_sync();
            if (warp < 4) {
                float v[16];
                tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
                float* dst = xchg + (warp * 32 + lane) * GRU_XLD;
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[j] = v[j];
                if (warp == 0) {
                    tc::tmem_ld16(tmem + 16, v);
                    if (lane < 4) {
                        float* d2 = xchg + (128 + lane) * GRU_XLD;
#pragma unroll
                        for (int j = 0; j < 16; ++j) d2[j] = v[j];
                    }
                }
                tc::fence_before_sync();
            }
            bar_sync(GRU_BAR_IN, 256);                    // exchange tile complete + this step's input projections staged
            if (gate_thr) {
                const float4 g_r = sin4[si], g_z = sin4[GRU_TILE4 + si], g_n = sin4[2 * GRU_TILE4 + si];
                const float gir[4] = {g_r.x, g_r.y, g_r.z, g_r.w}, giz[4] = {g_z.x, g_z.y, g_z.z, g_z.w};
                const float gin[4] = {g_n.x, g_n.y, g_n.z, g_n.w}, bh[4] = {bhn.x, bhn.y, bhn.z, bhn.w};
                float rr[4], zz[4], nn[4], hn[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float gr = xchg[(grp * 4 + i) * GRU_XLD + gb];
                    const float gz = xchg[(GRU_UNITS + grp * 4 + i) * GRU_XLD + gb];
                    const float gn = xchg[(2 * GRU_UNITS + grp * 4 + i) * GRU_XLD + gb];
                    rr[i] = sigmoid_approx(gir[i] + gr);
                    zz[i] = sigmoid_approx(giz[i] + gz);
                    hn[i] = gn + bh[i];
                    nn[i] = tanh_approx(fmaf(rr[i], hn[i], gin[i]));
                    h[i] = act ? fmaf(zz[i], h[i] - nn[i], nn[i]) : 0.f;            // (1 - z) n + z h
                }
                if (step + 1 < T) {                                                 // h_t into every CTA's operand buffer of step + 1
                    const uint32_t lo = tc::pack_bf16x2(h[0], h[1]), hi = tc::pack_bf16x2(h[2], h[3]);
                    const uint32_t boff = (cur ^ 1) * GRU_FB_BYTES, moff = (cur ^ 1) * 8;
#pragma unroll
                    for (int r = 0; r < GRU_CL; ++r) st_async_v2(raddr[r] + boff, lo, hi, rbar[r] + moff);
                }
                sout4[si] = make_float4(h[0], h[1], h[2], h[3]);
                sout4[GRU_TILE4 + si] = make_float4(rr[0], rr[1], rr[2], rr[3]);
                sout4[2 * GRU_TILE4 + si] = make_float4(zz[0], zz[1], zz[2], zz[3]);
                sout4[3 * GRU_TILE4 + si] = make_float4(nn[0], nn[1], nn[2], nn[3]);
                sout4[4 * GRU_TILE4 + si] = make_float4(hn[0], hn[1], hn[2], hn[3]);
            }
            __syncwarp();
            bar_arrive(GRU_BAR_OUT, 256);                 // the I/O warps take it from here
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    cluster_arrive();                                     // nobody leaves while a peer may still address its shared memory
    cluster_wait();
    if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- backward recurrence (BPTT)
struct GruBwdArgs {
    const float* dY;             // [T*B][704] gradient w.r.t. this layer's outputs
    const uint8_t* whhT;         // [2][8][GRU_BT_IMG + GRU_BS_BYTES] packed W_hh^T slices (tensor-memory image | smem tail slabs)
    const __nv_bfloat16* Y;      // [T*B][704] forward outputs
    const float *R, *Z, *N, *HN;
    __nv_bfloat16* dGi;          // [T*B][2*1056] gradient w.r.t. the input projections
    __nv_bfloat16* dGh;          // [T*B][2*1056] gradient w.r.t. the hidden projections (differs in the n gate: x r)
    int B, T;
};

constexpr size_t GRU_B_XCHG = 64 * GRU_XLD * 4;                            // 4352
constexpr size_t GRU_B_SIN = 5 * GRU_TILE4 * 16 + GRU_TILE4 * 8;           // staged r | z | n | hn | dy (fp32) + h_prev (bf16 x 4)
constexpr size_t GRU_B_SOUT = 4 * GRU_TILE4 * 8;                           // staged dg_r | dg_z | dg_n(input) | dg_n(hidden), bf16 x 4
constexpr size_t gru_bwd_smem() { return GRU_SMEM_FLOOR; }
static_assert(GRU_BS_BYTES + 2 * GRU_BB_BYTES + GRU_B_XCHG + GRU_B_SIN + GRU_B_SOUT + 1024 <= GRU_SMEM_FLOOR, "backward smem");

__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(GRU_THREADS, 1) gru_bwd_kernel(const GruBwdArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t wbar, mbar;
    __shared__ __align__(16) uint64_t hbar[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                   // the last GRU_BS_SLABS K-slabs of W_hh^T (the rest lives in TMEM)
    uint8_t* sB = smem + GRU_BS_BYTES;
    float* xchg = reinterpret_cast<float*>(sB + 2 * GRU_BB_BYTES);
    float4* sin4 = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(xchg) + GRU_B_XCHG);
    uint2* shp = reinterpret_cast<uint2*>(sin4 + 5 * GRU_TILE4);
    uint2* sout2 = shp + GRU_TILE4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cl = blockIdx.x / GRU_CL;
    const int dir = cl & 1, b0 = (cl >> 1) * GRU_NB;
    const int B = a.B, T = a.T;
    const uint8_t* wimg = a.whhT + (size_t)(dir * GRU_CL + rank) * (GRU_BT_IMG + GRU_BS_BYTES);

    if (tid == 0) {
        tc::mbar_init(&wbar, 1);
        tc::mbar_init(&mbar, 1);
        tc::mbar_init(&hbar[0], 1);
        tc::mbar_init(&hbar[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&hbar[0], GRU_B_TX);
        tc::mbar_expect_tx(&hbar[1], GRU_B_TX);
        tc::mbar_expect_tx(&wbar, GRU_BS_BYTES);
        tc::bulk_g2s(sA, wimg + GRU_BT_IMG, GRU_BS_BYTES, &wbar);
    }
    if (warp == 8) tc::tmem_alloc(&tmem_slot, 512);
    for (int i = tid; i < (int)(2 * GRU_BB_BYTES / 16); i += GRU_THREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_all();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp < 2) {                                       // rows 0..63 of W_hh^T (44 valid), K = 0..959 -> 480 TMEM columns
        tmem_load_image(tmem + ((uint32_t)(warp * 32) << 16) + GRU_BT_A0, reinterpret_cast<const uint32_t*>(wimg), 64, GRU_BT_COLS, warp * 32 + lane);
        tc::tmem_wait_st();
        tc::fence_before_sync();
    }
    __syncthreads();
    tc::fence_after_sync();
    cluster_arrive();
    cluster_wait();

    if (warp == 8) {
        // ============================================================ MMA issuer: (dGh of the step before) . W_hh -> D[unit k][sample]
        const uint64_t dA = desc_base(tc::smem_u32(sA));
        const uint64_t dB0 = desc_base(tc::smem_u32(sB)), dB1 = desc_base(tc::smem_u32(sB) + GRU_BB_BYTES);
        constexpr uint32_t idesc = tc::idesc_bf16(128, GRU_NB, 0, 0);
        tc::mbar_wait(&wbar, 0);
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            if (step > 0) {
                uint64_t* hb = &hbar[step & 1];
                tc::mbar_wait(hb, (uint32_t)((step - 1) >> 1) & 1u);
                if (lane == 0 && step + 2 < T) tc::mbar_expect_tx(hb, GRU_B_TX);
                tc::fence_async_smem();
            }
            tc::fence_after_sync();
            if (elect_one()) {
                const uint64_t dB = (step & 1) ? dB1 : dB0;
#pragma unroll
                for (int kt = 0; kt < GRU_BT_K / 16; ++kt)                             // A from tensor memory
                    tc::mma_bf16_ts(tmem, tmem + GRU_BT_A0 + kt * 8, desc_adv(dB, (kt >> 2) * GRU_B_SLAB + (kt & 3) * 32), idesc, kt != 0);
#pragma unroll
                for (int kt = GRU_BT_K / 16; kt < GRU_G / 16; ++kt) {                  // the K tail from shared memory
                    const int ks = kt - GRU_BT_K / 16;
                    tc::mma_bf16(tmem, desc_adv(dA, (ks >> 2) * GRU_BA_SLAB + (ks & 3) * 32), desc_adv(dB, (kt >> 2) * GRU_B_SLAB + (kt & 3) * 32),
                                 idesc, true);
                }
                tc::mma_commit(&mbar);
            }
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ============================================================ I/O warps
        const int it = tid - GRU_IO_T0;
        float4 g[5][3];
        uint2 hp[3];
        auto load_in = [&](int t) {
            const int tp = dir ? t + 1 : t - 1;          // the time step whose output was this step's h_{t-1}
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                const int gb = i / GRU_ROW4, grp = i - gb * GRU_ROW4;
                const bool ok = i < GRU_TILE4 && b0 + gb < B;
                const size_t col = dir * GRU_HP + (int)rank * GRU_UNITS + grp * 4;
                const size_t o = ((size_t)t * B + (b0 + gb)) * (2 * GRU_HP) + col;
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                g[0][k] = ok ? __ldg(reinterpret_cast<const float4*>(a.R + o)) : zero;
                g[1][k] = ok ? __ldg(reinterpret_cast<const float4*>(a.Z + o)) : zero;
                g[2][k] = ok ? __ldg(reinterpret_cast<const float4*>(a.N + o)) : zero;
                g[3][k] = ok ? __ldg(reinterpret_cast<const float4*>(a.HN + o)) : zero;
                g[4][k] = ok ? __ldg(reinterpret_cast<const float4*>(a.dY + o)) : zero;
                hp[k] = (ok && tp >= 0 && tp < T) ? __ldg(reinterpret_cast<const uint2*>(a.Y + ((size_t)tp * B + (b0 + gb)) * (2 * GRU_HP) + col))
                                                  : make_uint2(0u, 0u);
            }
        };
        auto stage_in = [&]() {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                if (i < GRU_TILE4) {
#pragma unroll
                    for (int q = 0; q < 5; ++q) sin4[q * GRU_TILE4 + i] = g[q][k];
                    shp[i] = hp[k];
                }
            }
        };
        load_in(dir ? 0 : T - 1);
        stage_in();
        bar_arrive(GRU_BAR_IN, 256);
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            const int t = dir ? step : T - 1 - step;     // reverse of the forward order
            if (step + 1 < T) load_in(dir ? t + 1 : t - 1);
            bar_sync(GRU_BAR_OUT, 256);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int i = it + k * GRU_IO_THREADS;
                const int gb = i / GRU_ROW4, grp = i - gb * GRU_ROW4;
                if (i < GRU_TILE4 && b0 + gb < B) {
                    const size_t o = ((size_t)t * B + (b0 + gb)) * (2 * GRU_G) + dir * GRU_G + (int)rank * GRU_UNITS + grp * 4;
                    const uint2 vr = sout2[i], vz = sout2[GRU_TILE4 + i], vni = sout2[2 * GRU_TILE4 + i], vnh = sout2[3 * GRU_TILE4 + i];
                    *reinterpret_cast<uint2*>(a.dGi + o) = vr; *reinterpret_cast<uint2*>(a.dGh + o) = vr;
                    *reinterpret_cast<uint2*>(a.dGi + o + GRU_HP) = vz; *reinterpret_cast<uint2*>(a.dGh + o + GRU_HP) = vz;
                    *reinterpret_cast<uint2*>(a.dGi + o + 2 * GRU_HP) = vni;
                    *reinterpret_cast<uint2*>(a.dGh + o + 2 * GRU_HP) = vnh;
                }
            }
            if (step + 1 < T) stage_in();
            bar_arrive(GRU_BAR_IN, 256);
        }
    } else {
        // ============================================================ gate warps
        const bool gate_thr = tid < GRU_GATE_THREADS;
        const int gb = tid & 15, grp = tid >> 4;
        const bool act = gate_thr && b0 + gb < B;
        const int ju = (int)rank * GRU_UNITS + grp * 4;
        const int si = gb * GRU_ROW4 + grp;
        uint32_t loff[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) loff[q] = tc::smem_u32(sB) + bop_off(gb, q * GRU_HP + ju);
        float carry[4] = {0.f, 0.f, 0.f, 0.f};             // dh * z of the step before (direct path to h_{t-1})
        uint32_t mphase = 0;
#pragma unroll 1
        for (int step = 0; step < T; ++step) {
            const int cur = step & 1;
            tc::mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc::fence_after_sync();
            if (warp < 2) {
                float v[16];
                tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
                float* dst = xchg + (warp * 32 + lane) * GRU_XLD;
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[j] = v[j];
                tc::fence_before_sync();
            }
            bar_sync(GRU_BAR_IN, 256);
            if (gate_thr) {
                const float4 r4 = sin4[si], z4 = sin4[GRU_TILE4 + si], n4 = sin4[2 * GRU_TILE4 + si], hn4 = sin4[3 * GRU_TILE4 + si],
                             dy4 = sin4[4 * GRU_TILE4 + si];
                const uint2 hp = shp[si];
                const float r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w}, n[4] = {n4.x, n4.y, n4.z, n4.w};
                const float hn[4] = {hn4.x, hn4.y, hn4.z, hn4.w}, dy[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
                const __nv_bfloat162 h01 = *reinterpret_cast<const __nv_bfloat162*>(&hp.x), h23 = *reinterpret_cast<const __nv_bfloat162*>(&hp.y);
                const float hprev[4] = {__low2float(h01), __high2float(h01), __low2float(h23), __high2float(h23)};
                float dgr[4], dgz[4], dgni[4], dgnh[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dh = act ? dy[i] + carry[i] + xchg[(grp * 4 + i) * GRU_XLD + gb] : 0.f;
                    const float dn = dh * (1.f - z[i]);
                    const float dz = dh * (hprev[i] - n[i]);
                    const float dnp = dn * (1.f - n[i] * n[i]);
                    dgni[i] = dnp;
                    dgnh[i] = dnp * r[i];
                    dgr[i] = dnp * hn[i] * r[i] * (1.f - r[i]);
                    dgz[i] = dz * z[i] * (1.f - z[i]);
                    carry[i] = dh * z[i];
                }
                const uint2 vr = make_uint2(tc::pack_bf16x2(dgr[0], dgr[1]), tc::pack_bf16x2(dgr[2], dgr[3]));
                const uint2 vz = make_uint2(tc::pack_bf16x2(dgz[0], dgz[1]), tc::pack_bf16x2(dgz[2], dgz[3]));
                const uint2 vnh = make_uint2(tc::pack_bf16x2(dgnh[0], dgnh[1]), tc::pack_bf16x2(dgnh[2], dgnh[3]));
                if (step + 1 < T) {
                    const uint32_t boff = (cur ^ 1) * GRU_BB_BYTES, lbar = tc::smem_u32(&hbar[cur ^ 1]);
#pragma unroll
                    for (int rk = 0; rk < GRU_CL; ++rk) {
                        const uint32_t rb = mapa(lbar, (uint32_t)rk);
                        st_async_v2(mapa(loff[0], (uint32_t)rk) + boff, vr.x, vr.y, rb);
                        st_async_v2(mapa(loff[1], (uint32_t)rk) + boff, vz.x, vz.y, rb);
                        st_async_v2(mapa(loff[2], (uint32_t)rk) + boff, vnh.x, vnh.y, rb);
                    }
                }
                sout2[si] = vr;
                sout2[GRU_TILE4 + si] = vz;
                sout2[2 * GRU_TILE4 + si] = make_uint2(tc::pack_bf16x2(dgni[0], dgni[1]), tc::pack_bf16x2(dgni[2], dgni[3]));
                sout2[3 * GRU_TILE4 + si] = vnh;
            }
            __syncwarp();
            bar_arrive(GRU_BAR_OUT, 256);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- packing / unpacking
// padded input width of layer l: layer 0 = I rounded up to 8; deeper layers = 704 ([fwd 350, 0, 0 | bwd 350, 0, 0])
__host__ __device__ inline int gru_in_pad(int l, int I) { return l == 0 ? ((I + 7) & ~7) : 2 * GRU_HP; }
// column of the padded input that holds PyTorch input column c
__host__ __device__ inline int gru_in_col(int l, int H, int c) { return l == 0 ? c : (c < H ? c : GRU_HP + (c - H)); }

// W_ih of both directions -> bf16 [2*1056][Ipad] (row = dir*1056 + gate*352 + unit); bias_comb[2*1056]; bhn[2*352]
__global__ void gru_pack_wih_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ bi0,
                                    const float* __restrict__ bi1, const float* __restrict__ bh0, const float* __restrict__ bh1,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bias, float* __restrict__ bhn, int l, int I, int Iin, int H)
{
    const int Ipad = gru_in_pad(l, I);
    const int row = blockIdx.x;                       // 0 .. 2*1056-1
    const int dir = row / GRU_G, q = (row % GRU_G) / GRU_HP, j = row % GRU_HP;
    const float* w = dir ? w1 : w0;
    const bool valid = j < H;
    __nv_bfloat16* dst = wp + (size_t)row * Ipad;
    for (int c = threadIdx.x; c < Ipad; c += blockDim.x) dst[c] = __float2bfloat16_rn(0.f);
    __syncthreads();
    if (valid) {
        const float* src = w + (size_t)(q * H + j) * Iin;
        for (int c = threadIdx.x; c < Iin; c += blockDim.x) dst[gru_in_col(l, H, c)] = __float2bfloat16_rn(src[c]);
    }
    if (threadIdx.x == 0) {
        const float* bi = dir ? bi1 : bi0;
        const float* bh = dir ? bh1 : bh0;
        float v = 0.f;
        if (valid) v = bi[q * H + j] + (q < 2 ? bh[q * H + j] : 0.f);
        bias[row] = v;
        if (q == 2) bhn[dir * GRU_HP + j] = valid ? bh[2 * H + j] : 0.f;
    }
}

// W_hh -> per (dir, CTA) operand images: forward = tensor-memory image of the rows [r|z|n] of the CTA's units; backward =
// shared-memory slabs [unit k][gate column] of W_hh^T
__global__ void gru_pack_whh_kernel(const float* __restrict__ w0, const float* __restrict__ w1, uint8_t* __restrict__ fimg,
                                    uint8_t* __restrict__ bimg, int H)
{
    const int dir = blockIdx.y, c = blockIdx.x;
    const float* w = dir ? w1 : w0;
    const int t0 = blockIdx.z * blockDim.x + threadIdx.x, tstride = gridDim.z * blockDim.x;
    uint32_t* fi = reinterpret_cast<uint32_t*>(fimg + (size_t)(dir * GRU_CL + c) * GRU_FT_IMG);
    for (int idx = t0; idx < 160 * (int)GRU_FT_COLS; idx += tstride) {       // tensor-memory image, see tmem_load_image
        const int row = idx / (int)GRU_FT_COLS, kp = idx % (int)GRU_FT_COLS;
        float v0 = 0.f, v1 = 0.f;
        if (row < 3 * GRU_UNITS) {
            const int q = row / GRU_UNITS, j = c * GRU_UNITS + row % GRU_UNITS;
            if (j < H) {
                const float* wr = w + (size_t)(q * H + j) * H;
                if (2 * kp < H) v0 = wr[2 * kp];
                if (2 * kp + 1 < H) v1 = wr[2 * kp + 1];
            }
        }
        const int ch = kp >> 4, wd = kp & 15;
        const size_t at = row < 128 ? ((size_t)ch * 128 + row) * 16 + wd : (size_t)128 * GRU_FT_COLS + ((size_t)ch * 32 + (row - 128)) * 16 + wd;
        fi[at] = tc::pack_bf16x2(v0, v1);
    }
    if (!bimg) return;
    uint8_t* bi = bimg + (size_t)(dir * GRU_CL + c) * (GRU_BT_IMG + GRU_BS_BYTES);
    uint32_t* bt = reinterpret_cast<uint32_t*>(bi);
    auto wt = [&](int row, int g) -> float {               // W_hh^T element: unit k = c*44 + row, gate column g
        if (row >= GRU_UNITS || g >= GRU_G) return 0.f;
        const int k = c * GRU_UNITS + row, q = g / GRU_HP, u = g % GRU_HP;
        return (k < H && u < H) ? w[(size_t)(q * H + u) * H + k] : 0.f;
    };
    for (int idx = t0; idx < 64 * (int)GRU_BT_COLS; idx += tstride) {
        const int row = idx / (int)GRU_BT_COLS, kp = idx % (int)GRU_BT_COLS;
        bt[((size_t)(kp >> 4) * 64 + row) * 16 + (kp & 15)] = tc::pack_bf16x2(wt(row, 2 * kp), wt(row, 2 * kp + 1));
    }
    uint8_t* bs = bi + GRU_BT_IMG;
    for (int idx = t0; idx < (int)GRU_BS_BYTES / 2; idx += tstride) {
        float v = 0.f;
        const int e = idx;                                 // bf16 element index inside the 16 KB tail region
        const int slab = e / (GRU_BA_ROWS * 64);
        if (slab < GRU_BS_SLABS) {
            const int row = (e % (GRU_BA_ROWS * 64)) / 64, col = e % 64;
            v = wt(row, GRU_BT_K + slab * 64 + col);
            const uint32_t off = (uint32_t)slab * GRU_BA_SLAB + tc::slab_chunk_off(row, col >> 3) + (uint32_t)(col & 7) * 2u;
            *reinterpret_cast<__nv_bfloat16*>(bs + off) = __float2bfloat16_rn(v);
        } else {
            *reinterpret_cast<__nv_bfloat16*>(bs + (size_t)e * 2) = __float2bfloat16_rn(0.f);
        }
    }
}

// x (B, T, I) fp32 batch-first -> bf16 [T*B][Ipad] time-major
__global__ void gru_pack_x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xp, int B, int T, int I, int Ipad)
{
    const size_t n = (size_t)B * T * Ipad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ipad);
        const size_t r = i / Ipad;
        const int b = (int)(r % B), t = (int)(r / B);
        xp[i] = __float2bfloat16_rn(c < I ? x[((size_t)b * T + t) * I + c] : 0.f);
    }
}
// dout (B, T, 2H) fp32 batch-first -> fp32 [T*B][704] time-major (pad columns zero)
__global__ void gru_pack_dy_kernel(const float* __restrict__ dy, float* __restrict__ dp, int B, int T, int H)
{
    const size_t n = (size_t)B * T * 2 * GRU_HP;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % (2 * GRU_HP));
        const size_t r = i / (2 * GRU_HP);
        const int b = (int)(r % B), t = (int)(r / B);
        const int dir = c / GRU_HP, j = c % GRU_HP;
        dp[i] = j < H ? dy[((size_t)b * T + t) * (2 * H) + dir * H + j] : 0.f;
    }
}
// dX [T*B][Ipad] fp32 time-major -> (B, T, I) batch-first
__global__ void gru_unpack_dx_kernel(const float* __restrict__ dxp, float* __restrict__ dx, int B, int T, int I, int Ipad)
{
    const size_t n = (size_t)B * T * I;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % I);
        const size_t r = i / I;
        const int t = (int)(r % T), b = (int)(r / T);
        dx[i] = dxp[((size_t)t * B + b) * Ipad + c];
    }
}
// packed gradients -> PyTorch parameter layouts.  dwih [2*1056][Ipad], dwhh [2*1056][352], dbi / dbh [2*1056]
__global__ void gru_unpack_grads_kernel(const float* __restrict__ dwih, const float* __restrict__ dwhh, const float* __restrict__ dbi,
                                        const float* __restrict__ dbh, float* gw_ih0, float* gw_ih1, float* gw_hh0, float* gw_hh1,
                                        float* gb_ih0, float* gb_ih1, float* gb_hh0, float* gb_hh1, int l, int I, int Iin, int H)
{
    const int Ipad = gru_in_pad(l, I);
    const int orow = blockIdx.x;                       // 0 .. 2*3H-1: (dir, gate*H + unit)
    const int dir = orow / (3 * H), gj = orow % (3 * H), q = gj / H, j = gj % H;
    const int prow = dir * GRU_G + q * GRU_HP + j;
    float* wi = dir ? gw_ih1 : gw_ih0;
    float* wh = dir ? gw_hh1 : gw_hh0;
    for (int c = threadIdx.x; c < Iin; c += blockDim.x) wi[(size_t)gj * Iin + c] = dwih[(size_t)prow * Ipad + gru_in_col(l, H, c)];
    for (int c = threadIdx.x; c < H; c += blockDim.x) wh[(size_t)gj * H + c] = dwhh[(size_t)prow * GRU_HP + c];
    if (threadIdx.x == 0) {
        (dir ? gb_ih1 : gb_ih0)[gj] = dbi[prow];
        (dir ? gb_hh1 : gb_hh0)[gj] = dbh[prow];
    }
}

// ---------------------------------------------------------------- workspace layout
struct GruLayout {
    int Ipad[HOPK_GRU_MAX_LAYERS];
    size_t xb, gi, y[HOPK_GRU_MAX_LAYERS], r[HOPK_GRU_MAX_LAYERS], z[HOPK_GRU_MAX_LAYERS], n[HOPK_GRU_MAX_LAYERS], hn[HOPK_GRU_MAX_LAYERS];
    size_t wih[HOPK_GRU_MAX_LAYERS], whh[HOPK_GRU_MAX_LAYERS], bias[HOPK_GRU_MAX_LAYERS], bhn[HOPK_GRU_MAX_LAYERS], total;
    // backward scratch
    size_t s_whhT, s_dya, s_dyb, s_dgi[2], s_dgh[2], s_dwih, s_dwhh, s_dbi, s_dbh, s_total;   // dGi / dGh double-buffered by layer parity
};
static size_t gbump(size_t& cur, size_t bytes)
{
    size_t at = cur;
    cur += (bytes + 1023) & ~size_t(1023);
    return at;
}
static GruLayout gru_layout(const HopkGruShape* s)
{
    GruLayout g;
    memset(&g, 0, sizeof(g));
    const size_t TB = (size_t)s->T * s->B;
    size_t cur = 0;
    int ipmax = 2 * GRU_HP;                          // the dY buffers also hold the packed output gradient (704 wide)
    for (int l = 0; l < s->L; ++l) { g.Ipad[l] = gru_in_pad(l, s->I); if (g.Ipad[l] > ipmax) ipmax = g.Ipad[l]; }
    g.xb = gbump(cur, TB * g.Ipad[0] * 2);
    g.gi = gbump(cur, TB * 2 * GRU_G * 4);
    for (int l = 0; l < s->L; ++l) {
        g.y[l] = gbump(cur, TB * 2 * GRU_HP * 2);
        if (s->save) {
            g.r[l] = gbump(cur, TB * 2 * GRU_HP * 4); g.z[l] = gbump(cur, TB * 2 * GRU_HP * 4);
            g.n[l] = gbump(cur, TB * 2 * GRU_HP * 4); g.hn[l] = gbump(cur, TB * 2 * GRU_HP * 4);
        }
        g.wih[l] = gbump(cur, (size_t)2 * GRU_G * g.Ipad[l] * 2);
        g.whh[l] = gbump(cur, (size_t)2 * GRU_CL * GRU_FT_IMG);
        g.bias[l] = gbump(cur, (size_t)2 * GRU_G * 4);
        g.bhn[l] = gbump(cur, (size_t)2 * GRU_HP * 4);
    }
    g.total = cur;
    cur = 0;
    g.s_whhT = gbump(cur, (size_t)2 * GRU_CL * (GRU_BT_IMG + GRU_BS_BYTES));
    g.s_dya = gbump(cur, TB * ipmax * 4);
    g.s_dyb = gbump(cur, TB * ipmax * 4);
    for (int k = 0; k < 2; ++k) {
        g.s_dgi[k] = gbump(cur, TB * 2 * GRU_G * 2);
        g.s_dgh[k] = gbump(cur, TB * 2 * GRU_G * 2);
    }
    g.s_dwih = gbump(cur, (size_t)2 * GRU_G * ipmax * 4);
    g.s_dwhh = gbump(cur, (size_t)2 * GRU_G * GRU_HP * 4);
    g.s_dbi = gbump(cur, (size_t)2 * GRU_G * 4);
    g.s_dbh = gbump(cur, (size_t)2 * GRU_G * 4);
    g.s_total = cur;
    return g;
}

static int gru_check(const HopkGruShape* s)
{
    HOPK_REQUIRE(s->B >= 1 && s->T >= 1 && s->I >= 1, "gru sizes");
    HOPK_REQUIRE(s->H >= 1 && s->H <= GRU_HP, "gru: hidden size must be <= 352");
    HOPK_REQUIRE(s->L >= 1 && s->L <= HOPK_GRU_MAX_LAYERS, "gru: layer count");
    return 0;
}

static int grid_1d(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : b); }

}  // namespace hopk

using namespace hopk;

extern "C" size_t hopk_gru_workspace_bytes(const HopkGruShape* s) { return gru_layout(s).total; }
extern "C" size_t hopk_gru_scratch_bytes(const HopkGruShape* s) { return gru_layout(s).s_total; }

extern "C" int hopk_gru_forward(const HopkGruShape* s, const HopkGruParams* p, const float* x, float* out, void* ws_, void* stream)
{
    if (int rc = gru_check(s)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GruLayout g = gru_layout(s);
    char* ws = (char*)ws_;
    const int B = s->B, T = s->T, H = s->H, L = s->L;
    const int TB = T * B;
    HOPK_CUDA(configure_smem_once((const void*)gru_fwd_kernel, gru_fwd_smem()));
    gru_pack_x_kernel<<<grid_1d((size_t)TB * g.Ipad[0]), 256, 0, st>>>(x, (__nv_bfloat16*)(ws + g.xb), B, T, s->I, g.Ipad[0]);
    HOPK_LAUNCH_CHECK("gru_pack_x");
    const int slices = cdiv(B, GRU_NB);
    for (int l = 0; l < L; ++l) {
        const int Iin = l == 0 ? s->I : 2 * H;
        __nv_bfloat16* wih = (__nv_bfloat16*)(ws + g.wih[l]);
        float* bias = (float*)(ws + g.bias[l]);
        float* bhn = (float*)(ws + g.bhn[l]);
        gru_pack_wih_kernel<<<2 * GRU_G, 128, 0, st>>>(p->w_ih[l][0], p->w_ih[l][1], p->b_ih[l][0], p->b_ih[l][1], p->b_hh[l][0],
                                                        p->b_hh[l][1], wih, bias, bhn, l, s->I, Iin, H);
        HOPK_LAUNCH_CHECK("gru_pack_wih");
        gru_pack_whh_kernel<<<dim3(GRU_CL, 2, 24), 256, 0, st>>>(p->w_hh[l][0], p->w_hh[l][1], (uint8_t*)(ws + g.whh[l]), nullptr, H);
        HOPK_LAUNCH_CHECK("gru_pack_whh");
        const __nv_bfloat16* X = l == 0 ? (const __nv_bfloat16*)(ws + g.xb) : (const __nv_bfloat16*)(ws + g.y[l - 1]);
        // Gi = X . W_ih^T + bias   (both directions: N = 2112)
        if (int rc = gemm_bf16_launch(X, wih, ws + g.gi, bias, nullptr, TB, 2 * GRU_G, g.Ipad[l], g.Ipad[l], g.Ipad[l], 2 * GRU_G, 0, 0,
                                      0, 0, 0, 0.f, 1, st)) return rc;
        GruFwdArgs a;
        a.Gi = (const float*)(ws + g.gi); a.whh = (const uint8_t*)(ws + g.whh[l]); a.bhn = bhn;
        a.Y = (__nv_bfloat16*)(ws + g.y[l]);
        a.R = s->save ? (float*)(ws + g.r[l]) : nullptr; a.Z = s->save ? (float*)(ws + g.z[l]) : nullptr;
        a.N = s->save ? (float*)(ws + g.n[l]) : nullptr; a.HN = s->save ? (float*)(ws + g.hn[l]) : nullptr;
        a.out = l == L - 1 ? out : nullptr;
        a.B = B; a.T = T; a.H = H;
        gru_fwd_kernel<<<2 * slices * GRU_CL, GRU_THREADS, gru_fwd_smem(), st>>>(a);
        HOPK_LAUNCH_CHECK("gru_fwd");
    }
    return 0;
}

// Side stream of the backward pass: the weight-gradient work of layer l (dW_ih, dW_hh, bias column sums, unpacking) is not
// on the dependent chain recurrence(l) -> dX -> recurrence(l-1); it runs beside the next layer's recurrence kernel, whose
// 128 latency-bound CTAs leave 20 SMs idle.  One stream + events per device, created on first use (same conventions as the
// Graph-WaveNet side streams: non-blocking, calls on a device come from one host thread at a time).
struct GruSide {
    cudaStream_t s;
    cudaEvent_t ev_rec[HOPK_GRU_MAX_LAYERS], ev_done[HOPK_GRU_MAX_LAYERS];
};
static GruSide* gru_side()
{
    constexpr int MAXDEV = 64;
    static GruSide table[MAXDEV];
    static int states[MAXDEV] = {0};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    GruSide& sd = table[dev];
    int& state = states[dev];
    if (state == 0) {
        state = 1;
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) state = -1;
        for (int l = 0; l < HOPK_GRU_MAX_LAYERS; ++l) {
            if (cudaEventCreateWithFlags(&sd.ev_rec[l], cudaEventDisableTiming) != cudaSuccess) state = -1;
            if (cudaEventCreateWithFlags(&sd.ev_done[l], cudaEventDisableTiming) != cudaSuccess) state = -1;
        }
    }
    return state == 1 ? &sd : nullptr;
}

extern "C" int hopk_gru_backward(const HopkGruShape* s, const HopkGruParams* p, const float* dout, void* ws_, void* scratch_,
                                 const HopkGruGrads* gr, float* dx, void* stream)
{
    if (int rc = gru_check(s)) return rc;
    HOPK_REQUIRE(s->save, "gru backward needs a forward run with save = 1");
    cudaStream_t st = (cudaStream_t)stream;
    GruLayout g = gru_layout(s);
    char* ws = (char*)ws_;
    char* sc = (char*)scratch_;
    const int B = s->B, T = s->T, H = s->H, L = s->L;
    const int TB = T * B;
    HOPK_CUDA(configure_smem_once((const void*)gru_bwd_kernel, gru_bwd_smem()));
    float* dy_cur = (float*)(sc + g.s_dya);
    float* dy_next = (float*)(sc + g.s_dyb);
    gru_pack_dy_kernel<<<grid_1d((size_t)TB * 2 * GRU_HP), 256, 0, st>>>(dout, dy_cur, B, T, H);
    HOPK_LAUNCH_CHECK("gru_pack_dy");
    const int slices = cdiv(B, GRU_NB);
    float* dwih = (float*)(sc + g.s_dwih);
    float* dwhh = (float*)(sc + g.s_dwhh);
    float* dbi = (float*)(sc + g.s_dbi);
    float* dbh = (float*)(sc + g.s_dbh);
    GruSide* side = gru_side();
    HOPK_REQUIRE(side != nullptr, "gru backward: side stream");
    cudaStream_t ss = side->s;
    struct JoinOnError {                                 // an early error return still joins the side stream (graph capture)
        GruSide* sd; cudaStream_t st; bool armed;
        ~JoinOnError()
        {
            if (!armed) return;
            if (cudaEventRecord(sd->ev_done[0], sd->s) == cudaSuccess) cudaStreamWaitEvent(st, sd->ev_done[0], 0);
            cudaGetLastError();
        }
    } join_guard{side, st, true};
    for (int l = L - 1; l >= 0; --l) {
        const int Iin = l == 0 ? s->I : 2 * H;
        const int Ipad = g.Ipad[l];
        __nv_bfloat16* dgi = (__nv_bfloat16*)(sc + g.s_dgi[l & 1]);
        __nv_bfloat16* dgh = (__nv_bfloat16*)(sc + g.s_dgh[l & 1]);
        gru_pack_whh_kernel<<<dim3(GRU_CL, 2, 24), 256, 0, st>>>(p->w_hh[l][0], p->w_hh[l][1], (uint8_t*)(ws + g.whh[l]), (uint8_t*)(sc + g.s_whhT), H);
        HOPK_LAUNCH_CHECK("gru_pack_whhT");
        if (l + 2 < L) HOPK_CUDA(cudaStreamWaitEvent(st, side->ev_done[l + 2], 0));    // the side stream has finished with this dGi / dGh pair
        GruBwdArgs a;
        a.dY = dy_cur; a.whhT = (const uint8_t*)(sc + g.s_whhT); a.Y = (const __nv_bfloat16*)(ws + g.y[l]);
        a.R = (const float*)(ws + g.r[l]); a.Z = (const float*)(ws + g.z[l]); a.N = (const float*)(ws + g.n[l]); a.HN = (const float*)(ws + g.hn[l]);
        a.dGi = dgi; a.dGh = dgh; a.B = B; a.T = T;
        gru_bwd_kernel<<<2 * slices * GRU_CL, GRU_THREADS, gru_bwd_smem(), st>>>(a);
        HOPK_LAUNCH_CHECK("gru_bwd");
        HOPK_CUDA(cudaEventRecord(side->ev_rec[l], st));
        HOPK_CUDA(cudaStreamWaitEvent(ss, side->ev_rec[l], 0));
        const __nv_bfloat16* X = l == 0 ? (const __nv_bfloat16*)(ws + g.xb) : (const __nv_bfloat16*)(ws + g.y[l - 1]);
        // ---- side stream: parameter gradients of this layer
        // dW_ih[2112][Ipad] = dGi^T . X   (contraction over the T*B rows: both operands MN-major)
        if (int rc = gemm_bf16_launch(dgi, X, dwih, nullptr, nullptr, 2 * GRU_G, Ipad, TB, 2 * GRU_G, Ipad, Ipad, 1, 1, 0, 0, 0, 0.f, 2, ss))
            return rc;
        // dW_hh[dir][1056][352] = dGh_dir^T . h_{t-1}: the stored outputs shifted by one time step (B rows), zero outside
        for (int dir = 0; dir < 2; ++dir) {
            if (int rc = gemm_bf16_launch(dgh + dir * GRU_G, (const __nv_bfloat16*)(ws + g.y[l]) + dir * GRU_HP, dwhh + (size_t)dir * GRU_G * GRU_HP,
                                          nullptr, nullptr, GRU_G, GRU_HP, TB, 2 * GRU_G, 2 * GRU_HP, GRU_HP, 1, 1, 0, 0, 0, 0.f, 8, ss,
                                          nullptr, 0, dir ? B : -B))
                return rc;
        }
        if (int rc = hopk_colsum(dgi, dbi, TB, 2 * GRU_G, 2 * GRU_G, 1, ss)) return rc;
        if (int rc = hopk_colsum(dgh, dbh, TB, 2 * GRU_G, 2 * GRU_G, 1, ss)) return rc;
        gru_unpack_grads_kernel<<<2 * 3 * H, 128, 0, ss>>>(dwih, dwhh, dbi, dbh, gr->w_ih[l][0], gr->w_ih[l][1], gr->w_hh[l][0],
                                                            gr->w_hh[l][1], gr->b_ih[l][0], gr->b_ih[l][1], gr->b_hh[l][0], gr->b_hh[l][1],
                                                            l, s->I, Iin, H);
        HOPK_LAUNCH_CHECK("gru_unpack_grads");
        HOPK_CUDA(cudaEventRecord(side->ev_done[l], ss));
        // ---- main stream: the input gradient the next layer's recurrence waits for
        // dX[T*B][Ipad] = dGi . W_ih   (W_ih packed is [2112][Ipad] = [contraction][out]: B MN-major)
        if (l > 0 || dx) {
            if (int rc = gemm_bf16_launch(dgi, ws + g.wih[l], dy_next, nullptr, nullptr, TB, Ipad, 2 * GRU_G, 2 * GRU_G, Ipad, Ipad, 0, 1,
                                          0, 0, 0, 0.f, 1, st)) return rc;
            if (l == 0) {
                gru_unpack_dx_kernel<<<grid_1d((size_t)TB * s->I), 256, 0, st>>>(dy_next, dx, B, T, s->I, Ipad);
                HOPK_LAUNCH_CHECK("gru_unpack_dx");
            }
            float* tmp = dy_cur; dy_cur = dy_next; dy_next = tmp;
        }
    }
    for (int l = 0; l < L && l < 2; ++l) HOPK_CUDA(cudaStreamWaitEvent(st, side->ev_done[l], 0));   // join (ev_done[l] orders all later layers)
    join_guard.armed = false;
    return 0;
}
