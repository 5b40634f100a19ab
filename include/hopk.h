/* hopk.h -- C ABI of libhopk.so: the sm_100a kernels behind HOP's training hot path.
 *
 * The reference (Chenghyyy/HOP-...) has no native code and no plugin/FFI registry: its
 * "operator interface" for this path is the Python class surface
 *     model/gwnet.py:8-46    nconv / linear / gcn
 *     model/gwnet.py:49-249  gwnet.__init__ / gwnet.forward
 *     model/HOP.py:255-299   ReprogrammingLayer.__init__ / forward / reprogramming
 * and autograd's backward of each.  Every entry point below replaces the ATen/cuDNN/cuBLAS call
 * sequence behind one of those methods (file:line cited per function).  The host-side mirror in
 * hop_b200/{gwnet,HOP}.py binds these symbols with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers + sizes only; no torch types; all pointers are DEVICE pointers unless noted.
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the host, never allocates: outputs and workspaces are caller-owned.
 *   - return 0 on success; non-zero = error, message via hopk_last_error() (thread-local).
 *   - dtype: 0 = fp32 storage + fp32 FFMA math (1e-5 parity mode),
 *            1 = bf16 storage / tensor-core math with fp32 accumulation (2e-2 mode).
 *   - "rows layout": an NCHW tensor (B, C, V, T) of the reference stored as rows
 *     r = (b*T + t)*V + v with the C channels contiguous (i.e. the contiguous (B, T, V, C) array).
 */
#ifndef HOPK_H
#define HOPK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOPK_MAX_LAYERS 16

const char* hopk_last_error(void);
int hopk_version(void);
long long hopk_launch_count(void);
/* hopk_launch_count: kernels this library has launched in this process (host-side counter) */
#ifdef HOPK_DEBUG
int hopk_debug_set(void* mapped_host_ints);   /* debug builds only (nvcc -DHOPK_DEBUG): progress markers of the tensor-core kernels */
#endif

/* ------------------------------------------------------------------ Graph-WaveNet block */
typedef struct HopkGwnetShape {
    int B, V, T;                 /* batch, nodes (bones), input time steps (before left-padding) */
    int in_dim, out_dim;         /* 173 / 173 in HOP (model/HOP.py:141-143) */
    int C;                       /* residual_channels == dilation_channels (64) */
    int S, E;                    /* skip_channels (256), end_channels (512) */
    int L;                       /* blocks*layers (8) */
    int dil[HOPK_MAX_LAYERS];    /* dilation of each gated conv (1,2,1,2,...), gwnet.py:98-123 */
    int rank;                    /* adaptive-adjacency embedding width (10), gwnet.py:82-83 */
    int training;                /* 1: BatchNorm batch statistics + running-stat update */
    int dtype;                   /* see above */
    float bn_momentum, bn_eps;   /* of the module's BatchNorm2d layers (nn.BatchNorm2d defaults 0.1 / 1e-5, gwnet.py:120) */
} HopkGwnetShape;

/* Parameter / buffer pointers, named after the reference's state_dict keys (gwnet.py:50-139).
 * All fp32 in their native PyTorch layouts: conv weights (out, in, 1, k) contiguous. */
typedef struct HopkGwnetParams {
    const float *nodevec1, *nodevec2;                    /* (V, rank), (rank, V) */
    const float *start_w, *start_b;                      /* (C, in_dim, 1, 1), (C) */
    const float *filter_w[HOPK_MAX_LAYERS], *filter_b[HOPK_MAX_LAYERS];   /* (C, C, 1, 2) */
    const float *gate_w[HOPK_MAX_LAYERS], *gate_b[HOPK_MAX_LAYERS];       /* (C, C, 1, 2) */
    const float *skip_w[HOPK_MAX_LAYERS], *skip_b[HOPK_MAX_LAYERS];       /* (S, C, 1, 1) */
    const float *mlp_w[HOPK_MAX_LAYERS], *mlp_b[HOPK_MAX_LAYERS];         /* gconv.i.mlp.mlp (C, 3C, 1, 1) */
    const float *bn_w[HOPK_MAX_LAYERS], *bn_b[HOPK_MAX_LAYERS];           /* (C) */
    float *bn_mean[HOPK_MAX_LAYERS], *bn_var[HOPK_MAX_LAYERS];            /* running stats, updated in place */
    int64_t *bn_nbt[HOPK_MAX_LAYERS];                                     /* num_batches_tracked */
    const float *end1_w, *end1_b;                        /* (E, S, 1, 1) */
    const float *end2_w, *end2_b;                        /* (out_dim, E, 1, 1) */
} HopkGwnetParams;

/* Gradient destinations (fp32, same shapes as the parameters, OVERWRITTEN not accumulated).
 * Entries may be NULL for tensors the reference never reaches (last layer's mlp / bn affine). */
typedef struct HopkGwnetGrads {
    float *nodevec1, *nodevec2;
    float *start_w, *start_b;
    float *filter_w[HOPK_MAX_LAYERS], *filter_b[HOPK_MAX_LAYERS];
    float *gate_w[HOPK_MAX_LAYERS], *gate_b[HOPK_MAX_LAYERS];
    float *skip_w[HOPK_MAX_LAYERS], *skip_b[HOPK_MAX_LAYERS];
    float *mlp_w[HOPK_MAX_LAYERS], *mlp_b[HOPK_MAX_LAYERS];
    float *bn_w[HOPK_MAX_LAYERS], *bn_b[HOPK_MAX_LAYERS];
    float *end1_w, *end1_b, *end2_w, *end2_b;
    /* optional: when every buffer above is a view into one allocation, its base and size; it is then cleared with a
     * single memset (the weight-gradient kernels accumulate split-K partials).  NULL: each buffer is cleared separately. */
    void* flat; size_t flat_bytes;
} HopkGwnetGrads;

/* bytes of the forward workspace (holds everything backward re-reads) and of backward's scratch */
size_t hopk_gwnet_workspace_bytes(const HopkGwnetShape* s);
size_t hopk_gwnet_scratch_bytes(const HopkGwnetShape* s);
int hopk_gwnet_out_steps(const HopkGwnetShape* s);       /* max(T, receptive_field) - rf + 1 */

/* Where a saved activation lives inside the forward workspace (for tools and tests that inspect what backward will
 * re-read).  name: "x0" (start conv output), "u" (pre-BatchNorm output of layer `layer`), "tf" / "sg" (tanh f, sigmoid g
 * of layer `layer`), "ycat" (last-T_out slices of every layer's gated output, 8C wide), "r0" (relu(skip)), "r1"
 * (relu(end_conv_1)), "mr" (BatchNorm mean | rstd of layer `layer`), "A" (adaptive adjacency).  All rows layout;
 * elem_bytes is 4 (fp32) or 2 (bf16).  Returns non-zero for an unknown name. */
int hopk_gwnet_ws_field(const HopkGwnetShape* s, const char* name, int layer, size_t* offset, size_t* bytes, int* elem_bytes);

/* gwnet.forward (model/gwnet.py:143-249).
 *   x: input viewed as (B, in_dim, V, T) through element strides xs = {sB, sC, sV, sT}
 *      (so both an NCHW tensor and HOP.Model's permuted (B,T,V,C) buffer are read in place).
 *   out: (B, out_dim, V, T_out) contiguous NCHW, fp32.
 *   ws: hopk_gwnet_workspace_bytes() bytes; must be kept untouched until backward.           */
int hopk_gwnet_forward(const HopkGwnetShape* s, const HopkGwnetParams* p,
                       const float* x, const int64_t xs[4], float* out, void* ws, void* stream);

/* autograd backward of the call above.
 *   dout: (B, out_dim, V, T_out) contiguous.   dx: (B, T, V, in_dim) contiguous ("rows layout"),
 *   may be NULL when the input needs no gradient.                                             */
int hopk_gwnet_backward(const HopkGwnetShape* s, const HopkGwnetParams* p,
                        const float* x, const int64_t xs[4], const float* dout,
                        void* ws, void* scratch, const HopkGwnetGrads* g, float* dx, void* stream);

/* ------------------------------------------------------------------ Graph-WaveNet operators one at a time
 * The kernels hopk_gwnet_forward / backward chain together, exposed per operator for isolated tests and timing.  Rows layout
 * everywhere: x (B, Ti, V, C), outputs of a layer (B, To = Ti - d, V, C).  ss = BatchNorm scale | shift (2C floats) folded into
 * every read of x (pass ones | zeros for a plain input).  dtype as above.
 *
 * adaptive adjacency (gwnet.py:161-164): A5 = [A | A^2 | A^T | (A^2)^T | Z = E1 E2], 5 V*V floats, A = softmax(relu(Z), dim=1) */
int hopk_adp_softmax_fwd(const float* e1, const float* e2, int V, int R, float* A5, void* stream);
int hopk_adp_softmax_bwd(const float* e1, const float* e2, const float* A5, const float* dA, int V, int R, float* de1, float* de2, void* stream);
/* gated dilated conv (gwnet.py:186-200): tf = tanh(filter conv), sg = sigmoid(gate conv), y = tf * sg; weights (C, C, 1, 2).
 * bwd: dx = gradient w.r.t. the BN-folded input (nullable), dw / db overwritten; scratch = hopk_gated_tcn_bwd_scratch_bytes */
int hopk_gated_tcn_fwd(const float* x, const float* ss, const float* wf, const float* bf, const float* wg, const float* bg, int B, int V,
                       int Ti, int d, int C, int dtype, float* tf, float* sg, float* y, void* stream);
size_t hopk_gated_tcn_bwd_scratch_bytes(int B, int V, int Ti, int d, int C);
int hopk_gated_tcn_bwd(const float* x, const float* ss, const float* wf, const float* wg, const float* tf, const float* sg, const float* dy,
                       int B, int V, int Ti, int d, int C, int dtype, void* scratch, float* dx, float* dwf, float* dbf, float* dwg,
                       float* dbg, void* stream);
/* graph convolution + residual + BatchNorm statistics (gwnet.py:33-46, 233, 237): x1 = A^T y, x2 = (A^2)^T y (per node group),
 * u = Wm [y | x1 | x2] + bm + (ss-folded) xres[t + d]; stats = per-channel sum | sum of squares of u (2C doubles).
 * bwd (from du): dy, dWm (C, 3C), dbm, dA (V, V); the residual branch receives du itself.  scratch = hopk_gcn_scratch_bytes */
size_t hopk_gcn_scratch_bytes(int B, int V, int To, int C);
int hopk_gcn_diffuse_mlp_res_bnstat_fwd(const float* y, const float* A5, const float* xres, const float* ss, const float* wm,
                                        const float* bm, int B, int V, int Ti, int d, int C, int dtype, float* x1, float* x2, float* u,
                                        double* stats, void* stream);
int hopk_gcn_diffuse_mlp_res_bnstat_bwd(const float* du, const float* y, const float* x1, const float* x2, const float* A5,
                                        const float* wm, int B, int V, int To, int C, int dtype, void* scratch, float* dy, float* dwm,
                                        float* dbm, float* dA, void* stream);
/* BatchNorm2d finalize (gwnet.py:120, 237): statistics -> mean | rstd, folded scale | shift for the next layer's reads,
 * running-statistics update (training) */
int hopk_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float* rmean, float* rvar, int64_t* nbt,
                     float* mean_rstd, float* scale_shift, int C, int training, float momentum, float eps, void* stream);

/* nconv.forward (model/gwnet.py:12-14): out[n,c,w,l] = sum_v x[n,c,v,l] * A[v,w], contiguous NCHW.
 * hopk_nconv_bwd: dx (same layout) and dA (V x V, overwritten). */
int hopk_nconv_fwd(const float* x, const float* A, float* out, int N, int C, int V, int T, void* stream);
int hopk_nconv_bwd(const float* x, const float* A, const float* dout, float* dx, float* dA,
                   int N, int C, int V, int T, void* stream);

/* ------------------------------------------------------------------ dense layers
 * y[M,N] = act_in(x[M,K]) * w[N,K]^T + b ; flags: 1 = ReLU on the input (HOP.py:284-285 fuses the
 * activation before out_projection), 2 = ReLU on the output.  gwnet's `linear` (gwnet.py:16-22),
 * the reprogramming Q/K/V/O projections (HOP.py:262-265, 276-278, 285). */
int hopk_linear_fwd(const float* x, const float* w, const float* b, float* y,
                    int M, int N, int K, int flags, void* stream);
/* dx (nullable), dw, db (nullable) overwritten.  `y` is needed only with flag 2. */
int hopk_linear_bwd(const float* x, const float* w, const float* y, const float* dy,
                    float* dx, float* dw, float* db, int M, int N, int K, int flags, void* stream);
/* 1x1 Conv2d on NCHW (gwnet.py:16-22 `linear`, standalone use): x (B,K,V,T) -> y (B,N,V,T) */
int hopk_conv1x1_nchw_fwd(const float* x, const float* w, const float* b, float* y,
                          int B, int K, int N, int V, int T, void* stream);
int hopk_conv1x1_nchw_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                          int B, int K, int N, int V, int T, void* stream);

/* ------------------------------------------------------------------ dense bf16 GEMM (TMA + tcgen05)
 * C[M][N] (+)= sum_k A(m,k) B(n,k) (+ bias[n] + addend[m][n]) -> activation.   A, B: bf16 in global memory, 16-byte aligned,
 * leading dimensions (elements) multiples of 8.  Default operand layout is "K-major" (A is [M][K], B is [N][K] row-major:
 * x @ W^T with nn.Linear's weight, HOP.py:130-134,200,202,262-265); HOPK_GEMM_A_MN / _B_MN say the operand is stored with
 * the contraction index as the row ([K][M] / [K][N]), which covers dX = dY @ W and dW = dY^T @ X without transposes.
 * C: fp32 (default) or bf16 (HOPK_GEMM_OUT_BF16), leading dimension ldc.  addend: optional fp32 [M][ldc] added before the
 * activation.  mask: optional [M][ldc] (fp32, or bf16 with HOPK_GEMM_MASK_BF16); elements with mask <= 0 are multiplied by
 * `slope` (0: gradient of a fused ReLU; 0.2: of LeakyReLU(0.2)).
 * splits > 1: split-K with vector atomics into an fp32 C (cleared by the call unless HOPK_GEMM_ACCUMULATE). */
#define HOPK_GEMM_A_MN 1
#define HOPK_GEMM_B_MN 2
#define HOPK_GEMM_OUT_BF16 4
#define HOPK_GEMM_ACCUMULATE 8
#define HOPK_GEMM_RELU 16
#define HOPK_GEMM_LEAKY 32     /* LeakyReLU(slope) */
#define HOPK_GEMM_GELU 64      /* exact (erf) GELU */
#define HOPK_GEMM_BIAS_ROW 128 /* bias indexed by the output row m (length M) instead of the column n */
#define HOPK_GEMM_MASK_BF16 256 /* mask is stored as bf16 */
#define HOPK_GEMM_MASK_GELU 512 /* mask is a bf16 pre-activation x: the result is multiplied by gelu'(x) (exact erf form) */
int hopk_gemm_bf16(const void* A, const void* B, void* C, const float* bias, const float* addend, const void* mask,
                   int M, int N, int K, long lda, long ldb, long ldc, int flags, float slope, int splits, void* stream);
/* fp32 (rows x cols, ld lds) -> bf16 (rows x cols_out, ld ldd), zero padding for cols <= c < cols_out, optional ReLU */
int hopk_cast_bf16(const float* src, void* dst, long rows, int cols, long lds, int cols_out, long ldd, int relu, void* stream);
/* overlapping windows of B signals as a bf16 matrix: out[b*nwin + k][c] = x[b][k*hop + c], c < win (in_audio.unfold(1, 3400,
 * 2191) of HOP.py:210 plus the cast); ldx / ldo: leading dimensions of x and out in elements (ldo % 8 == 0) */
int hopk_unfold_bf16(const float* x, void* out, int B, int nwin, int win, int hop, long ldx, long ldo, void* stream);
/* out[c] = sum_r src[r][c] (bias gradients); src fp32 or bf16 */
int hopk_colsum(const void* src, float* out, long rows, int cols, long ld, int src_bf16, void* stream);

/* ------------------------------------------------------------------ beat features -> Graph-WaveNet rows (model/HOP.py:210-217)
 * feat: (B*16, F) beat-MLP output per audio window; seed: (B, 16, 3J) seed bones; rows: (B, 16, J, 3 + F) with
 * rows[b,t,j,:3] = seed[b,t,3j:3j+3] and rows[b,t,j,3:] = feat[b*16 + (t*J + j) % 16]  (the reference's repeat + view).
 * Backward: dfeat (bf16, leading dimension ldd) = the J-fold gather-sum of drows[..., 3:]; dbias (nullable) = its column sums. */
int hopk_beat_rows_fwd(const float* feat, const float* seed, float* rows, int B, int J, int F, void* stream);
int hopk_beat_rows_bwd(const float* drows, void* dfeat_bf16, float* dbias, int B, int J, int F, long ldd, void* stream);

/* The generator's loss terms of the training step, train_eval/train_llm.py:46-79, as one kernel forward and one backward:
 *   huber = mean smooth_l1(out / 0.1, tgt / 0.1) * 0.1;  pose[b] = sum smooth_l1(out / 0.05, rnd / 0.05) * 0.05;
 *   zl1[b] = mean |zc - zr|;  div_reg = mean_b clamp(-pose[b] / (zl1[b] + 1e-5), min = -1000);
 *   kld = -0.5 mean(1 + logvar - mu^2 - exp(logvar));  vals = {w_reg huber + w_div div_reg + w_kld kld, huber, div_reg, kld}.
 * out, tgt, rnd: (B, TP) fp32; zc, zr, mu, logvar: (B, Z); rnd / zc / zr NULL: no diversity term; mu / logvar NULL: no KLD.
 * per: (B, 4) fp32 scratch kept for backward; ticket: one zero-initialised uint32 (left zero).  Backward writes the
 * gradients of vals[0] with respect to out, mu, logvar, scaled by the device scalar *gl. */
int hopk_step_losses_fwd(const float* out, const float* tgt, const float* rnd, const float* zc, const float* zr, const float* mu,
                         const float* logvar, int B, int TP, int Z, float w_reg, float w_div, float w_kld, float* per, float* vals,
                         unsigned int* ticket, void* stream);
int hopk_step_losses_bwd(const float* out, const float* tgt, const float* rnd, const float* mu, const float* logvar, const float* per,
                         const float* gl, int B, int TP, int Z, float w_reg, float w_div, float w_kld, float* dout, float* dmu,
                         float* dlogvar, void* stream);

/* ------------------------------------------------------------------ frozen BERT encoder: the kernels between its GEMMs
 * (model/HOP.py:204 `self.llm_model(inputs_embeds=...)`; weights frozen, HOP.py:90-91: backward = dX only)
 * layer norm over rows of C = 128 k columns: v = x (+ add[row % period]); y = (v - mean) * rstd * gamma + beta; fp32 and / or
 * bf16 copies of y; stat[row] = {mean, rstd} for backward.  hopk_ln_bwd: dX of the same (fp32 and / or bf16). */
int hopk_ln_fwd(const float* x, const float* add, int period, const float* gamma, const float* beta, float eps, float* y32,
                void* y16, float* stat, int M, int C, void* stream);
int hopk_ln_bwd(const float* dy, const float* x, const float* add, int period, const float* gamma, const float* stat,
                float* dx32, void* dx16, int M, int C, void* stream);
/* exact GELU on n bf16 elements (n % 8 == 0) */
int hopk_gelu_bf16(const void* pre, void* out, long n, void* stream);
/* BertSelfAttention core without mask / dropout: qkv bf16 (B*S, 3*H*D) = [q heads | k heads | v heads] per row; ctx bf16
 * (B*S, H*D); P fp32 (B*H, S, S): optional copy of the softmax probabilities for tools (NULL: not stored).  Backward
 * recomputes the probabilities from qkv and writes dqkv (same layout as qkv) from dctx.  D = 64, S <= 64. */
int hopk_bert_attn_fwd(const void* qkv, void* ctx, float* P, int B, int S, int H, int D, void* stream);
int hopk_bert_attn_bwd(const void* qkv, const void* dctx, void* dqkv, int B, int S, int H, int D, void* stream);

/* ------------------------------------------------------------------ GRU decoder (model/HOP.py:166-167, 248)
 * Multi-layer bidirectional GRU, batch_first, zero initial state, PyTorch gate order (r, z, n); dtype-1 arithmetic (bf16
 * tensor-core operands, fp32 accumulation and fp32 recurrent state).  Hidden size <= 352 (HOP: 350).
 * Parameters are nn.GRU's own tensors: weight_ih_l{k}{_reverse} (3H, I_k), weight_hh (3H, H), bias_ih / bias_hh (3H);
 * index [layer][direction].  I_0 = I, I_k = 2H. */
#define HOPK_GRU_MAX_LAYERS 8
typedef struct HopkGruShape {
    int B, T, I, H, L;
    int save;                    /* 1: keep what backward needs in the workspace; 0: inference / no-grad pass */
} HopkGruShape;
typedef struct HopkGruParams {
    const float* w_ih[HOPK_GRU_MAX_LAYERS][2]; const float* w_hh[HOPK_GRU_MAX_LAYERS][2];
    const float* b_ih[HOPK_GRU_MAX_LAYERS][2]; const float* b_hh[HOPK_GRU_MAX_LAYERS][2];
} HopkGruParams;
typedef struct HopkGruGrads {    /* same shapes as the parameters, overwritten */
    float* w_ih[HOPK_GRU_MAX_LAYERS][2]; float* w_hh[HOPK_GRU_MAX_LAYERS][2];
    float* b_ih[HOPK_GRU_MAX_LAYERS][2]; float* b_hh[HOPK_GRU_MAX_LAYERS][2];
} HopkGruGrads;
size_t hopk_gru_workspace_bytes(const HopkGruShape* s);
size_t hopk_gru_scratch_bytes(const HopkGruShape* s);
/* x: (B, T, I) fp32; out: (B, T, 2H) fp32 (forward | reverse halves, like nn.GRU); ws: workspace, kept until backward */
int hopk_gru_forward(const HopkGruShape* s, const HopkGruParams* p, const float* x, float* out, void* ws, void* stream);
/* dout: (B, T, 2H); dx: (B, T, I) or NULL */
int hopk_gru_backward(const HopkGruShape* s, const HopkGruParams* p, const float* dout, void* ws, void* scratch,
                      const HopkGruGrads* g, float* dx, void* stream);

/* ------------------------------------------------------------------ reprogramming cross-attention
 * ReprogrammingLayer.reprogramming (model/HOP.py:289-299):
 *   O[b,l,h,:] = dropout(softmax_s(Q[b,l,h,:].K[s,h,:] / sqrt(E))) . V[s,h,:]
 * q,o: (B, L, H, E)   k,v: (S, H, E)   lse: (B, H, L) log-sum-exp saved for backward.
 * Dropout uses the counter-based mask documented in oracle/reprog_np.py (seed, flat index). */
int hopk_xattn_fwd(const float* q, const float* k, const float* v, float* o, float* lse,
                   int B, int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream);
/* dq (B,L,H,E), dk, dv (S,H,E): overwritten.  delta: caller-owned scratch of B*H*L floats. */
int hopk_xattn_bwd(const float* q, const float* k, const float* v, const float* o, const float* lse,
                   const float* dout, float* dq, float* dk, float* dv, float* delta,
                   int B, int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream);

/* Dropout epoch: a device-side value added to the `seed` argument of every attention call.  A training step captured in
 * a CUDA graph bakes `seed` into its launch arguments; capturing one call of this function (reset = 0) at the top of the
 * step makes every replay draw a fresh mask.  reset = 1 puts the epoch back to 0 (the state tests and oracles assume). */
int hopk_dropout_epoch_advance(int reset, void* stream);

/* dtype-1 variants of the two calls above: bf16 operands on tcgen05 (UMMA 128x128, TMEM accumulators), fp32
 * accumulation and fp32 I/O; head dim must be 128.  Same arguments and semantics. */
/* kv_pack: scratch of hopk_xattn_pack_bytes(S, H) bytes (K/V re-packed as bf16 UMMA slab images and streamed with
 * cp.async.bulk); NULL selects the variant that stages K/V from fp32 with its own threads. */
size_t hopk_xattn_pack_bytes(int S, int H);
int hopk_xattn_fwd_tc(const float* q, const float* k, const float* v, float* o, float* lse, void* kv_pack,
                      int B, int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream);
/* backward v3: kv_pack = the forward's packed K/V records (K, V unchanged), scratch = hopk_xattn_bwd_scratch_bytes(B, L, H)
 * bytes (Q / dO records + per-row statistics).  Either NULL selects the kernels that stage fp32 operands themselves. */
size_t hopk_xattn_bwd_scratch_bytes(int B, int L, int H);
int hopk_xattn_bwd_tc(const float* q, const float* k, const float* v, const float* o, const float* lse,
                      const float* dout, float* dq, float* dk, float* dv, float* delta, void* kv_pack, void* scratch,
                      int B, int L, int H, int E, int S, float p_drop, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HOPK_H */
