/*
 * gaot_b200.h -- C ABI of libgaot_b200.so: the B200 (sm_100a) hot path of GAOT-3D.
 *
 * The reference (Shizheng-Wen/GAOT-3D) is pure Python; it has no FFI layer.  Each
 * entry point below replaces one third-party kernel call site of the reference
 * (file:line relative to the reference root).  A maintainer binds them from
 * Python with ctypes; gaot_3d_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered,
 *     nothing synchronises unless a *_host out-parameter is non-NULL;
 *   - positions are float32 [n,3] row-major (D=2 inputs are zero-padded by the host);
 *   - edge lists are int64 (the reference's edge_index dtype), CSR side-band is int32;
 *   - return value: 0 = ok, otherwise a gaot_status code; gaot_last_error() gives text;
 *   - no call allocates device memory: temporaries live in caller-provided workspaces
 *     whose size comes from the matching *_workspace_bytes() function.
 */
#ifndef GAOT_B200_H
#define GAOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum gaot_status {
    GAOT_OK = 0,
    GAOT_ERR_INVALID = 1,      /* bad argument (shape / range / null) */
    GAOT_ERR_WORKSPACE = 2,    /* workspace too small */
    GAOT_ERR_CUDA = 3,         /* CUDA runtime error (see gaot_last_error) */
    GAOT_ERR_UNSUPPORTED = 4   /* configuration outside the kernel envelope */
};

const char* gaot_last_error(void);
int  gaot_abi_version(void);
/* number of kernel launches issued by this library since load / last reset (bench.py gpu_launches) */
int64_t gaot_launch_count(void);
void gaot_launch_count_reset(void);
/* launches of this library's kernels that a CUDA-graph replay issued on the caller's behalf (counted at capture time) */
void gaot_launch_count_add(int64_t n);
/* optional CUDA-event timing of the main kernels on their launching stream; summary lines are
 * "kernel_name calls total_ms total_algorithmic_work" (work = bytes for HBM-bound kernels, FLOPs
 * for tensor-bound ones).  enable(on) also clears the records. */
void gaot_profile_enable(int on);
int gaot_profile_summary(char* buf, size_t buf_bytes);

/* ------------------------------------------------------------------ graph build
 * Replaces torch_cluster.radius via torch_geometric.nn.radius
 *   (reference src/model/layers/magno.py:193-200, :253-260).
 * For every query y: all sources x with fp32 ((dx*dx+dy*dy)+dz*dz) < fl32(r*r),
 * at most `cap` (PyG default 32), the FIRST `cap` by ascending x index, emitted
 * grouped by ascending y, ascending x inside a group.
 * Two phases: count (-> rowptr[ny+1], E) then emit (-> E edges); the workspace
 * must be left untouched between the two calls.
 */
size_t gaot_radius_workspace_bytes(int64_t nx, int64_t ny);
int gaot_radius_count(const float* x, int64_t nx, const float* y, int64_t ny,
                      double r, int cap, void* ws, size_t ws_bytes,
                      int32_t* rowptr /* [ny+1] out */, int64_t* E_host /* may be NULL */,
                      void* stream);
int gaot_radius_emit(const float* x, int64_t nx, const float* y, int64_t ny,
                     double r, int cap, void* ws, size_t ws_bytes,
                     const int32_t* rowptr, int64_t* out_y /* [E] */, int64_t* out_x /* [E] */,
                     void* stream);

/* Replaces torch_cluster.knn via torch_geometric.nn.knn (magno.py:183-189, :242-248).
 * For every query y the k nearest sources x (same fp32 distance), ascending
 * distance, ties -> lower x index.  Writes exactly ny*min(k,nx) edges grouped by y.
 * k <= 128.
 */
size_t gaot_knn_workspace_bytes(int64_t nx, int64_t ny);
int gaot_knn(const float* x, int64_t nx, const float* y, int64_t ny, int k,
             void* ws, size_t ws_bytes, int64_t* out_y, int64_t* out_x, void* stream);

/* Replaces torch_geometric.utils.coalesce (magno.py:220, :293): lexicographic
 * (row0,row1) sort + duplicate removal of an int64 edge list.  Outputs are sized E_in
 * by the caller; the unique count is written to *E_out_dev (device) and, when
 * E_out_host != NULL, copied to the host (synchronises the stream).
 */
size_t gaot_coalesce_workspace_bytes(int64_t E_in);
int gaot_coalesce(const int64_t* row0, const int64_t* row1, int64_t E_in,
                  int64_t max_row0, int64_t max_row1,
                  void* ws, size_t ws_bytes, int64_t* out0, int64_t* out1,
                  int64_t* E_out_dev, int64_t* E_out_host, void* stream);

/* Replaces torch_geometric.utils.dropout_edge (magno.py:367): keep[e] = u(e) >= p with a
 * counter-based Philox stream (seed, offset); compacts both rows preserving order.
 */
size_t gaot_edge_mask_workspace_bytes(int64_t E_in);
int gaot_edge_mask(const int64_t* row0, const int64_t* row1, int64_t E_in, double p_drop,
                   uint64_t seed, uint64_t offset, void* ws, size_t ws_bytes,
                   int64_t* out0, int64_t* out1, int64_t* E_out_dev, int64_t* E_out_host,
                   void* stream);

/* Query-major CSR side-band of an arbitrary-order edge list (precomputed int32/int64
 * edges arrive in any order: magno.py:506-516, stat.py:191).  Stable: inside a query
 * the original edge order is kept, so reductions are deterministic.
 *   rowptr[nq+1], csr_src[E], csr_qry[E], perm[E] (original edge id of CSR slot)
 */
size_t gaot_csr_workspace_bytes(int64_t E, int64_t nq);
int gaot_csr_from_edges(const int64_t* src, const int64_t* qry, int64_t E, int64_t n_src, int64_t nq,
                        int flags /* bit0: edges already grouped by ascending qry; bit1: VALIDATE -- range-check
                                     every index (0 <= qry < nq, 0 <= src < n_src) and the bit0 claim in the same
                                     pass; one 4-byte read-back; GAOT_ERR_INVALID instead of an out-of-bounds
                                     access (the reference's torch indexing raises IndexError there) */,
                        void* ws, size_t ws_bytes, int32_t* rowptr, int32_t* csr_src,
                        int32_t* csr_qry, int32_t* perm, void* stream);

/* ------------------------------------------------------------------ GNO (IntegralTransform)
 * Replaces the gather -> cat -> LinearChannelMLP -> (* f_y) -> scatter(mean) chain of
 *   reference src/model/layers/integral_transform.py:114-171 (torch index + cuBLAS +
 *   torch_scatter) by one fused kernel.
 *   out[q] = (1/max(cnt_q,1)) * sum_{e in CSR row q} MLP(cat[y_pos[src_e], x_pos[q] (,f_y[src_e])]) (* f_y[src_e])
 * MLP: n_layers Linear layers, exact-erf GELU between them (mlp.py:327-335).
 * `params` is one flat float32 buffer: for each layer W[out,in] row-major then b[out].
 * transform: 0 = linear, 1 = nonlinear, 2 = nonlinear_kernelonly, 3 = kernel only (f_y == NULL).
 * reduce: 0 = mean, 1 = raw sum (sharded encoder: partial sums, counts come from rowptr).
 * precision: 0 = fp32 CUDA cores, 1 = bf16 operands / fp32 accumulate on tcgen05.
 */
typedef struct gaot_mlp_desc {
    int32_t n_layers;        /* number of Linear layers, 1..6 */
    int32_t dims[8];         /* dims[0] = input width, dims[l+1] = output width of layer l */
} gaot_mlp_desc;

size_t gaot_gno_workspace_bytes(int64_t E, int64_t nq, const gaot_mlp_desc* mlp);
int gaot_gno_forward(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq,
                     const float* f_y, int32_t c_f,
                     const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry, int64_t E,
                     const gaot_mlp_desc* mlp, const float* params,
                     int transform, int reduce, int precision,
                     void* ws, size_t ws_bytes, float* out /* [nq, dims[n_layers]] */, void* stream);
/* Backward of the above: d_params (same flat layout as params), d_f_y [n_src, c_f] (may be NULL). */
int gaot_gno_backward(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq,
                      const float* f_y, int32_t c_f,
                      const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry, int64_t E,
                      const gaot_mlp_desc* mlp, const float* params,
                      int transform, int reduce, int precision,
                      const float* d_out, void* ws, size_t ws_bytes,
                      float* d_params, float* d_f_y, void* stream);

/* Attentional integral transform (reference integral_transform.py:128-141,:161-165, use_attn): the same fused kernels with
 * a per-edge weight (CSR order) multiplied into each edge's value before a SUM reduction (reduce must be 1); the weights
 * are the segment-softmax of the cosine / projected-dot scores, computed by the host.  The backward optionally returns
 * d loss / d edge_w [E] (needed for the learnable 'dot_product' projections).  These run on the FP32 kernels. */
int gaot_gno_forward_weighted(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                              int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                              int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                              int precision, const float* edge_w, void* ws, size_t ws_bytes, float* out, void* stream);
int gaot_gno_backward_weighted(const float* y_pos, int64_t n_src, const float* x_pos, int64_t nq, const float* f_y,
                               int32_t c_f, const int32_t* rowptr, const int32_t* csr_src, const int32_t* csr_qry,
                               int64_t E, const gaot_mlp_desc* mlp, const float* params, int transform, int reduce,
                               int precision, const float* edge_w, const float* d_out, void* ws, size_t ws_bytes,
                               float* d_params, float* d_f_y, float* d_edge_w, void* stream);

/* ------------------------------------------------------------------ geometric embedding statistics
 * Replaces the 5 scatters + batched eigvalsh of reference src/model/layers/geoembed.py:99-175:
 * per query [N_i, mean|y-x|, var|y-x|, centroid - x (3), eig(cov + 1e-6 I) descending (3)],
 * zero rows for empty queries; `feat` is [nq, 9] BEFORE the global z-score (geoembed.py:177-180),
 * which gaot_geo_zscore applies in place (unbiased std, std < 1e-6 -> 1).
 */
int gaot_geo_stats(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq,
                   const int32_t* rowptr, const int32_t* csr_src, float* feat, void* stream);
/* The two halves of gaot_geo_stats for a query set whose sources are sharded over ranks (intra-sample sharding,
 * no reference counterpart; the statistics are those of geoembed.py:126-162): per-query moment SUMS
 * moments[nq,12] = {n, sum d, sum d^2, sum (y-x) (3), sum (y-x)(y-x)^T upper triangle (6)} centred on the
 * query, which add across shards (all-reduce), then moments -> the same [nq,9] features. */
int gaot_geo_moments(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq,
                     const int32_t* rowptr, const int32_t* csr_src, float* moments, void* stream);
int gaot_geo_from_moments(const float* moments, int64_t nq, float* feat, void* stream);
size_t gaot_geo_zscore_workspace_bytes(int64_t nq);
int gaot_geo_zscore(float* feat, int64_t nq, int32_t nfeat, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ PointNet-style geometric embedding (pooling part)
 * Replaces the gathers + per-edge `pointnet_mlp` + scatter(max|mean) of reference src/model/layers/geoembed.py:184-213
 * (method='pointnet'): pooled[q, 0..31] = pool_e relu(W2 relu(W1 (y[src(e)] - x[q]) + b1) + b2) over the CSR row of q,
 * 0 for an empty query.  params = [W1 (32 x 3) | b1 (32) | W2 (32 x 32) | b2 (32)] fp32; pooling 0 = max, 1 = mean.
 * `argmax` [nq, 32] int32 (max pooling: CSR position of the first maximal edge per channel, -1 if none) is written by the
 * forward and read by the backward.  Backward: parameter gradients only (coordinates carry none), same layout as params. */
size_t gaot_pointnet_workspace_bytes(void);
int gaot_pointnet_forward(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq, const int32_t* rowptr,
                          const int32_t* csr_src, const float* params, int pooling, float* pooled, int32_t* argmax, void* stream);
int gaot_pointnet_backward(const float* src_pos, int64_t n_src, const float* qry_pos, int64_t nq, const int32_t* rowptr,
                           const int32_t* csr_src, const float* params, int pooling, const float* d_pooled, const int32_t* argmax,
                           void* ws, size_t ws_bytes, float* d_params, void* stream);

/* ------------------------------------------------------------------ fused two-layer node MLP
 * y = W2 gelu(W1 x + b1) + b2 on every row of x [n, c_in]: the decoder's projection head
 *   (reference src/model/layers/magno.py:640-644 `self.projection`, applied at :796-797; LinearChannelMLP / ChannelMLP
 *   forward mlp.py:327-335, :283-305).  W1 [hidden, c_in], W2 [c_out, hidden] row-major fp32 (a kernel-size-1 Conv1d
 *   weight has the same layout).  f16/bf16 tensor-core operands, fp32 accumulation, tanh-form GELU (rtol 2e-2 tier).
 * Envelope: c_in = 32, hidden = 256, c_out <= 8 (gaot_node_mlp2_supported); outside it the entry points return
 * GAOT_ERR_UNSUPPORTED and the caller keeps its library GEMMs.
 * backward: d_params = [dW1 (hidden x c_in) | db1 (hidden) | dW2 (c_out x hidden) | db2 (c_out)]; d_x may be null. */
int gaot_node_mlp2_supported(int32_t c_in, int32_t hidden, int32_t c_out);
size_t gaot_node_mlp2_workspace_bytes(int32_t c_in, int32_t hidden, int32_t c_out);
int gaot_node_mlp2_forward(const float* x, int64_t n, int32_t c_in, int32_t hidden, int32_t c_out, const float* w1,
                           const float* b1, const float* w2, const float* b2, float* y, void* stream);
int gaot_node_mlp2_backward(const float* x, const float* d_y, int64_t n, int32_t c_in, int32_t hidden, int32_t c_out,
                            const float* w1, const float* b1, const float* w2, void* ws, size_t ws_bytes, float* d_x,
                            float* d_params, void* stream);

/* ------------------------------------------------------------------ one-layer node MLP with a tiny input width
 * y = W x + b on every row of x [n, k_in], k_in <= 16, c_out <= 64, fp32 FMA (both precision tiers): the encoder's lifting
 *   layer (reference src/model/layers/magno.py:421-424 `self.lifting`, applied at :540-545; LinearChannelMLP mlp.py:327-335 /
 *   ChannelMLP :283-305 with one layer).  W [c_out, k_in] row-major (a kernel-size-1 Conv1d weight has the same layout), b may
 *   be null.  Streaming kernels: 4 (k_in + c_out) bytes per row each way; the weight gradient is a fixed-order (deterministic)
 *   two-stage reduction.  d_x and d_b may be null. */
int gaot_node_linear_supported(int32_t k_in, int32_t c_out);
size_t gaot_node_linear_workspace_bytes(int32_t k_in, int32_t c_out);
int gaot_node_linear_forward(const float* x, int64_t n, int32_t k_in, int32_t c_out, const float* w, const float* b, float* y,
                             void* stream);
int gaot_node_linear_backward(const float* x, const float* d_y, int64_t n, int32_t k_in, int32_t c_out, const float* w,
                              void* ws, size_t ws_bytes, float* d_x, float* d_w, float* d_b, void* stream);

/* ------------------------------------------------------------------ all-to-all over peer memory (intra-sample sharding)
 * Data movement of one all-to-all among the `world` GPUs of a node: block j (block_bytes, a multiple of 16) of `send` is stored
 *   into slot `rank` of peer j's receive buffer, peer_recv[j] being THIS process's mapping of that buffer (symmetric memory /
 *   CUDA IPC); peer_recv is a HOST array of `world` device pointers.  One kernel, every destination in parallel.  The caller
 *   orders the stores before the consumers with a cross-rank barrier on the same stream (tblock.py).  New with the sharded
 *   path: the reference has no intra-sample parallelism (src/trainer/stat.py:431-436 is sample-level DDP). */
int gaot_a2a_put(const void* send, const void* const* peer_recv, int32_t rank, int32_t world, int64_t block_bytes, void* stream);
/* the same store kernel; bcast != 0: the ONE block of `send` goes to slot `rank` of every peer (all-gather) */
int gaot_p2p_put(const void* send, const void* const* peer_recv, int32_t rank, int32_t world, int64_t block_bytes, int32_t bcast,
                 void* stream);
/* out[i] = sum_{r = 0..world-1} peer_src[r][offset_floats + i], i < n_floats (fp32, fixed rank order -> deterministic): the reduce
 *   half of a reduce-scatter / two-shot all-reduce, pulled from the peers' symmetric buffers.  Offset and count in multiples of 4. */
int gaot_p2p_reduce(const void* const* peer_src, int32_t world, int64_t offset_floats, int64_t n_floats, float* out, void* stream);

/* ------------------------------------------------------------------ latent attention
 * Replaces rotary_emb + F.scaled_dot_product_attention of reference
 *   src/model/layers/attn.py:110-128.  q [B,S,H*d], k,v [B,S,Hkv*d] float32 (projection
 *   outputs, token-major), out [B,S,H*d] float32.  d = 32 or 64, non-causal, no mask,
 *   scale 1/sqrt(d), optional 1-D RoPE over the sequence index (theta 10000, interleaved pairs).
 *   BF16 operands / FP32 accumulation on tcgen05 with TMEM accumulators.
 *   lse [B,H,S] (log2 domain: max + log2(sum)) is saved for the backward.
 */
size_t gaot_attn_workspace_bytes(int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d);
int gaot_attn_forward(const float* q, const float* k, const float* v, int64_t B, int64_t S,
                      int32_t H, int32_t Hkv, int32_t d,
                      const float* rope_freqs /* [d/2] = RotaryEmbedding.freqs, NULL: no RoPE */,
                      float dropout_p /* attn.py:122-126; 0 = off */, uint64_t dropout_seed,
                      void* ws, size_t ws_bytes, float* out, float* lse, void* stream);
int gaot_attn_backward(const float* q, const float* k, const float* v, const float* out,
                       const float* d_out, const float* lse, int64_t B, int64_t S,
                       int32_t H, int32_t Hkv, int32_t d, const float* rope_freqs,
                       float dropout_p, uint64_t dropout_seed,
                       void* ws, size_t ws_bytes, float* dq, float* dk, float* dv, void* stream);

/* ------------------------------------------------------------------ token-level dense layers (nn.Linear)
 * Replaces the fp32 cuBLAS SGEMMs behind nn.Linear of the latent transformer: q/k/v/o_proj
 *   (reference src/model/layers/attn.py:104-106,:129), SwiGLU w1/w2/w3 (:163), skip_proj over
 *   cat[x, skip] (:223), patch_linear (src/model/gaot_3d.py:205).  BF16 operands / FP32
 *   accumulation on tcgen05 (TMEM accumulators); inputs keep their row-major layout, no transposed copies.
 *   dtype codes: 0 = float32, 1 = bfloat16.  w is the nn.Linear weight [N, K] (row-major, ld = K).
 *   forward        : y[M,N]  = [x | x2][M,K] w^T (+ bias[N]) (+ residual[M,N]);  x2 (optional) supplies the
 *                    contraction columns k >= k_split, so cat[x, skip] is never materialised (k_split % 64 == 0)
 *   backward_input : dx[M,K] = dy[M,N] w (+ residual) (+= when accumulate)
 *   backward_weight: dw[N,K] = dy^T x (+= when accumulate), split-K with a fixed-order reduction (deterministic)
 *   Leading dimensions in elements; rows must be 16-byte aligned.  GAOT_ERR_UNSUPPORTED otherwise.
 */
size_t gaot_linear_workspace_bytes(int64_t M, int64_t N, int64_t K);
int gaot_linear_forward(const void* x, int x_dtype, int64_t ldx, const void* x2, int64_t ldx2, int64_t k_split,
                        const void* w, int w_dtype, int64_t M, int64_t N, int64_t K,
                        const float* bias, const float* residual, int64_t ldr,
                        void* y, int y_dtype, int64_t ldy, void* ws, size_t ws_bytes, void* stream);
int gaot_linear_backward_input(const void* dy, int dy_dtype, int64_t lddy, const void* w, int w_dtype,
                               int64_t M, int64_t N, int64_t K, const float* residual, int64_t ldr,
                               void* dx, int dx_dtype, int64_t lddx, int accumulate,
                               void* ws, size_t ws_bytes, void* stream);
int gaot_linear_backward_weight(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                                int64_t M, int64_t N, int64_t K, float* dw, int64_t lddw, int accumulate,
                                void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ TransformerBlock row-wise kernels
 * The pieces of reference src/model/layers/attn.py TransformerBlock.forward (:205-230) between the GEMMs, with
 * BF16 activations handed from kernel to kernel in the layout the tcgen05 dense kernel reads:
 *   RMSNorm (:167-178): y = x * rsqrt(mean(x^2) + eps) * w, statistics in fp32; y written as bf16 and/or fp32,
 *     rstd [M] kept for the backward.  backward: dx (+ dres, the residual-branch gradient) and dw (two-stage,
 *     fixed-order reduction).  H multiple of 128, <= 1024.
 *   SwiGLU gate (:163) on the fused [w1 x | w3 x] projection GU [M, 2F] bf16: a = silu(g) * u, and its backward.
 *   colsum: bias gradient of skip_proj (:223).   cast_bf16: fp32 weights -> bf16 tensor-core operands.
 */
int gaot_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* up to 8 independent casts in one launch (all weights of one transformer block); host arrays of device pointers */
int gaot_cast_bf16_batch(const float* const* src, void* const* dst_bf16, const int64_t* n, int32_t count, void* stream);
int gaot_rmsnorm_forward(const float* x, const float* w, int64_t M, int32_t H, float eps,
                         void* y_bf16 /* may be NULL */, float* y_f32 /* may be NULL */, float* rstd /* [M] */, void* stream);
size_t gaot_rmsnorm_backward_workspace_bytes(int32_t H);
int gaot_rmsnorm_backward(const float* dy, const float* x, const float* rstd, const float* w,
                          const float* dres /* may be NULL */, int64_t M, int32_t H,
                          float* dx, float* dw /* [H] */, void* ws, size_t ws_bytes, void* stream);
size_t gaot_colsum_workspace_bytes(int64_t M, int64_t N);
int gaot_colsum(const float* x, int64_t M, int64_t N, float* out /* [N] */, void* ws, size_t ws_bytes, void* stream);
int gaot_swiglu_forward(const void* gu_bf16, int64_t M, int32_t F, void* a_bf16, void* stream);
int gaot_swiglu_backward(const void* da_bf16, const void* gu_bf16, int64_t M, int32_t F, void* dgu_bf16, void* stream);

/* latent attention, fused-block flavour (same kernels as gaot_attn_forward/backward): the input is the bf16
 * output qkv [B*S, ld] of ONE [Wq;Wk;Wv] projection (columns H*d | Hkv*d | Hkv*d); `packed`
 * (gaot_attn_packed_bytes) receives the per-head bf16 operands with RoPE applied and is handed unchanged to the
 * backward (no re-packing); out_bf16 / d_out are bf16 token-major [B*S, H*d], out_f32 is the same output unrounded
 * (the backward forms D = rowsum(dO * O) from it; a rounded O costs q/k-gradient accuracy); d_qkv is the bf16 gradient of the
 * projection output (GQA group summed, RoPE undone), laid out like qkv. */
size_t gaot_attn_packed_bytes(int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d);
int gaot_attn_fused_forward(const void* qkv_bf16, int64_t ld, int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d,
                            const float* rope_freqs, float dropout_p, uint64_t dropout_seed,
                            void* packed, void* out_bf16, float* out_f32, float* lse, void* stream);
size_t gaot_attn_fused_backward_workspace_bytes(int64_t B, int64_t S, int32_t H, int32_t d);
int gaot_attn_fused_backward(const void* packed, const float* out_f32, const void* d_out_bf16, const float* lse,
                             int64_t B, int64_t S, int32_t H, int32_t Hkv, int32_t d, const float* rope_freqs,
                             float dropout_p, uint64_t dropout_seed, void* ws, size_t ws_bytes,
                             void* d_qkv_bf16, int64_t ld, void* stream);

/* ------------------------------------------------------------------ host-buffer convenience (e2e arm)
 * Same graph build with HOST inputs/outputs: copies positions H2D, runs the kernels,
 * copies the edge list D2H.  Returns E through *E_host; out rows sized by the caller
 * (ny*cap for radius, ny*k for knn).
 */
int gaot_radius_host(const float* x_host, int64_t nx, const float* y_host, int64_t ny, double r, int cap,
                     int64_t* out_y_host, int64_t* out_x_host, int64_t* E_host);
int gaot_knn_host(const float* x_host, int64_t nx, const float* y_host, int64_t ny, int k,
                  int64_t* out_y_host, int64_t* out_x_host, int64_t* E_host);

#ifdef __cplusplus
}
#endif
#endif /* GAOT_B200_H */
