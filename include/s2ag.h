/* libs2ag_b200.so -- C ABI of the B200-native Speech2AffectiveGestures GAN-step hot path.
 *
 * The reference (UttaranB127/speech2affective_gestures) has no FFI of its own: its boundary for
 * this path is the Python module surface of net/multimodal_context_net_v2.py and
 * processor_v2.py.  Each entry point below replaces the torch.nn operator
 * call sites named in its comment (file:line are relative to the reference checkout); the
 * Python mirror in speech2affective_gestures_b200/net/ binds them through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (fp32 unless stated otherwise);
 *     workspaces are caller-provided, nothing is allocated, nothing synchronises the host;
 *   - activations are CHANNELS-LAST: a Conv1d tensor the reference holds as [N, C, L] is
 *     [N, L, C] here, a Conv2d tensor [N, C, T, V] is [N, T, V, C]; weights keep the
 *     reference (PyTorch state_dict) layout unless stated;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 ok, <0 error (s2ag_last_error() has the text).  Never throws.
 *   - "+=" in a comment: the kernel ACCUMULATES into that buffer (parameter gradients live in
 *     one flat buffer per network that is zeroed once per optimiser step).
 */
#ifndef S2AG_H_
#define S2AG_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int s2ag_version(void);
const char* s2ag_last_error(void);
/* 1 when the library was compiled from CUDA sources for sm_100a; the CPU logic-emulation build
 * used by tests/emu reports 0 and is never loaded by the product path. */
int s2ag_is_device_build(void);
/* number of kernels this library has launched (or recorded into a graph capture) so far */
unsigned long long s2ag_launch_count(void);
/* debugging aid: cudaStreamIsCapturing status of `stream` (0 none, 1 active, 2 invalidated, <0 error) */
int s2ag_stream_capture_status(void* stream);

/* Dense-contraction engine: 0 = auto (tcgen05 tensor-core kernel wherever a 128 x BN tile is worthwhile,
 * exact-fp32 SIMT kernel for tiny shapes), 1 = SIMT everywhere (A/B parity checks). */
int s2ag_set_engine(int engine);
/* Scratch for the tensor-core operand images of weight operands (gemm_umma_packed.cuh): a device buffer owned by the
 * caller, registered for ONE stream (kernels of that stream use it in stream order).  Without a registered buffer
 * that is large enough the contraction stages both operands on the fly.  (NULL, 0) unregisters. */
int s2ag_register_scratch(void* stream, void* buf, long bytes);
/* Tensor-core operand precision: 0 = "bf16x3" (fp32 operands split hi+lo on the fly, three MMAs, fp32
 * accumulate in TMEM: fp32-grade results, the default and the parity configuration), 1 = "bf16x1"
 * (single bf16 pass, BASELINE config 3). */
int s2ag_set_precision(int mode);
/* Bring-up / A-B switches of the tcgen05 kernels (0 = production behaviour).  Bits:
 *   2   clock64 / globaltimer timeline of the persistent GRU forward (s2ag_debug_read_timeline)
 *   4   stride-1 convolutions and their weight gradients through the implicit-GEMM engine (no shifted-window kernels)
 *   8   timeline of the persistent BPTT kernel
 *   16  no FAST (LDG.256, unguarded) operand-load variant of the contraction kernels
 *   32  one stderr line per dense contraction with its device time (eager mode only; synchronises)
 *   64  conv_wgrad_shift_kernel skips its red.global epilogue (timing experiment; results are wrong)
 *   128 conv_wgrad_shift_kernel wherever it is structurally applicable (default: only where it was measured faster)
 *   256 no packed-weight route: both operands of every contraction are converted on the fly
 *   1024 L2-exchange GRU forward (umma_gru.cu) also where the cluster kernel (umma_gru_cluster.cu) is the default
 *   2048 cluster GRU forward for every supported hidden size (default: H <= 80; see umma_gru_cluster.cu)
 *   4096 cluster GRU forward: wait for every h slice before the first tcgen05.mma of a step (measurement)
 *   8192 fused TCN block: CTA-pair variant (tcgen05 cta_group::2, M = 256; measured slower than one CTA per tile)
 *   16384 two-TMA contraction kernel (gemm_umma_tt.cuh: both operands pre-packed, 256-row tiles, three MMA issuers) for the
 *         big weight-side contractions; measured on par with the packed-B kernel, hence opt-in */
int s2ag_debug_flags(int flags);
/* bring-up aid: clock64 timeline (64 steps x 16 marks) of one CTA of the last persistent GRU forward launched with
 * s2ag_debug_flags bit 1 set; copies n values to HOST memory (synchronises). */
int s2ag_debug_read_timeline(long long* host, int n);
/* bring-up aid: number of co-resident clusters of the cluster GRU forward (backward = 0) / BPTT (1) kernel for hidden size H */
int s2ag_debug_gru_cluster_occupancy(int H, int backward);

#define S2AG_ACT_NONE 0
#define S2AG_ACT_RELU 1
#define S2AG_ACT_LEAKY 2

/* ---- nn.Linear (net/multimodal_context_net_v2.py:49,57,78,90,473-475,483-485,562-563) --------
 * y[M,N] = act(x[M,K] @ w[N,K]^T + bias[N]); row strides ldx / ldy let x and y be column
 * slices of wider buffers (the torch.cat at :526,:539 is never materialised). */
int s2ag_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* y, long ldy,
                    int M, int N, int K, int act, float slope, void* stream);
/* dx[M,K] (+)= dy[M,N] @ w[N,K] */
int s2ag_linear_bwd_data(const float* dy, long lddy, const float* w, float* dx, long lddx,
                         int M, int N, int K, int accumulate, void* stream);
/* dw[N,K] += dy^T x ; db[N] += colsum(dy) (db may be NULL) */
int s2ag_linear_bwd_weight(const float* dy, long lddy, const float* x, long ldx, float* dw, float* db,
                           int M, int N, int K, void* stream);
/* Linear over the ROW axis of each batch item (MFCCEncoder.linear1, :49,57: the reference applies
 * Linear(num_mfcc -> 32) to the last axis of [B, C, L]; channels-last holds that tensor as x[B, L, C]):
 *   y[b, c, n] = act( sum_l x[b, l, c] * w[n, l] + bias[n] ),   y row (b,c) at y + (b*C + c)*ldy */
int s2ag_linear_t_fwd(const float* x, const float* w, const float* bias, float* y, long ldy,
                      int B, int L, int C, int N, int act, float slope, void* stream);
/* dx[b, l, c] = sum_n dy[b, c, n] w[n, l] */
int s2ag_linear_t_bwd_data(const float* dy, long lddy, const float* w, float* dx,
                           int B, int L, int C, int N, void* stream);
/* dw[n, l] += sum_{b,c} dy[b, c, n] x[b, l, c] ; db[n] += sum_{b,c} dy[b, c, n] */
int s2ag_linear_t_bwd_weight(const float* dy, long lddy, const float* x, float* dw, float* db,
                             int B, int L, int C, int N, void* stream);
/* dpre = dy * act'(y) evaluated from the activation OUTPUT y; [M,N] with row strides */
int s2ag_act_bwd(const float* dy, long lddy, const float* y, long ldy, float* dpre, long ldd,
                 int M, int N, int act, float slope, void* stream);
/* y[M,H] = x[M, 0:H] + x[M, H:2H]  (sum of the bidirectional GRU halves, :542, :579) */
int s2ag_add_halves(const float* x, float* y, int M, int H, void* stream);
/* y = x * mask(seed)/(1-p) (nn.Dropout / nn.GRU inter-layer dropout); stateless counter RNG so
 * the backward call regenerates the mask from the same seed. In-place allowed. */
int s2ag_dropout(const float* x, float* y, long n, float p, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* seed_dev[0] += inc.  Every dropout-bearing call takes (seed, seed_dev): the effective seed is
 * seed + *seed_dev (seed_dev may be NULL), so a captured CUDA graph draws fresh masks on every
 * replay once the step nonce is advanced inside the graph. */
int s2ag_seed_advance(uint64_t* seed_dev, uint64_t inc, void* stream);

/* ---- nn.Conv1d / nn.Conv2d over channels-last activations -----------------------------------
 * (WavEncoder :17-27, MFCCEncoder :39-45, AffEncoder :146-150, STGraphConv tgcn.py:64-68,
 *  181-187,199-203, ConvDiscriminator pre_conv :397-404).
 * x[N,H,W,Cin] (pixel stride ldpix_x >= Cin), w[Cout,Cin,KH,KW] (reference layout),
 * y[N,Ho,Wo,Cout] (pixel stride ldpix_y).  Conv1d: W = KW = 1. */
int s2ag_conv_fwd(const float* x, long ldpix_x, int N, int H, int W, int Cin,
                  const float* w, const float* bias, float* y, long ldpix_y, int Cout,
                  int KH, int KW, int sh, int sw, int ph, int pw, int dh, int dw,
                  int act, float slope, void* stream);
/* dx (+)= conv_transpose(dy) ; stride must be 1 */
int s2ag_conv_bwd_data(const float* dy, long ldpix_dy, int N, int H, int W, int Cin,
                       const float* w, float* dx, long ldpix_dx, int Cout,
                       int KH, int KW, int ph, int pw, int dh, int dw, int accumulate, void* stream);
/* dw += , db += (db may be NULL) */
int s2ag_conv_bwd_weight(const float* dy, long ldpix_dy, const float* x, long ldpix_x,
                         int N, int H, int W, int Cin, float* dw, float* db, int Cout,
                         int KH, int KW, int sh, int sw, int ph, int pw, int dh, int dwd, void* stream);

/* ---- nn.BatchNorm1d/2d, train and eval, fused with the following activation -----------------
 * (:19-25,:40-46,:123,:131,:144,:149; tgcn.py:179,188,204).  x is [M,C] (M = all non-channel
 * positions), per-column statistics.  param_map (int32[C], may be NULL = identity) gives the
 * parameter index of each column and col_map (int32[C], may be NULL) the OUTPUT column, which
 * is how AffEncoder's view/permute regrouping (:155-171) is folded in.
 *   y[m, col_map[c]] = act( (x[m,c]-mean)/sqrt(var+eps)*gamma[pm[c]] + beta[pm[c]] + add[m, col_map[c]] )
 * training != 0: batch statistics (biased var), running stats updated with `momentum`
 * (unbiased var), save_mean/save_invstd[groups*C] written for the backward pass.
 * groups >= 1: the M rows are `groups` consecutive, equally sized, INDEPENDENT batches that the reference pushes
 * through the same module one call after the other (processor_v2.py:808-809: D(target), D(out)): statistics are per
 * group, the running statistics receive the groups' momentum updates in order.  ws: double[2*C*groups] scratch. */
int s2ag_bn_fwd(const float* x, long ldx, int M, int C, const float* gamma, const float* beta,
                const int32_t* param_map, float* running_mean, float* running_var,
                int training, float momentum, float eps,
                const float* add, long ldadd, float* y, long ldy, const int32_t* col_map,
                int act, float slope, float* save_mean, float* save_invstd, double* ws, int groups, void* stream);
/* A/B switch of the BatchNorm kernels: bit 0 = scalar kernels everywhere (default 0: float4 variants -- 4 channels per
 * thread, 16-byte loads -- whenever there is no column map, C % 4 == 0 and pitches / pointers are 16-byte aligned).
 * Returns the previous flags. */
int s2ag_debug_bn_flags(int flags);
/* dx = BN'(dy * act'(y)); dgamma += ; dbeta += ; dadd (may be NULL) = dy * act'(y).
 * ws: double[2*C*groups] scratch; groups as in s2ag_bn_fwd. */
int s2ag_bn_bwd(const float* dy, long lddy, const float* y, long ldy, const int32_t* col_map,
                const float* x, long ldx, int M, int C, const float* gamma, const int32_t* param_map,
                const float* save_mean, const float* save_invstd, int training, int act, float slope,
                float* dx, long lddx, float* dgamma, float* dbeta, float* dadd, long lddadd,
                double* ws, int groups, void* stream);

/* ---- causal-TCN residual block forward as ONE kernel (net/tcn.py:16-46; umma_tcn.cu): same semantics and dropout
 * streams as s2ag_weight_norm_fwd x2 + s2ag_tcn_block_fwd, but y1 stays on the SM (conv2 consumes it from shared
 * memory) and weight_norm is folded into the operand packing.  v1,g1,b1 / v2,g2,b2: weight_norm parameters ([C][C][2], [C], [C]);
 * w1/w2 [C][2][C] and n1/n2 [C] receive the effective fp32 weights and norms (inputs of s2ag_tcn_block_bwd /
 * s2ag_weight_norm_bwd); y1 / y2 may be NULL when no backward pass follows; ws: s2ag_tcn_fused_ws_floats(T, C, dilation)
 * floats, 16-byte aligned (0 = shape not covered: use the two-launch entry).  Whole clips per CTA: T + dilation <= 128. */
long s2ag_tcn_fused_ws_floats(int T, int C, int dilation);
int s2ag_tcn_block_fused_fwd(const float* x, const float* v1, const float* g1, const float* b1, const float* v2,
                             const float* g2, const float* b2, float* w1, float* w2, float* n1, float* n2, float* y1,
                             float* y2, float* out, float* ws, int B, int T, int C, int dilation, float p_drop,
                             uint64_t seed, const uint64_t* seed_dev, void* stream);

/* ---- fused WavEncoder forward (net/multimodal_context_net_v2.py:14-33; frozen inside PoseGeneratorTriModal :277,
 * called at :301) -- raw audio [B, L] -> features y[B, 34, 32] (row stride ldy floats), replacing the chain
 * Conv1d(1,16,15,s5,p1600) BN LReLU Conv1d(16,32,15,s6) BN LReLU Conv1d(32,64,15,s6) BN LReLU Conv1d(64,32,15,s6).
 * conv_w[4] / conv_b[4]: the four Conv1d weights [Cout][Cin][15] / biases; bn_*[3]: the three BatchNorm1d affine
 * parameters (gamma / beta entries may be NULL) and running statistics.  These six arguments are HOST arrays of DEVICE
 * pointers.  training != 0: batch statistics, running statistics updated with `momentum` (nn.BatchNorm1d train mode);
 * 0: running statistics.  slope: LeakyReLU.  ws: caller-owned workspace of s2ag_wavencoder_ws_floats(B, L) floats,
 * 32-byte aligned (raw conv2 / conv3 outputs + statistics; conv1's output is never stored).  4-5 launches (umma_wav.cu). */
long s2ag_wavencoder_ws_floats(int B, int L);
int s2ag_wavencoder_fwd(const float* audio, int B, int L, const float* const* conv_w, const float* const* conv_b,
                        const float* const* bn_gamma, const float* const* bn_beta, float* const* bn_rmean,
                        float* const* bn_rvar, int training, float momentum, float eps, float slope, float* y, long ldy,
                        float* ws, void* stream);

/* ---- ST-GCN adjacency contraction (tgcn.py:67-69: einsum 'nkctv,kvw->nctw') -----------------
 * x[M, V, K*C] (channel index k*C + c), A[K,V,V], y[M, V, C]:  y[m,w,c] = sum_{k,v} x[m,v,k*C+c] A[k,v,w] */
int s2ag_graph_fwd(const float* x, const float* A, float* y, int M, int V, int K, int C, void* stream);
int s2ag_graph_bwd(const float* dy, const float* A, float* dx, int M, int V, int K, int C, void* stream);
/* ConvTemporalGraphical (tgcn.py:15-71) as one temporal convolution: the (Kt x 1) Conv2d Cin -> K*C (weight W[K*C, Cin, Kt],
 * bias b) composed with the adjacency contraction is the Conv1d x[n, t, v*Cin + ci] -> y[n, t, w*C + c] with
 *   Weff[w*C + c][v*Cin + ci][dt] = sum_k A[k,v,w] W[k*C + c][ci][dt],   beff[w*C + c] = sum_k (sum_v A[k,v,w]) b[k*C + c]
 * (run it through s2ag_conv_fwd / _bwd_data / _bwd_weight).  _bwd folds the gradients of the composed tensors back:
 * dW += A-contraction of dWeff, db += ... (db / dbeff may be NULL). */
int s2ag_gcn_compose_fwd(const float* W, const float* b, const float* A, float* Weff, float* beff, int V, int K,
                         int C, int Cin, int Kt, void* stream);
int s2ag_gcn_compose_bwd(const float* dWeff, const float* dbeff, const float* A, float* dW, float* db, int V,
                         int K, int C, int Cin, int Kt, int layout, void* stream);
/* layout 0: dWeff in the convolution's weight layout [V*C][V*Cin][Kt]; layout 1: dWeffT[Kt*V*Cin][V*C] as written by
 * s2ag_window_wgrad.
 * Temporal-convolution weight gradient as a plain contraction over a sliding-window view: x_pad [N, Tp = T + 2p, C] (+ Kt - 1
 * zero rows; s2ag_pad_time with off = p, tail_rows = Kt - 1), dy_pad [N, Tp, Cout] (off = 0: rows t >= T zero), R = N * Tp:
 *   dwT[dt*C + c][co] += sum_r x_pad[(r + dt)*C + c] * dy_pad[r*Cout + co]      (Kwin = Kt*C, ldx = C, lddy = Cout) */
int s2ag_pad_time(const float* src, long ld_src, float* dst, int N, int T, int C, int Tp, int off, int tail_rows,
                  void* stream);
int s2ag_window_wgrad(const float* dy_pad, long lddy, const float* x_pad, long ldx, float* dwT, long R, int Cout,
                      int Kwin, void* stream);
/* db[N] += column sums of dy[M, N] */
int s2ag_colsum(const float* dy, long lddy, float* db, int M, int N, void* stream);

/* ---- weight_norm + causal dilated TemporalBlock (net/tcn.py:16-46) ---------------------------
 * weight_norm (old style, dim=0): w[co] = g[co] * v[co] / ||v[co]||.  v is [Co,Ci,k] (reference
 * layout); the effective weight is emitted TAP-MAJOR as w[Co][k][Ci] so that a tap is a
 * contiguous K-slab for the conv-as-GEMM. norm[Co] is saved for the backward. */
int s2ag_weight_norm_fwd(const float* v, const float* g, float* w, float* norm, int Co, int Ci, int k,
                         void* stream);
/* dv += , dg += from dw[Co][k][Ci] */
int s2ag_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* norm,
                         float* dv, float* dg, int Co, int Ci, int k, void* stream);
/* One residual block over x[B,T,C] (channels-last), kernel size 2, dilation d:
 *   y1 = drop(relu(W1[:,1] x[t] + W1[:,0] x[t-d] + b1)); y2 = drop(relu(conv(y1; W2) + b2));
 *   out = relu(y2 + x).   The chomped columns are never computed.  y1,y2 saved for backward. */
int s2ag_tcn_block_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                       float* y1, float* y2, float* out, int B, int T, int C, int dilation,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream);
/* dx = d(out)/d(x); dw1,dw2 [C][2][C] +=, db1,db2 += ; ws: float[2*B*T*C] scratch */
int s2ag_tcn_block_bwd(const float* dout, const float* x, const float* y1, const float* y2, const float* out,
                       const float* w1, const float* w2, float* dx, float* dw1, float* db1, float* dw2,
                       float* db2, float* ws, int B, int T, int C, int dilation, float p_drop, void* stream);
/* the same in phases: 1 = data path (dx; leaves the g2 / g1 planes in ws), 2 = parameter path (dw1, db1, dw2, db2 from
 * ws), 3 = both.  Phase 2 may run on another stream, ordered after phase 1 by an event (pointers a phase does not use
 * may be NULL). */
int s2ag_tcn_block_bwd_phased(const float* dout, const float* x, const float* y1, const float* y2, const float* out,
                              const float* w1, const float* w2, float* dx, float* dw1, float* db1, float* dw2,
                              float* db2, float* ws, int B, int T, int C, int dilation, float p_drop, int phases,
                              void* stream);

/* ---- nn.Embedding (+ nn.Dropout) (:70-73,:88; :470-472) ------------------------------------- */
int s2ag_embedding_fwd(const int64_t* idx, const float* table, float* out, long ldo, long n, int D, long V,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream);
int s2ag_embedding_bwd(const int64_t* idx, const float* dout, long ldo, float* dtable, long n, int D, long V,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, void* stream);

/* ---- nn.GRU, one bidirectional layer (:480-481,:281-282,:558-560) ---------------------------
 * x[B,T,In] (row stride ldx), weights in the reference layout, gate order r,z,n, h0 = 0.
 * out[B,T,2H] = [forward | reverse].  gi_ws: float[s2ag_gru_fwd_ws_floats(B,T,H)] scratch (input projections
 * + the hidden-state exchange images of the persistent tcgen05 recurrence).  gates (may be NULL when no
 * backward will follow): float[T*2*4*H*B], layout [t][dir][gate r,z,n,(W_hn h + b_hn)][j][b]. */
long s2ag_gru_fwd_ws_floats(int B, int T, int H);
int s2ag_gru_layer_fwd(const float* x, long ldx, const float* w_ih_f, const float* w_ih_r,
                       const float* b_ih_f, const float* b_ih_r, const float* w_hh_f, const float* w_hh_r,
                       const float* b_hh_f, const float* b_hh_r, float* gi_ws, float* out, float* gates,
                       int B, int T, int In, int H, void* stream);
/* The recurrence of one bidirectional layer alone: gi_ws as s2ag_gru_layer_fwd leaves it (the time-batched input
 * projection [B*T][2][3H] followed by the exchange workspace; s2ag_gru_fwd_ws_floats(B,T,H) floats in total).  The
 * latency-bound "GRU step" kernel of the north star, exposed for measurement and for callers that batch the input
 * projection of several passes themselves.  gates may be NULL (inference). */
int s2ag_gru_recurrence_fwd(const float* gi_ws, const float* w_hh_f, const float* w_hh_r, const float* b_hh_f,
                            const float* b_hh_r, float* out, float* gates, int B, int T, int H, void* stream);
/* dout[B,T,*]: gradient of the layer output; direction `d` reads columns d*dir_stride .. +H of a
 * row of stride lddout (dir_stride = H normally; 0 when both halves share the gradient of their
 * sum).  dx (may be NULL) [B,T,In] row stride lddx.  All dw / db += ; pass all eight as NULL to
 * skip the parameter gradients (discriminator inside the generator step).
 * phases (bit mask): 1 = the recurrence (BPTT; fills the dgi/dgh rows in ws), 2 = dx from dgi, 4 = the time-batched weight / bias
 * gradients from dgi/dgh.  7 = everything; a caller may issue 3 on its main stream and 4 on a second stream (ordered after
 * the first call) so that the weight-gradient GEMMs overlap the next layer's latency-bound recurrence.
 * ws: float[s2ag_gru_bwd_ws_floats(B,T,H)] = B*T*6H (dgi) + B*T*6H (dgh) + 4*B*H (dh ping-pong) + the partial-product
 * exchange buffers of the persistent tcgen05 BPTT kernel */
long s2ag_gru_bwd_ws_floats(int B, int T, int H);
int s2ag_gru_layer_bwd(const float* dout, long lddout, int dir_stride, const float* x, long ldx,
                       const float* out, const float* gates,
                       const float* w_ih_f, const float* w_ih_r, const float* w_hh_f, const float* w_hh_r,
                       float* dx, long lddx, float* dw_ih_f, float* dw_ih_r, float* db_ih_f, float* db_ih_r,
                       float* dw_hh_f, float* dw_hh_r, float* db_hh_f, float* db_hh_r, float* ws,
                       int B, int T, int In, int H, int phases, void* stream);

/* ---- speaker style vector (:513-516, :536-539; net/embedding_net.py:10-13) -------------------
 * z = mu + eps * exp(0.5*logvar); z[B,Z] is also tiled over T into dst[B,T,ld] columns [off, off+Z) */
int s2ag_reparam_tile_fwd(const float* mu, const float* logvar, const float* eps, float* z,
                          float* dst, long ld, int off, int B, int T, int Z, void* stream);
/* dz_total = dz (may be NULL) + sum_t ddst[b,t,off:off+Z]; dmu = dz_total; dlogvar = dz_total*eps*0.5*exp(0.5 logvar) */
int s2ag_reparam_tile_bwd(const float* ddst, long ld, int off, const float* dz, const float* logvar,
                          const float* eps, float* dmu, float* dlogvar, int B, int T, int Z, void* stream);

/* ---- discriminator head (:579-585): sigmoid(Linear_T(Linear_H(fwd+rev))) ---------------------
 * g[B,T,2H]; lin1[B,T] saved; out[B] */
int s2ag_dhead_fwd(const float* g, const float* w1, const float* b1, const float* w2, const float* b2,
                   float* lin1, float* out, int B, int T, int H, void* stream);
/* dg[B,T,2H] (both halves get the same gradient); dw1[H] += , db1[1] +=, dw2[T] +=, db2[1] += */
int s2ag_dhead_bwd(const float* dout, const float* out, const float* g, const float* lin1,
                   const float* w1, const float* w2, float* dg, float* dw1, float* db1, float* dw2,
                   float* db2, int B, int T, int H, void* stream);

/* ---- losses (processor_v2.py:811, 893-937, 956) ---------------------------------------------
 * D step: loss[0] = -mean(log(real+1e-8) + log(1-fake+1e-8)); d_real, d_fake [B] gradients */
int s2ag_dis_loss(const float* d_real, const float* d_fake, float* loss, float* g_real, float* g_fake,
                  int B, void* stream);
/* G step.  out,tgt,out_rand [B,TP]; z,z_rand,mu,logvar [B,Z]; dis_out [B] (may be NULL: warm-up).
 * losses[0..4] = huber, gen, kld, div_reg, total (weights w_* applied only in total).
 * Gradients of `total`: g_out[B,TP], g_dis[B], g_mu[B,Z], g_logvar[B,Z]. */
int s2ag_gen_loss(const float* out, const float* tgt, const float* out_rand, const float* z,
                  const float* z_rand, const float* mu, const float* logvar, const float* dis_out,
                  float w_huber, float w_kld, float w_div, float w_gan,
                  float* losses, float* g_out, float* g_dis, float* g_mu, float* g_logvar,
                  int B, int TP, int Z, void* stream);
/* dst[0] = mean |a - b| over n elements */
int s2ag_l1_mean(const float* a, const float* b, float* dst, long n, void* stream);

/* ---- optim.Adam over a flat parameter buffer (processor_v2.py:215-220) -----------------------
 * step_count: device int32, incremented by this call BEFORE use (bias correction reads it on
 * device so that a captured CUDA graph replays correctly). grad_scale multiplies g (1/world). */
int s2ag_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                   float beta2, float eps, float grad_scale, int32_t* step_count, void* stream);

/* ---- attention + softmax over time (net/ser_att_conv_rnn_v2.py:16-34; named off-path) --------
 * x[N,T,Hd]; v = sigmoid(x W1^T + b1) [A]; e = v . w2 + b2; alpha = softmax_t(e); out[n] = sum_t alpha x */
int s2ag_attention_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                       float* out, float* alpha, int N, int T, int Hd, int A, void* stream);

/* backward of s2ag_attention_fwd.  d_out[N,Hd], d_alpha[N,T] (may be NULL) -> dx[N,T,Hd];
 * dw1[A,Hd] +=, db1[A] +=, dw2[A] +=, db2[1] +=.  A <= 64. */
int s2ag_attention_bwd(const float* x, const float* w1, const float* b1, const float* w2, const float* alpha,
                       const float* d_out, const float* d_alpha, float* dx, float* dw1, float* db1,
                       float* dw2, float* db2, int N, int T, int Hd, int A, void* stream);

/* ---- input front-end ---------------------------------------------------
 * utils/common.py:340-349 get_mfcc_features == librosa.feature.mfcc(y, sr, n_mfcc)/1000 (STFT n_fft 2048, hann,
 * centre + reflect padding, |.|^2, slaney mel bank, power_to_db(top_db), DCT-II ortho) and its 1st / 2nd row differences;
 * called per chunk at processor_v2.py:1249-1252.  audio[B, lda >= L]; mel_fb[n_mels, 1025] with mel_span[n_mels][2] =
 * [lo, hi) non-zero bin range of each filter; dct[n_mfcc, n_mels]; logmel_ws[B, F, n_mels] workspace, F = 1 + L/hop;
 * out[B, 3*n_mfcc - 5, F] (the layout PoseGenerator.forward takes); scale = 1/1000. */
int s2ag_mfcc_features(const float* audio, long lda, int B, int L, int hop, const float* mel_fb,
                       const int* mel_span, int n_mels, const float* dct, int n_mfcc, float top_db,
                       float scale, float* logmel_ws, float* out, void* stream);
/* processor_v2.py:606-610 on the device: audio_out[B,L] = int16 audio * audio_max[B] / 32767 (audio_max fp32, or fp64
 * when max_is_f64: same roundings as numpy), mfcc_out[n] = fp16 -> fp32.  Either half may be skipped with NULL. */
int s2ag_expand_inputs(const void* audio_i16, const void* audio_max, int max_is_f64, float* audio_out,
                       long B, long L, const void* mfcc_f16, float* mfcc_out, long n_mfcc, void* stream);

/* ---- long-form synthesis on the device (processor_v2.py:1282-1327, 1334-1391) -----------
 * chunk `chunk` of every clip: out[B,T,P] -> result[b, chunk*(T-n_pre) + t, :] with the first n_pre frames blended
 * linearly with what the previous chunk left there (prev*(n-j)/(n+1) + next*(j+1)/(n+1)); pre_next[B,T,P+1] (may be
 * NULL) = seed of the next chunk: zeros, first n_pre frames = RAW last n_pre frames of out, constraint bit 1.
 * n_chunks[B] (may be NULL): clips with chunk >= n_chunks[b] are left untouched (ragged clip lengths). */
int s2ag_longform_blend(const float* out, float* result, long ld_result, float* pre_next,
                        const int* n_chunks, int chunk, int B, int T, int P, int n_pre, void* stream);
/* fade to the mean pose: seq[b] (ld_seq floats per clip, Lmax frames capacity, len[b] valid frames) is zero-padded to
 * end = start_frame[b] + 2*n_smooth, frames [end - n_smooth, len) zeroed, then [start, end) replaced by its weighted
 * (5,1,...,1,5) quadratic least-squares fit; len_out[b] = max(len[b], end). */
int s2ag_fade_out(float* seq, long ld_seq, const int* len, const int* start_frame, int* len_out, int B, int P,
                  int n_smooth, int Lmax, void* stream);
/* utils/ted_db_utils.py:81-102: vec[N,27] (+ mean[27] when not NULL) -> pose[N,10,3] */
int s2ag_dir_vec_to_pose(const float* vec, const float* mean, float* pose, long N, void* stream);
/* processor_v2.py:738-774 push_samples: dst[0] = L1(out,tgt), dst[1] = joint MAE over frames >= n_pre,
 * dst[2] = mean |accel(tgt) - accel(out)|; out,tgt [B,T,27]; acc_ws: 3 doubles */
int s2ag_pose_metrics(const float* out, const float* tgt, const float* mean, double* acc_ws, float* dst,
                      int B, int T, int n_pre, void* stream);

/* ---- Frechet gesture distance on the device (net/embedding_space_evaluator.py) -----------
 * Replaces the host lists of push_samples (:45-61) and the numpy / scipy get_scores (:73-101): a batch of paired latent
 * features gen_feat / real_feat [N, D] (fp32, row pitches ld_*; D <= 32) is folded into the fp64 moment buffer `acc`
 * (s2ag_fgd_acc_doubles(D) doubles, zeroed by the caller on reset): n, the paired L1 sum, per-dimension sums and the
 * two D x D second-moment matrices. */
long s2ag_fgd_acc_doubles(int D);
int s2ag_fgd_accumulate(const float* gen_feat, long ld_gen, const float* real_feat, long ld_real, int N, int D,
                        double* acc, void* stream);
/* get_scores (:73-101): out2[0] = Frechet distance between N(mean, np.cov) of the generated and of the real features,
 * out2[1] = mean per-sample L1 distance of the paired features ("feat_dist"). */
int s2ag_fgd_scores(const double* acc, int D, double* out2, void* stream);
/* calculate_frechet_distance (:104-152) for caller-given moments (fp64, device): out2[0] = |mu1 - mu2|^2 + Tr(sigma1)
 * + Tr(sigma2) - 2 Tr(sqrtm(sigma1 sigma2)), the trace evaluated as the sum of the square roots of the eigenvalues of
 * sigma1^(1/2) sigma2 sigma1^(1/2) (two fp64 Jacobi eigen-decompositions in one warp); out2[1] = 0. */
int s2ag_frechet_distance(const double* mu1, const double* sigma1, const double* mu2, const double* sigma2, int D,
                          double* out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif
