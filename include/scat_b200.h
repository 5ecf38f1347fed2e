/*
 * scat_b200 C ABI  --  B200 (sm_100a) kernels for SCAT's reg_transformer hand-pose head.
 *
 * The reference (tomguluson92/SCAT) is pure Python/PyTorch and defines no native interface; the
 * boundary its hot path sits behind is the nn.Module `EncoderTransformer` (models/hand_net.py:315-398).
 * This header is the C-ABI a maintainer binds (ctypes) to replace the ATen calls
 * on that path.  Each entry cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t or a negative SCAT_ERR_* code;
 *     nothing throws; scat_last_error_string() describes the last failure on the calling thread.
 *   - all tensor arguments are DEVICE pointers to contiguous fp32 (int32 for indices) unless a
 *     leading dimension is given; the caller owns every buffer including the workspace.
 *   - no function allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so call sequences can be captured into a CUDA graph.
 *   - there is no CPU fallback: without a CUDA device every compute entry fails with a CUDA error.
 */
#ifndef SCAT_B200_H_
#define SCAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCAT_B200_ABI_VERSION 2

#define SCAT_ERR_BAD_ARG     (-1)
#define SCAT_ERR_WORKSPACE   (-2)
#define SCAT_ERR_UNSUPPORTED (-3)

/* arithmetic of the large GEMMs; the regressor and the last feed-forward are always fp32 */
#define SCAT_PREC_FP32 0   /* CUDA-core FFMA everywhere ("parity mode")                         */
#define SCAT_PREC_TF32 1   /* tcgen05 kind::tf32, operands stay fp32 in HBM, fp32 accumulate    */
#define SCAT_PREC_BF16 2   /* tcgen05 kind::f16 with bf16 operands, fp32 accumulate             */
/* scat_gemm only: fp32-grade result from the warp-level tensor instruction (mma.sync TF32) with the 3xTF32 operand
 * split; any strides, no alignment requirement.  Measured slower than the fp32 FFMA kernel on B200 (the legacy
 * mma.sync path has FFMA-class throughput there), so the head does not use it */
#define SCAT_PREC_TF32X3 3

/* dtype of the backbone seam tensors x2 / x2.grad (resnet.py:151 emits fp32, a bf16 /
 * autocast backbone emits bf16): both are NCHW-contiguous */
#define SCAT_DTYPE_F32  0
#define SCAT_DTYPE_BF16 1
/* scat_gemm only, OR-ed into SCAT_PREC_TF32: the fp32 operands are already TF32-representable (their producer
 * rounded them), so the kernel skips its in-shared-memory rounding pass -- how the head itself runs its GEMMs */
#define SCAT_PREC_FLAG_PREROUNDED 0x100
/* scat_gemm only, OR-ed into the precision: the kernel may slice K over CTAs and combine the slices with reductions
 * in L2 (how the head runs its weight-gradient GEMMs); C must be zero on entry, epilogue NONE */
#define SCAT_PREC_FLAG_SPLIT_K 0x200

/* GEMM epilogues (scat_gemm) */
#define SCAT_EPI_NONE       0
#define SCAT_EPI_BIAS       1
#define SCAT_EPI_BIAS_RESID 2
#define SCAT_EPI_BIAS_GELU  3
#define SCAT_EPI_DGELU      4
#define SCAT_EPI_RESID      5

/* Number of parameter tensors of the head, in the reference's state_dict order (hand_net.py:329-353):
 *   0 mask_token[dim]  1 conv1x1_channel_reduction.weight[T,C]
 *   layer i=0,1 (base 2+11i): norm_a.w norm_a.b to_qkv.w[3*64h,d] to_out.w[d,64h] to_out.b
 *                             norm_f.w norm_f.b fc1.w[3d/4,d] fc1.b fc2.w[d/2,3d/4] fc2.b
 *   layer 2 (base 24):        norm_a.w norm_a.b to_qkv.w to_out.w to_out.b fc1.w fc1.b fc2.w[3,hid] fc2.b
 *   33 regressor.weight[66,1090]  34 regressor.bias[66] */
#define SCAT_HEAD_NUM_PARAMS 35

typedef struct ScatHeadDesc {
    int32_t batch;          /* B (may vary call to call: train.py:143-150 filters empty images)     */
    int32_t n_tokens;       /* 21  (hand_net.py:328)                                                */
    int32_t channels;       /* 512 (x2 channels, hand_net.py:329); 0 = token input, no conv         */
    int32_t token_dim;      /* 784 = 28*28 (hand_net.py:331)                                        */
    int32_t heads;          /* opt.vit_heads                                                        */
    int32_t iteration;      /* opt.iteration; 0 with main_feat_dim=0 = no regressor (token path)    */
    int32_t pos_embed;      /* opt.pos_embed                                                        */
    int32_t n_masked;       /* int(mask_rate*n_tokens) if 0.1<=mask_rate<=0.9 else 0                */
    int32_t pl_reg;         /* opt.pl_reg: also return d(sum feat_out)/d feat_visual                */
    int32_t precision;      /* SCAT_PREC_*                                                          */
    int32_t main_feat_dim;  /* 1024 (resnet fc1)                                                    */
    int32_t n_out;          /* 66 = 3 camera + 63 joint coordinates                                 */
    int32_t x2_dtype;       /* SCAT_DTYPE_*: storage of x2 AND x2_grad; BF16 needs precision tf32/bf16      */
} ScatHeadDesc;

int         scat_abi_version(void);
const char* scat_last_error_string(void);
/* number of CUDA kernels this library has launched (or captured into a graph) since load */
uint64_t    scat_launch_count(void);

/* ---- whole-head entry points ---------------------------------------------------------------- */

/* bytes of caller-owned workspace for scat_head_forward/backward/train_step with this descriptor */
size_t scat_head_workspace_bytes(const ScatHeadDesc* desc);

/* EncoderTransformer.forward after the backbone (hand_net.py:363-398).
 *   params[35]     device pointers, order above;  pe[T,dim] = positionalEncoding.pe[0];
 *   mean_params[66]; mask_idx[n_masked] int32 token indices drawn on the host (hand_net.py:370-372)
 *   x2[B,C,28,28] (fp32 or bf16 per desc.x2_dtype), main_feat[B,1024]
 *     ->  pred_params[B,66], feat_visual[B,T,28,28], pl_term[B,T,28,28]
 *   (pl_term may be NULL iff !pl_reg).  With pos_embed==0 and masking the reference overwrites
 *   feat_visual in place (hand_net.py:364,373); that aliasing is reproduced. */
int scat_head_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                      const float* mean_params, const int32_t* mask_idx, const void* x2,
                      const float* main_feat, float* pred_params, float* feat_visual, float* pl_term,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above (what loss.backward() does at train.py:206) from the forward's workspace.
 *   feat_visual: the forward's output (only read when pos_embed==0, where it IS the token matrix)
 *   grad_pred[B,66] (required), grad_feat_visual[B,T,784] (NULL = zero)
 *   grads[35] device pointers to receive parameter gradients (overwritten, not accumulated),
 *   x2_grad[B,C,784] (dtype desc.x2_dtype) / main_feat_grad[B,1024]: NULL to skip. */
int scat_head_backward(const ScatHeadDesc* desc, const float* const* params, const int32_t* mask_idx,
                       const void* x2, const float* main_feat, const float* feat_visual,
                       const float* grad_pred, const float* grad_feat_visual, float* const* grads,
                       void* x2_grad, float* main_feat_grad, void* workspace, size_t workspace_bytes,
                       void* stream);

/* projection + losses + d loss / d pred_params (train.py:112-120,165-203).
 *   labels[B,ld_labels]: ld_labels = 105 -> [63 3D | 42 2D]; 166 -> [61 pose | 63 3D | 42 2D] (train.py:188-199 picks
 *   the columns by the row width; any other width is rejected), pl_term NULL = no path-length term,
 *   losses[4] = {loss, l_3d, l_2d, l_pl} (device), grad_pred[B,66] scaled by grad_scale (1/world for DP),
 *   scratch: >= B floats. */
int scat_proj_loss(int32_t batch, int32_t n_tokens, int32_t token_dim, const float* pred_params,
                   const float* labels, int32_t ld_labels, const float* pl_term, float l_weight_3d,
                   float l_weight_2d, float grad_scale, float* losses, float* grad_pred, float* scratch,
                   void* stream);

/* One fused training-step body (train.py:159-206): forward, path-length VJP, losses, backward. */
int scat_head_train_step(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                         const float* mean_params, const int32_t* mask_idx, const void* x2,
                         const float* main_feat, const float* labels, int32_t ld_labels,
                         float l_weight_3d, float l_weight_2d, float grad_scale, float* pred_params,
                         float* feat_visual, float* pl_term, float* losses, float* const* grads,
                         void* x2_grad, float* main_feat_grad, void* workspace, size_t workspace_bytes,
                         void* stream);

/* The same step in three launches for data-parallel callers, so the gradient all-reduce overlaps the backward:
 * phase 0 = forward, losses and the backward down to transformer layer 1 (the gradients of layers 1, 2 and of the
 * regressor -- parameters 13..34 of the gradient list -- are final when it returns); phase 1 = layer 0 and the token
 * masking (parameters 0 and 2..12 and the losses are final); phase 2 = conv backward (parameter 1, x2_grad).
 * phase -1 = scat_head_train_step. */
int scat_head_train_step_phase(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                               const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                               const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d,
                               float grad_scale, float* pred_params, float* feat_visual, float* pl_term, float* losses,
                               float* const* grads, void* x2_grad, float* main_feat_grad, void* workspace,
                               size_t workspace_bytes, void* stream, int32_t phase);

/* The single-call step with a gradients-ready hook, for data-parallel callers that hide their gradient exchange under the
 * backward without cutting the step into phases.  `ready(user, part, side_stream)` is called on the host, while the step
 * is being ENQUEUED (also under stream capture), four times and in this order:
 *   part 0: parameters 13..34 (transformer layers 1, 2 and the regressor) are final in the order of `side_stream`;
 *   part 1: parameters 7..12 (the feed-forward half of transformer layer 0: its LayerNorm, fc1, fc2);
 *   part 2: parameters 2..6 (the attention half of layer 0: its LayerNorm, to_qkv, to_out);
 *   part 3: parameters 0, 1 (mask token, conv weight) -- the conv data gradient (x2_grad) is still to come.
 * Work the hook enqueues on `side_stream` (a library-owned stream, the same for the four calls, so the hook's launches
 * are serialised among themselves) runs beside the rest of the step; `stream` waits for it before the call returns, so
 * the step stays one capturable unit.  A non-zero return from the hook fails the call.  ready == NULL: scat_head_train_step. */
typedef int (*scat_grads_ready_fn)(void* user, int32_t part, void* side_stream);
int scat_head_train_step_hooked(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                                const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                                const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d,
                                float grad_scale, float* pred_params, float* feat_visual, float* pl_term, float* losses,
                                float* const* grads, void* x2_grad, float* main_feat_grad, void* workspace,
                                size_t workspace_bytes, scat_grads_ready_fn ready, void* ready_user, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (the reference imports
 * DistributedDataParallel at train.py:18 and never uses it, so this has no reference counterpart to replace) ----
 * One process per GPU.  Each rank allocates its gradient bucket and a signal area with scat_peer_alloc, exports both
 * (64-byte CUDA IPC handles, exchanged by the host through any channel), and maps every peer's with scat_peer_open.
 * scat_peer_allreduce then sums elements [lo, hi) (multiples of 4) of all buckets in place, in rank order, in ONE
 * kernel on `stream` (capturable): buckets[p] / signals[p] are rank p's bucket and signal area as mapped in this
 * process (own entries = the local allocations).  Every rank must issue the same sequence of calls.
 * A peer that never arrives does not hang the GPU: after 20 s the kernel gives up WITHOUT summing or storing anything,
 * scat_peer_error reports 1 from then on, and every later exchange on this rank returns immediately. */
size_t scat_peer_signal_bytes(void);
int scat_peer_alloc(size_t bytes, void** out);                 /* zero-filled device memory, IPC-exportable */
int scat_peer_free(void* ptr);
int scat_peer_export(void* ptr, uint8_t* handle64);
int scat_peer_open(const uint8_t* handle64, void** out);
int scat_peer_close(void* ptr);
int scat_peer_allreduce(float* const* buckets, uint32_t* const* signals, int32_t rank, int32_t world, long long lo,
                        long long hi, void* stream);
/* One of SEVERAL exchanges a step issues on the same stream (disjoint ranges, same order on every rank): with last == 0
 * the closing cross-GPU barrier is left to a later exchange of that stream -- the sums of this range may only be read,
 * and the gradients only be rewritten, after an exchange with last != 0 (or scat_peer_allreduce) on the same stream. */
int scat_peer_allreduce_part(float* const* buckets, uint32_t* const* signals, int32_t rank, int32_t world, long long lo,
                             long long hi, int32_t last, void* stream);
int scat_peer_error(const uint32_t* signal, int32_t* out);     /* synchronous read of the time-out flag */
/* device address of that flag inside a signal area: non-zero once an exchange on this rank has given up on a peer (it
 * stays set; later exchanges are then no-ops).  Pass it to scat_adam_step as abort_flag. */
const uint32_t* scat_peer_error_word(const uint32_t* signal);

/* Fused Adam step over the head's flat parameter / gradient / moment buffers (all in the gradient bucket's order):
 * replaces optim.Adam(self.net.parameters(), lr).step() (train.py:60,209) for the head's 35 tensors, arithmetic as
 * torch.optim.Adam's single-tensor path (weight_decay is the coupled L2 form, no amsgrad).  `step` is the 1-based
 * count of this update; if `step_dev` / `lr_dev` are non-null the kernel reads the count / learning rate from device
 * memory instead (a captured CUDA graph then follows a changing schedule).  Buffers 16-byte aligned.  The scalar
 * hyper-parameters are doubles, as torch holds them: 1 - beta is formed in double before it is rounded to fp32.
 * abort_flag (nullable, device): when the word it points to is non-zero the kernel changes nothing -- wire it to
 * scat_peer_error_word() so that a gradient exchange that gave up on a peer never reaches the weights. */
int scat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int32_t step, const float* lr_dev,
                   const int32_t* step_dev, const uint32_t* abort_flag, void* stream);

/* ---- evaluation metrics on the device (the reference round-trips every frame to numpy,
 * eval.py:691-753).  All joints are fp32 [batch, n_joints, 3]. ----
 * scat_eval_procrustes: batch_compute_similarity_transform_torch (eval.py:110-161): per sample the similarity transform
 *   (scale, rotation via 3x3 SVD with the det fix, translation) that takes `pred` closest to `gt`; writes the transformed
 *   prediction to `aligned` (may alias pred) and the scale to `scale` (nullable).
 * scat_eval_joint_errors: the device part of cal_PCK (eval.py:300-316) and of the MPJPE (eval.py:749):
 *   d = || unit_scale*pred - unit_scale*gt || per joint (unit_scale 1000: metres -> millimetres, as the reference);
 *   counts[k] = number of joints of the whole batch with d <= thresholds[k] (thresholds: HOST array of n_thresholds <= 64
 *   doubles; counts: device, zeroed by the call) -- PCK_k = 100 * counts[k] / (batch * n_joints); mpjpe[b] (nullable) =
 *   mean over the joints of || pred - gt ||.
 * scat_eval_accel: compute_error_accel (data_utils/eval_utils.py:20-47) before its visibility selection:
 *   out[i] = mean over joints of || (pred[i] - 2 pred[i+1] + pred[i+2]) - (gt[i] - 2 gt[i+1] + gt[i+2]) ||, i < n_frames-2;
 *   gt == NULL gives compute_accel (eval_utils.py:6-17). */
int scat_eval_procrustes(const float* pred, const float* gt, int32_t batch, int32_t n_joints, float* aligned, float* scale,
                         void* stream);
int scat_eval_joint_errors(const float* pred, const float* gt, int32_t batch, int32_t n_joints, float unit_scale,
                           const double* thresholds, int32_t n_thresholds, unsigned long long* counts, float* mpjpe,
                           void* stream);
int scat_eval_accel(const float* pred, const float* gt, int32_t n_frames, int32_t n_joints, float* out, void* stream);

/* Token-only transformer (HRNet-variant path up to feat.mean(dim=1), hand_net.py:193-203):
 *   tokens[B,n,dim] -> out[B,n,3], mean[B,3].  desc.channels = 0, desc.iteration = 0. */
int scat_tokens_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                        const int32_t* mask_idx, const float* tokens, float* out, float* mean,
                        void* workspace, size_t workspace_bytes, void* stream);

/* EncoderTransformerCoarse.forward after the backbone (hand_net.py:264-311, `--net reg_transformer_coarse`, the
 * attention-visualisation variant used by eval.py:788-834): same conv + PE + masking front end, the transformer of
 * models/vision_transformer_attn.py:104-113 (attention on the raw tokens, LayerNorm on the attention branch's output, then
 * the residual), camera = Linear(1027 -> 3) applied once, joints = mean template + transformer output relative to joint 1.
 *   params[35]: the head's layout above with slots norm_a.w / norm_a.b holding layers.i.1.norm (the post-attention norm),
 *   norm_f / fc1 / fc2 holding layers.i.2.*, and 33 / 34 = regressor.weight[3,1027] / regressor.bias[3].
 *   -> pred_params[B,66], feat_visual[B,21,28,28], attn[B,heads,21,21] (last layer's attention maps).
 * Inference only (desc.pl_reg must be 0: the reference itself fails with pl_reg under no_grad); precision fp32 or tf32. */
int scat_coarse_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe, const float* mean_params,
                        const int32_t* mask_idx, const void* x2, const float* main_feat, float* pred_params,
                        float* feat_visual, float* attn, void* workspace, size_t workspace_bytes, void* stream);

/* ---- single operators (used by the unit tests and by other callers of the same kernels) ------ */

/* C[M,N] = epilogue(sum_k A(m,k) B(n,k)),  A(m,k)=A[m*sam+k*sak], B(n,k)=B[n*sbn+k*sbk]
 * replaces nn.Linear forward / backward (vision_transformer.py:33-35,53-55). */
int scat_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C,
              int32_t ldc, int32_t M, int32_t N, int32_t K, int32_t epilogue, const float* bias,
              const float* aux_in, int32_t ld_aux_in, float* aux_out, int32_t ld_aux_out,
              int32_t accumulate, int32_t precision, void* stream);

/* Same contraction with bf16 operands already in HBM (tcgen05 kind::f16, fp32 accumulation in TMEM): the GEMM the
 * head runs in SCAT_PREC_BF16 mode, where every producer kernel stores the bf16 copy its consumer GEMM reads.
 * A, B: bf16 (uint16 storage), strides in elements, unit stride in one direction, 16-byte aligned rows.
 * C (fp32, NULL to skip) and/or C16 (bf16 copy, NULL to skip) receive epilogue(A B^T).  split_k != 0 lets the
 * kernel slice K over CTAs and combine with reductions in L2 (C must then be zero on entry, epilogue NONE). */
int scat_gemm_bf16(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk, float* C,
                   int32_t ldc, void* C16, int32_t ldc16, int32_t M, int32_t N, int32_t K, int32_t epilogue,
                   const float* bias, const float* aux_in, int32_t ld_aux_in, float* aux_out, int32_t ld_aux_out,
                   int32_t split_k, void* stream);

/* Profiling hook: when set to a device buffer of 8 int64, CTA (0,0,0) of every tensor-core GEMM launch records
 * clock64() at entry / setup done / dependency wait done / first stage landed / epilogue waiting / accumulator
 * ready / stores issued / teardown.  NULL (default) disables it.  See tools/gemm_timeline.py. */
void scat_debug_gemm_timeline(void* dev_int64x8);

/* hand_net.py:363-373 */
int scat_conv_pe_mask_fwd(const float* x2, const float* conv_w, const float* pe, const float* mask_token,
                          const int32_t* mask_idx, int32_t n_masked, int32_t pos_embed, float* feat_visual,
                          float* tokens_out, int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens,
                          void* stream);
/* The same front end on the tensor cores (csrc/conv_tc.cu: three persistent tcgen05 kernels that stream x2 / x2.grad
 * through a TMA ring, accumulators in tensor memory): what the head runs in SCAT_PREC_TF32 / SCAT_PREC_BF16.
 * x2_dtype SCAT_DTYPE_F32: x2 and x2_grad are fp32; x2 is rounded to TF32-NEAREST in shared memory (the tensor core
 * alone would truncate), results are TF32-grade (~3e-4 relative) for feat_visual / conv_w_grad and fp32-grade for x2_grad
 * (exact hi/lo split of d_tokens and the weight).  SCAT_DTYPE_BF16: x2 and x2_grad are bf16; bf16 operands are exact on
 * the tensor core and the fp32 weight / d_tokens enter as stacked bf16 split terms, so feat_visual and conv_w_grad are
 * fp32-grade for the given bf16 x2, and x2_grad carries only its final bf16 rounding.  Masked rows stay bit-exact
 * copies of the mask token.  T = 21, C in {256, 512}, hw % 8 == 0.  scratch: scat_conv_tc_scratch_floats() floats. */
size_t scat_conv_tc_scratch_floats(int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens);
int scat_conv_pe_mask_fwd_tc(const void* x2, int32_t x2_dtype, const float* conv_w, const float* pe, const float* mask_token,
                             const int32_t* mask_idx, int32_t n_masked, int32_t pos_embed, float* feat_visual,
                             float* tokens_out, float* scratch, int32_t batch, int32_t channels, int32_t hw,
                             int32_t n_tokens, void* stream);
int scat_conv_bwd_tc(const float* d_tokens, const void* x2, int32_t x2_dtype, const float* conv_w, const int32_t* mask_idx,
                     int32_t n_masked, void* x2_grad, float* conv_w_grad, float* mask_token_grad, float* scratch,
                     int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens, void* stream);

/* backward of the conv: d_tokens[B,T,hw] -> x2_grad (NULL to skip), conv_w_grad[T,C], mask_token_grad[hw];
 * scratch: scat_conv_bwd_scratch_floats() floats */
size_t scat_conv_bwd_scratch_floats(int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens);
int scat_conv_bwd(const float* d_tokens, const float* x2, const float* conv_w, const int32_t* mask_idx,
                  int32_t n_masked, float* x2_grad, float* conv_w_grad, float* mask_token_grad, float* scratch,
                  int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens, void* stream);

/* vision_transformer.py:23,26 */
int scat_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                       int32_t rows, int32_t dim, void* stream);
int scat_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* resid, float* dx, float* dgamma, float* dbeta, int32_t rows, int32_t dim,
                       void* stream);

/* vision_transformer.py:61-77: qkv[B*n,3*64h] -> o[B*n,64h], p[B,h,n,n] */
int scat_attention_fwd(const float* qkv, float* o, float* p, int32_t batch, int32_t n, int32_t heads, void* stream);
int scat_attention_bwd(const float* qkv, const float* p, const float* d_o, float* d_qkv, int32_t batch, int32_t n,
                       int32_t heads, void* stream);
/* The same for n = 21 tokens on the tensor cores (mma.sync m16n8k8 TF32, one warp per (batch, head) problem, operands
 * rounded to TF32-nearest): what the head runs in SCAT_PREC_TF32 / SCAT_PREC_BF16.  TF32-grade results.
 * scat_attention_fwd_tc also takes n = 128 (the HRNet-token variant, inference): one tcgen05 tile per (batch, head)
 * problem, S and O accumulated in tensor memory; p is then left untouched. */
int scat_attention_fwd_tc(const float* qkv, float* o, float* p, int32_t batch, int32_t n, int32_t heads, void* stream);
int scat_attention_bwd_tc(const float* qkv, const float* p, const float* d_o, float* d_qkv, int32_t batch, int32_t n,
                          int32_t heads, void* stream);

/* hand_net.py:379-393 (root_relative=1, n_out=66) and hand_net.py:53-57 (root_relative=0, n_out=61).
 * states[B,iteration,n_out] may be NULL when no backward follows; scratch: batch*n_out floats. */
int scat_regressor_fwd(const float* main_feat, const float* feat_out, const float* mean_params, const float* w,
                       const float* b, float* pred, float* states, float* scratch, int32_t batch, int32_t feat_dim,
                       int32_t n_out, int32_t iteration, int32_t root_relative, void* stream);

/* MANO linear blend skinning, models/mano.py:280-391 (rot_pose_beta_to_mesh).
 * asset (as in MANO_RIGHT.pkl, row-major fp32): v_template[778,3] shapedirs[778,3,10] posedirs[778,3,135]
 * J_regressor[16,778] (dense) weights[778,16] hands_mean[45].
 * scat_lbs_prepare fills derived[scat_lbs_derived_floats()] once per asset (regressed template joints and
 * vertex-contiguous copies of the blend-shape tables); scat_lbs_fwd then maps
 * rots[B,3] poses[B,45] betas[B,10] -> out[B,799,3] (21 joints then 778 vertices, joint 1 at the origin). */
size_t scat_lbs_derived_floats(void);
int scat_lbs_prepare(const float* v_template, const float* shapedirs, const float* posedirs,
                     const float* j_regressor, const float* weights, float* derived, void* stream);
int scat_lbs_fwd(const float* derived, const float* hands_mean, const float* rots, const float* poses,
                 const float* betas, float* out, int32_t batch, void* stream);
/* The same forward with the two blend-shape contractions (0.65 of the 1.19 MFLOP per sample) on the tensor cores: one
 * tcgen05 GEMM [B,145] x [145,2334] at fp32 grade (TF32 hi/lo split of both operands stacked along K), then a skinning
 * kernel.  table: scat_lbs_tc_table_floats() floats filled once per asset by scat_lbs_tc_prepare; scratch:
 * scat_lbs_tc_scratch_floats(batch) floats (a smaller scratch just means more, smaller chunks; a chunk's intermediates are
 * sized to stay in L2).  Same output as scat_lbs_fwd to ~1e-7 m. */
size_t scat_lbs_tc_table_floats(void);
size_t scat_lbs_tc_scratch_floats(int32_t batch);
int scat_lbs_tc_prepare(const float* shapedirs, const float* posedirs, float* table, void* stream);
int scat_lbs_fwd_tc(const float* derived, const float* table, const float* hands_mean, const float* rots, const float* poses,
                    const float* betas, float* out, int32_t batch, float* scratch, size_t scratch_floats, void* stream);
/* Backward of scat_lbs_fwd: mano.py:280-391 is plain differentiable PyTorch in the reference (autograd); this is its
 * vector-Jacobian product.  grad_out[B,799,3] -> grad_rots[B,3], grad_poses[B,45], grad_betas[B,10] (overwritten).
 * Nothing is saved by the forward: the kernel recomputes from (rots, poses, betas).  Rows whose axis-angle norm is
 * below 1e-30 take the gradient of the Taylor branch (mano.py:258-265), where the reference's autograd returns NaN. */
int scat_lbs_bwd(const float* derived, const float* hands_mean, const float* rots, const float* poses,
                 const float* betas, const float* grad_out, float* grad_rots, float* grad_poses,
                 float* grad_betas, int32_t batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCAT_B200_H_ */
