// Internal host-side launcher API shared by the .cu translation units.  Not part of the C ABI
// (see include/scat_b200.h for that); everything here takes raw device pointers and a stream and
// never allocates or synchronises, so a caller may capture any sequence into a CUDA graph.
#pragma once
#include "common.cuh"

namespace scat {

// ------------------------------------------------------------------------------------------
// GEMM:  C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )   with fully strided operands
//   A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
//   forward  y = x W^T      : A=x (sam=ldx,sak=1)    B=W  (sbn=ldw,sbk=1)      "K-major / K-major"
//   dgrad    dx = dy W      : A=dy (sam=ld,sak=1)    B=W  (sbn=1,  sbk=ldw)    "K-major / MN-major"
//   wgrad    dW = dy^T x    : A=dy (sam=1, sak=ld)   B=x  (sbn=1,  sbk=ldx)    "MN-major / MN-major"
// ------------------------------------------------------------------------------------------
enum Epilogue : int {
    EPI_NONE = 0,        // C = acc
    EPI_BIAS = 1,        // C = acc + bias[n]
    EPI_BIAS_RESID = 2,  // C = acc + bias[n] + aux_in[m,n]
    EPI_BIAS_GELU = 3,   // z = acc + bias[n]; aux_out[m,n] = z; C = gelu(z)
    EPI_DGELU = 4,       // C = acc * gelu'(aux_in[m,n])
    EPI_RESID = 5,       // C = acc + aux_in[m,n]
    EPI_PE_MASK = 6,     // tensor-core kernel only, rows = tokens (conv front end, hand_net.py:363-373):
                         //   tok = row in mask_idx ? bias[n] (mask token) : acc (+ aux_in[m,n] = positional encoding);
                         //   aux_out set: C = acc, aux_out = tok;   aux_out null (token matrix aliases C): C = tok
};

enum Precision : int {
    PREC_FP32 = 0,  // CUDA-core FFMA, fp32 throughout (parity mode)
    PREC_TF32 = 1,  // tcgen05 kind::tf32, fp32 storage, fp32 accumulate in TMEM
    PREC_BF16 = 2,  // tcgen05 kind::f16 (bf16 operands), fp32 accumulate
};

struct GemmArgs {
    const void* A = nullptr; long long sam = 0, sak = 0;     // fp32, or bf16 when operand_bf16 (strides in elements)
    const void* B = nullptr; long long sbn = 0, sbk = 0;
    int operand_bf16 = 0;    // tensor-core kernel only: A and B hold bf16 (tcgen05 kind::f16), else fp32 (kind::tf32)
    float* C = nullptr; int ldc = 0;                         // fp32 output (tensor-core kernel: may be null if C16 is set)
    void* C16 = nullptr; int ldc16 = 0;                      // tensor-core kernel only: bf16 copy of the output
    int M = 0, N = 0, K = 0;
    int epilogue = EPI_NONE;
    const float* bias = nullptr;
    const float* aux_in = nullptr; int ld_aux_in = 0;
    int aux_row_mod = 0;     // > 0: aux_in row = m % aux_row_mod (stacked cotangents share one saved activation)
    float* aux_out = nullptr; int ld_aux_out = 0;
    int accumulate = 0;  // C += epilogue(...)
    int c_zeroed = 0;        // C is known to be zero: split-K slices may reduce into it without a memset
    int allow_split_k = 0;   // FFMA kernel may split K over CTAs and combine with atomics (weight gradients)
    int prerounded = 0;      // tensor-core kernel: operands are already TF32-representable, skip the in-kernel rounding
    int round_out = 0;       // store C rounded to TF32 (nearest): it feeds a tensor-core GEMM next
    // ---- tensor-core kernel only: a stack of `batch` problems in one launch (blockIdx.z) ----
    int batch = 1;
    int a_row_z = 0, a_k_z = 0, b_row_z = 0, b_k_z = 0;   // per-problem offsets of the operands along rows / k (elements)
    long long c_z = 0, aux_out_z = 0;                    // per-problem element offsets of C / aux_out
    int batch_accumulate = 0;    // every problem reduces into the same C (red.global.add; C pre-zeroed): K split over the batch
    const int32_t* mask_idx = nullptr; int n_masked = 0;   // EPI_PE_MASK
    int gelu_saves_grad = 0;     // EPI_BIAS_GELU: aux_out receives gelu'(z) instead of z;  EPI_DGELU: aux_in already IS gelu'(z)
                                 // (the head: the backward's epilogue becomes a multiply, the erf is evaluated once, in the forward)
    int force_bn = 0;            // 64 / 128: tile width override
    int allow_wide = 0;          // tensor-core kernel: 192- / 256-wide single-wave tiles may be chosen (one 200 KB CTA per SM:
                                 // only where nothing runs beside this GEMM -- measured slower in the backward, where the
                                 // weight-gradient GEMMs of the side stream then cannot share the SMs)
    int b_static = 0;            // tensor-core kernel: B was NOT written by the kernel preceding this launch on its stream (a
                                 // weight, a saved activation): its first stages are requested before the PDL dependency wait
};

#ifdef __CUDACC__
// fp32 -> nearest TF32-representable fp32 (10-bit mantissa).  The tensor core truncates, so operands are rounded
// by whoever produces them (or inside the GEMM kernel when they are not).
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
#endif

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);
// fp32-grade small GEMM on mma.sync TF32 with the 3xTF32 split (gemm_mma3.cu): any strides, epilogues NONE..RESID
int launch_gemm_mma3(const GemmArgs& g, cudaStream_t stream);
// tcgen05 path; returns kErrUnsupported if the operand strides cannot be expressed as TMA tensor maps
int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream);
bool gemm_tc_supported(const GemmArgs& g);
void gemm_tc_set_debug_buffer(long long* dev8);   // 8 x int64 device buffer, or null (profiling hook)
int launch_gemm(const GemmArgs& g, int precision, cudaStream_t stream);  // dispatch on precision/shape

// column sums: out[n] (+)= sum_m X[m*ld + n]
int launch_colsum(const float* X, int ld, int M, int N, float* out, int accumulate, cudaStream_t stream,
                  int x_bf16 = 0);   // x_bf16: X holds bf16 (ld in elements)

// ------------------------------------------------------------------------------------------
// 1x1 conv + positional encoding + token masking (hand_net.py:363-373)
// ------------------------------------------------------------------------------------------
int launch_conv_pe_mask_fwd(const float* x2, const float* Wc, const float* pe, const float* mask_token,
                            const int32_t* mask_idx, int n_masked, int pos_embed, float* feat_visual, float* X0,
                            int B, int C, int HW, int T, cudaStream_t stream);
// dFv = dX0 with masked token rows zeroed (unless keep_masked); d mask_token = sum over b, masked t of dX0
int launch_mask_bwd(const float* dX0, const int32_t* mask_idx, int n_masked, int keep_masked, float* dFv,
                    float* d_mask_token, int B, int T, int HW, cudaStream_t stream, int token_grad_zeroed = 0);
int launch_conv_dgrad(const float* dFv, const float* Wc, float* x2_grad, int B, int C, int HW, int T,
                      cudaStream_t stream);
size_t conv_wgrad_scratch_floats(int C, int T);
int launch_conv_wgrad(const float* dFv, const float* x2, float* dWc, float* scratch, int B, int C, int HW, int T,
                      cudaStream_t stream);
// Tensor-core versions (conv_tc.cu): three persistent tcgen05 kernels, HBM streams of x2 / x2.grad.  x2_bf16 selects
// the seam dtype: 0 = fp32 x2 / x2.grad (kind::tf32, x2 rounded to TF32-nearest in shared memory), 1 = bf16 x2 / x2.grad
// (kind::f16, exact operands, the fp32 weight and d tokens enter as stacked bf16 split terms).
//   Wprep  = launch_conv_weight_prep(conv weight)   conv_weight_prep_floats(C, T) floats, once per forward
//   dsplit = launch_conv_bwd_prep(d token matrix [B,T,HW])   conv_split_floats(B, HW, T) floats; zeroes the masked token
//            rows on the way and reduces them into d mask_token (nullable; zero on entry)
size_t conv_weight_prep_floats(int C, int T);
size_t conv_split_floats(int B, int HW, int T);
int launch_conv_weight_prep(const float* Wc, void* Wprep, int C, int T, int x2_bf16, cudaStream_t stream);
int launch_conv_bwd_prep(const float* dX0, const int32_t* mask_idx, int n_masked, void* dsplit, float* d_mask_token, int B,
                         int T, int HW, int x2_bf16, cudaStream_t stream);
int launch_conv_pe_mask_fwd_tc(const void* x2, int x2_bf16, const void* Wprep, const float* pe, const float* mask_token,
                               const int32_t* mask_idx, int n_masked, int pos_embed, float* feat_visual, float* X0,
                               int B, int C, int HW, int T, cudaStream_t stream);
int launch_conv_dgrad_tc(const void* dsplit, const void* Wprep, int x2_bf16, void* x2_grad, int B, int C, int HW, int T,
                         cudaStream_t stream);
int launch_conv_wgrad_tc(const void* dsplit, const void* x2, int x2_bf16, float* dWc /* pre-zeroed */, int B, int C, int HW,
                         int T, cudaStream_t stream);
// token-only front end (config 4): X0 = tokens (+pe) with masked rows replaced
int launch_pe_mask_tokens(const float* tokens, const float* pe, const float* mask_token, const int32_t* mask_idx,
                          int n_masked, int pos_embed, float* X0, int B, int T, int D, cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// LayerNorm (vision_transformer.py:23,26; eps 1e-5, affine)
// ------------------------------------------------------------------------------------------
// resid (nullable): Y = LayerNorm(X) + resid
int launch_layernorm_fwd(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                         float* mean, float* rstd, int M, int D, int round_out, cudaStream_t stream,
                         const float* resid = nullptr, int ldr = 0);
// dX = LN'(dY) (+ resid); dgamma/dbeta accumulated with atomics when non-null (must be pre-zeroed)
int launch_layernorm_bwd(const float* dY, int lddy, const float* X, int ldx, const float* gamma, const float* mean,
                         const float* rstd, const float* resid, int ldr, float* dX, int lddx, float* dgamma,
                         float* dbeta, int M, int D, int round_out, cudaStream_t stream, int act_rows = 0,
                         void* dX16 = nullptr, int lddx16 = 0);   // dX16: optional bf16 shadow copy of dX
// the parameter gradients alone, as column sums (dgamma / dbeta zero on entry): what the head runs on its side stream
// next to launch_layernorm_bwd(..., dgamma = dbeta = nullptr, ...) on the critical path
// E / esum (optional): a same-shaped [M, D] tensor whose column sums (a Linear bias gradient) accumulate into esum
int launch_layernorm_param_grads(const float* dY, int lddy, const float* X, int ldx, const float* mean, const float* rstd,
                                 float* dgamma, float* dbeta, int M, int D, cudaStream_t stream, const float* E = nullptr,
                                 int lde = 0, float* esum = nullptr);
// act_rows > 0: dY/dX/resid have M rows, the saved activations (X, mean, rstd) have act_rows rows and row m uses
// activation row m % act_rows; dgamma/dbeta only accumulate rows m < act_rows (the real cotangent)

// ------------------------------------------------------------------------------------------
// softmax attention over n tokens, heads of 64 (vision_transformer.py:61-77)
//   QKV [B*n, 3*inner] (q | k | v, head g = columns g*64..g*64+63), O [B*n, inner], P [B,h,n,n]
// ------------------------------------------------------------------------------------------
int launch_attention_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int round_out,
                         cudaStream_t stream);
int launch_attention_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                         int round_out, cudaStream_t stream, int act_batch = 0);
// act_batch > 0: dO/dQKV have B samples, QKV/P have act_batch samples and sample b uses activations of b % act_batch
// the n = 21 tensor-core kernels (mma.sync TF32) directly, any out_mode (attention_mma.cu)
int launch_attention_mma_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int out_mode, cudaStream_t stream);
int launch_attention_mma_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                             int out_mode, cudaStream_t stream, int act_batch = 0);
// n = 128 on tcgen05 (attention_tc128.cu), forward only, P not written
int launch_attention_tc128_fwd(const float* QKV, float* O, int B, int n, int heads, int out_mode, cudaStream_t stream);
// dst[r, c] = round_tf32(src[r, c]) (pad columns zero-filled) for a list of weight matrices, one launch;
// the job table travels by value as a kernel argument (no device-side table, graph-capturable)
// mode 0: dst = TF32-nearest(src) as fp32; 1: dst = bf16(src) (dst is a bf16 array); 2: dst = TF32-nearest(src - TF32-nearest(src))
struct RoundJob { const float* src; float* dst; int rows, cols, ld_src, ld_dst; int to_bf16 = 0; };
struct RoundJobs { RoundJob job[16]; int n; };
int launch_round_copy(const RoundJobs& jobs, cudaStream_t stream);
// TF32 hi / lo split of a matrix, stacked along K for an fp32-grade (3xTF32) tensor-core product.  mode 0: first operand
// [rows, 3 cp] = [hi | lo | hi]; mode 1: K-major second operand [rows, 3 cp] = [hi | hi | lo]; mode 2: MN-major second
// operand [3 rp, cols] = [hi; hi; lo] (cp / rp = cols / rows padded to 4, pads zero)
int launch_split3(const float* src, int ld_src, float* dst, int rows, int cols, int mode, cudaStream_t stream);

// ------------------------------------------------------------------------------------------
// autoregressive regressor (hand_net.py:379-393) and its backward
// ------------------------------------------------------------------------------------------
// pred0 = mean + [0,0,0,feat_out]; `iteration` x pred += [mf|pred] Wr^T + br; root-relative joints.
// states [B, iteration, P] keeps pred before each step (for backward).  P = n_out (66), F = feature width.
int launch_regressor_fwd(const float* main_feat, const float* feat_out, const float* mean_params, const float* Wr,
                         const float* br, float* pred, float* states, float* h_scratch /* [B,P] */, int B, int F, int P,
                         int iteration, int root_relative, cudaStream_t stream, int hoisted = 0);
// h_scratch = main_feat Wr[:, :F]^T + br (iteration invariant); pass hoisted = 1 to launch_regressor_fwd afterwards
int launch_regressor_hoist(const float* main_feat, const float* Wr, const float* br, float* h_scratch, int B, int F, int P,
                           cudaStream_t stream);
// g_pred [B,P] -> d_feat_out [B,P-3], d_main_feat [B,F] (nullable), gsum [B,P], gsteps [B,iteration,P]
int launch_regressor_bwd(const float* g_pred, const float* Wr, float* d_feat_out, float* d_main_feat, float* gsum,
                         float* gsteps, int B, int F, int P, int iteration, int root_relative, cudaStream_t stream,
                         int skip_main_feat_gemm = 0);   // 1: the caller computes d_main_feat = gsum Wr[:, :F] itself

// fused train-step tail: recurrence, root-relative step, closed-form d loss / d pred_params and the reverse recurrence
// in one kernel per sample (h_scratch from launch_regressor_hoist).  Writes pred, states, d_feat_out [B,P-3], gsum,
// gsteps and, if ones_out != null, a ones cotangent [B,P-3] for the stacked path-length sweep.  The loss values are
// launch_proj_loss(..., g_pred = nullptr) on pred, off the critical path.
int launch_regressor_train(const float* feat_out, const float* mean_params, const float* Wr, const float* h_scratch,
                           const float* labels, int ld_labels, float w3d, float w2d, float grad_scale, float* pred,
                           float* states, float* d_feat_out, float* ones_out, float* gsum, float* gsteps, int B, int F, int P,
                           int iteration, cudaStream_t stream);

// d main_feat (nullable), d regressor.weight, d regressor.bias from the backward's gsum / gsteps / states: one launch
int launch_regressor_param_grads(const float* gsum, const float* gsteps, const float* states, const float* main_feat,
                                 const float* Wr, float* d_main_feat, float* dWr, float* dbr, int B, int F, int P, int iteration,
                                 cudaStream_t stream);
// the last feed-forward's second Linear (out = 3, vision_transformer.py:37-42) and its data gradient with the GELU
// derivative fused (act_rows > 0: row m of dY / dZ uses activation row m % act_rows), always fp32
int launch_ff_out3_fwd(const float* H, int ldh, const float* W2, const float* b2, float* Y, int M, int K, cudaStream_t stream);
// dZs (nullable): also the TF32 hi / lo split [MR, 3 x (N padded to 8)] = [hi | lo | hi] of dZ (pad columns must be zero already)
int launch_ff_out3_bwd(const float* dY, const float* W2, const float* Z, int ldz, float* dZ, int lddz, int MR, int N, int act_rows,
                       float* dZs, cudaStream_t stream, int z_is_grad = 0);      // z_is_grad: Z holds gelu'(z) (see GemmArgs)

// ------------------------------------------------------------------------------------------
// projection + losses (train.py:112-120,165-203) with closed-form gradient w.r.t. pred_params
// ------------------------------------------------------------------------------------------
// losses[4] = {loss, l_3d, l_2d, l_pl}; pl_term may be null (then l_pl = 0 and the 10*l_pl term is absent)
int launch_proj_loss(const float* pred, const float* labels, int ld_labels, const float* pl_term, int pl_row_elems,
                     int n_tokens, float w3d, float w2d, float grad_scale, float* losses, float* g_pred,
                     float* pl_scratch, int B, cudaStream_t stream);
// losses[3] = l_pl; losses[0] += 10 * l_pl   (the path-length term alone, train.py:178-183,201)
int launch_pl_loss_add(const float* pl_term, int pl_row_elems, int n_tokens, float* losses, float* pl_scratch, int B,
                       cudaStream_t stream);

// MANO linear blend skinning (mano.py:280-391) lives in lbs.cu with its own C-ABI wrappers.

}  // namespace scat
