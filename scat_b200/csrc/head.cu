// Orchestration of the reg_transformer head on one GPU and the C ABI of include/scat_b200.h.
//
// The forward follows EncoderTransformer.forward after the backbone (hand_net.py:363-398) and
// Transformer.forward (vision_transformer.py:97-101); the backward is the autograd graph of that code
// written out by hand.  All activations needed by the backward live in a
// caller-owned workspace laid out by HeadPlan; nothing here allocates or synchronises.
//
// Precision plan (desc.precision != FP32): the large GEMMs run on tcgen05 kind::tf32.  The tensor core
// truncates fp32 operands, so every tensor that FEEDS a tensor-core GEMM is stored already rounded to
// TF32-nearest by the kernel that produces it (LayerNorm, attention, GELU / dGELU epilogues, LayerNorm
// backward) and the GEMM weights get a rounded (and 16-byte-padded) copy once per forward.  The regressor,
// the last feed-forward (196->147->3) and the conv front end stay fp32.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "../../include/scat_b200.h"
#include "kernels.h"

namespace scat {

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
unsigned long long g_launch_count = 0;
static int read_pdl_env() {
    const char* e = getenv("SCAT_PDL");
    return (e && e[0] == '0') ? 0 : 1;
}
int g_use_pdl = read_pdl_env();
thread_local cudaStream_t tl_side_stream = nullptr;
int g_prio_low = 0, g_prio_high = 0;
static void init_priorities() {
    static std::once_flag once;
    std::call_once(once, [] {
        // Opt-in (SCAT_PRIORITIES=1).  Measured on B200 (profiles/README.md): with the critical chain at the highest and the
        // side stream at the lowest launch priority the step gets SLOWER (0.785 -> 0.817 ms): the side stream starves (its
        // first kernel stretched from ~10 to 100 us), falls behind, and the step ends with ~100 us of side-stream tail
        const char* e = getenv("SCAT_PRIORITIES");
        int least = 0, greatest = 0;
        if (e && e[0] == '1' && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) {
            g_prio_low = least;
            g_prio_high = greatest;
        }
    });
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
    struct Entry { const void* fn; int dev; int bytes; };
    static std::mutex mu;
    static std::vector<Entry> seen;
    int dev = 0;
    SCAT_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : seen)
        if (e.fn == kernel && e.dev == dev && e.bytes >= bytes) return 0;
    SCAT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    seen.push_back(Entry{kernel, dev, bytes});
    return 0;
}

// SCAT_SAVE_DGELU=0 (A/B switch): the forward saves z and the backward evaluates gelu'(z) itself, as in round 1
int g_save_dgelu = [] { const char* e = getenv("SCAT_SAVE_DGELU"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
// experiment switch: bit 0 / 1 / 2 lets the dZ / dNf / dNa GEMM of the backward use the wide single-wave tiles
int g_exp_wide_bwd = [] { const char* e = getenv("SCAT_EXP_WIDE_BWD"); return e ? atoi(e) : 0; }();
// Every kernel asks for the maximum shared-memory carve-out, so that consecutive kernels of the chain never make an SM
// change its L1 / shared split between them: a GEMM CTA (96-192 KB of shared memory) cannot become resident on an SM that
// is still configured for a kernel without shared memory, which defeats the programmatic-launch overlap.  Measured on the
// B=96 step: 0.742 -> 0.722 ms.  SCAT_CARVEOUT=0 restores the driver's per-kernel choice.
int g_carveout = [] { const char* e = getenv("SCAT_CARVEOUT"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
// pct: preferred shared-memory share of the SM's 256 KB (100 = everything a kernel may get).  The first request for a
// kernel wins, so a launcher that wants L1 instead (table-reading FFMA kernels: the LBS skinning) asks before it launches.
void ensure_carveout(const void* kernel, int pct) {
    struct Entry { const void* fn; int dev; };
    static std::mutex mu;
    static std::vector<Entry> seen;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : seen)
        if (e.fn == kernel && e.dev == dev) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct >= 100 ? (int)cudaSharedmemCarveoutMaxShared : pct);
    seen.push_back(Entry{kernel, dev});
}

int launch_gemm(const GemmArgs& g, int precision, cudaStream_t stream) {
    // tiny problems (regressor / N=3 grads) stay on the FFMA kernel: a 128-row tensor tile would be mostly padding
    if (g.operand_bf16) return launch_gemm_tc(g, precision, stream);     // bf16 operands only exist for the tensor core
    if (precision != PREC_FP32 && (long long)g.M * g.N * g.K >= (1LL << 22) && gemm_tc_supported(g))
        return launch_gemm_tc(g, precision, stream);
    return launch_gemm_simt(g, stream);
}

// The contractions that stay fp32-accurate in every precision (last feed-forward, regressor gradients) run on the
// fp32 FFMA kernel.  The 3xTF32 mma.sync kernel (gemm_mma3.cu, scat_gemm precision SCAT_PREC_TF32X3) computes the same
// thing to the same accuracy but was measured SLOWER on B200 (44 us vs 25 us for the 2016x147x196 layer): the legacy
// mma.sync path issues one m16n8k8 TF32 instruction per ~55 cycles per SM sub-partition, i.e. FFMA-class throughput,
// and the split needs three of them.
int launch_gemm_exact(const GemmArgs& g, int head_precision, cudaStream_t stream) {
    (void)head_precision;
    return launch_gemm_simt(g, stream);
}

// ---------------------------------------------------------------------------------------------
// Second stream for work that is off the critical path (weight gradients, bias column sums, gradient zeroing, the
// conv weight gradient): the step is a chain of short, latency-bound kernels that each fill a fraction of the 148
// SMs, so independent launches overlap almost for free.  Fork / join are event record + wait pairs, which is also
// exactly how a side stream joins a CUDA-graph capture of the caller's stream.  One side stream and a small ring
// of events per device, created on first use (the only state this library keeps besides the launch counter);
// SCAT_SIDE_STREAM=0 keeps everything on the caller's stream.  Host threads: a whole-head call holds the device's
// side-stream mutex while it ENQUEUES (SideScope), so concurrent callers on one device interleave call by call; the
// event ring is only ever advanced under that mutex, and a stream wait binds to the event's state at the time of the
// call, so re-recording a ring slot later cannot disturb an earlier wait.
// ---------------------------------------------------------------------------------------------
struct SideStream {
    cudaStream_t s = nullptr;      // weight-gradient GEMMs, weight copies, regressor gradients
    cudaStream_t s2 = nullptr;     // the small reductions (bias column sums, LayerNorm parameter gradients): they fill a few SMs
                                   // each, so they overlap the GEMMs of `s` instead of queueing behind them
    cudaStream_t sx = nullptr;     // what a caller's gradients-ready hook enqueues (the data-parallel exchange)
    cudaEvent_t ev[128];
    int next = 0;
    bool ready = false;
    std::mutex mu;
};
static SideStream g_side[16];
static std::mutex g_side_create;
static int side_enabled() {
    static const int on = [] { const char* e = getenv("SCAT_SIDE_STREAM"); return (e && e[0] == '0') ? 0 : 1; }();
    return on;
}
static SideStream* get_side() {
    if (!side_enabled()) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideStream& sd = g_side[dev];
    std::lock_guard<std::mutex> lock(g_side_create);
    if (!sd.ready) {
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&sd.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&sd.sx, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 128; ++i)
            if (cudaEventCreateWithFlags(&sd.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sd.ready = true;
    }
    return &sd;
}
// the device's side stream for the duration of one whole-head call (enqueue phase), exclusive among host threads
struct SideScope {
    SideStream* sd;
    SideScope() : sd(get_side()) {
        init_priorities();
        if (sd) { sd->mu.lock(); tl_side_stream = sd->s; }
    }
    ~SideScope() {
        tl_side_stream = nullptr;
        if (sd) sd->mu.unlock();
    }
    SideScope(const SideScope&) = delete;
    SideScope& operator=(const SideScope&) = delete;
};
// everything enqueued on `from` so far happens before anything enqueued on `to` from now on
static int order_after(SideStream* sd, cudaStream_t from, cudaStream_t to) {
    if (sd == nullptr || from == to) return 0;
    cudaEvent_t e = sd->ev[sd->next];
    sd->next = (sd->next + 1) % 128;
    SCAT_CHECK_CUDA(cudaEventRecord(e, from));
    SCAT_CHECK_CUDA(cudaStreamWaitEvent(to, e, 0));
    return 0;
}

// join both side streams into `to`
static int join_side(SideStream* sd, cudaStream_t to) {
    if (sd == nullptr) return 0;
    SCAT_PROPAGATE(order_after(sd, sd->s, to));
    return order_after(sd, sd->s2, to);
}

// The caller's gradients-ready hook (scat_head_train_step_hooked): called on the host while the step is being enqueued,
// once per part of the gradient list, with a stream that is ordered after everything that writes that part.  What it
// enqueues there (a gradient exchange) runs beside the rest of the step; the step's end waits for it.
struct GradHook {
    scat_grads_ready_fn fn = nullptr;
    void* user = nullptr;
    bool used = false;
};
// `from`: the stream on which the part's last writer was enqueued (or that already waits for all of them)
static int fire_hook(GradHook* h, SideStream* sd, cudaStream_t from, int part) {
    if (h == nullptr || h->fn == nullptr) return 0;
    const cudaStream_t sx = sd ? sd->sx : from;
    SCAT_PROPAGATE(order_after(sd, from, sx));
    h->used = true;
    const int rc = h->fn(h->user, part, (void*)sx);
    SCAT_REQUIRE(rc == 0, kErrBadArg, "train_step: the gradients-ready hook failed for part %d (rc=%d)", part, rc);
    return 0;
}

// split form: mark a point on `from` now, make `to` wait for it later
static int side_mark(SideStream* sd, cudaStream_t from, cudaEvent_t* out) {
    *out = nullptr;
    if (sd == nullptr) return 0;
    cudaEvent_t e = sd->ev[sd->next];
    sd->next = (sd->next + 1) % 128;
    SCAT_CHECK_CUDA(cudaEventRecord(e, from));
    *out = e;
    return 0;
}
static int side_wait(cudaStream_t to, cudaEvent_t e) {
    if (e == nullptr) return 0;
    SCAT_CHECK_CUDA(cudaStreamWaitEvent(to, e, 0));
    return 0;
}

namespace {

constexpr int kDepth = 3;   // hand_net.py:331 depth=3 (opt.vit_depth is ignored by the reference)
constexpr int kDimHead = 64;

enum ParamIdx {
    P_MASK_TOKEN = 0, P_CONV_W = 1,
    // per layer offsets
    L_NA_W = 0, L_NA_B = 1, L_QKV_W = 2, L_OUT_W = 3, L_OUT_B = 4,
    L_NF_W = 5, L_NF_B = 6, L_FC1_W = 7, L_FC1_B = 8, L_FC2_W = 9, L_FC2_B = 10,
    // last layer has no feed-forward norm
    LL_FC1_W = 5, LL_FC1_B = 6, LL_FC2_W = 7, LL_FC2_B = 8,
    P_REG_W = 33, P_REG_B = 34,
};
inline int layer_base(int l) { return 2 + 11 * l; }
// leading dimension (in floats) of an fp32 tensor-core operand: rows start on 32-byte sectors.  TMA accepts any 16-byte
// pitch, but a pitch that is an odd multiple of 16 bytes (d = 196, hidden 588 / 147, the LBS K = 444) costs the bulk
// loads dearly: tools/gemm_pitch_probe.py, 8192 x 2334 x 444: 68.8 us at pitch 444, 44.7 us at pitch 448
inline int padp(int x) { return (x + 7) / 8 * 8; }
// the same for a bf16 operand (16 elements = 32 bytes); the bf16 tensor lives in the first half of an fp32-sized slot
inline int padh(int x) { return (x + 15) / 16 * 16; }

struct LayerPlan {
    int d, hid, out, ldh;
    size_t X, Na, mean_a, rstd_a, QKV, P, O, X1, Nf, mean_f, rstd_f, Z, H;   // float offsets
    size_t w_qkv, w_out, w_fc1, w_fc2;   // TF32-rounded weight copies (tensor-core layers only)
    size_t w_fc1s, w_fc1d;               // last layer: 3xTF32 split copies of fc1.weight, [hid, 3 d] (forward) and [3 hidp, d] (dgrad)
    int ld_qkv, ld_out, ld_fc1, ld_fc2;  // their leading dimensions (padded to 4 floats)
    bool last;
    int p_na_w, p_na_b, p_qkv, p_out_w, p_out_b, p_nf_w, p_nf_b, p_fc1_w, p_fc1_b, p_fc2_w, p_fc2_b;
};

struct HeadPlan {
    int B, T, C, D, heads, inner, M, it, F, NP;
    LayerPlan L[kDepth];
    // cotangent scratch of the backward, two sets: transformer layer l works in set l & 1, so the side stream may still read
    // one layer's buffers (weight gradients, column sums) while the next layer already writes its own
    struct CotSet { size_t dZ, dNf, dX1, dO, dQKV, dNa, dX, dX16, dX1_16; } cot[2];
    size_t feat_out, states, gsum, gsteps, dfeat, dZ, dNf, dX1, dO, dQKV, dNa, dNa2, dX, dFv, conv_scratch, pl_scratch,
        g_pred, ones, hreg, up2, dX16, dX1_16, w_conv, dFv2, x1s, dZs, total;   // x1s / dZs: 3xTF32 split operands of the last feed-forward   // w_conv: [2T,C] TF32-rounded conv weight, twice; dFv2: hi/lo split of dFv   // dX16 / dX1_16: bf16 shadows of dX / dX1 (PREC_BF16 only)
};

size_t take(size_t& cur, size_t n) {
    const size_t at = cur;
    cur += (n + 63) / 64 * 64;   // 256-byte granules keep every buffer float4 / TMA aligned
    return at;
}

int make_plan(const ScatHeadDesc& d, HeadPlan& p) {
    SCAT_REQUIRE(d.batch > 0 && d.n_tokens > 0 && d.n_tokens <= 128 && d.token_dim > 0 && d.heads > 0, kErrBadArg,
                 "desc: bad batch/n_tokens/token_dim/heads (%d %d %d %d)", d.batch, d.n_tokens, d.token_dim, d.heads);
    SCAT_REQUIRE(d.token_dim % 4 == 0 && d.token_dim <= 1024, kErrUnsupported,
                 "desc: token_dim %d must be a multiple of 4 and <= 1024", d.token_dim);
    SCAT_REQUIRE(d.n_masked >= 0 && d.n_masked <= d.n_tokens, kErrBadArg, "desc: n_masked %d", d.n_masked);
    SCAT_REQUIRE(d.precision >= PREC_FP32 && d.precision <= PREC_BF16, kErrBadArg, "desc: precision %d", d.precision);
    SCAT_REQUIRE(d.x2_dtype == SCAT_DTYPE_F32 || d.x2_dtype == SCAT_DTYPE_BF16, kErrBadArg, "desc: x2_dtype %d", d.x2_dtype);
    SCAT_REQUIRE(d.x2_dtype == SCAT_DTYPE_F32 || d.precision != PREC_FP32, kErrUnsupported,
                 "desc: the bf16 seam (x2_dtype) needs a tensor-core precision (tf32 / bf16), fp32 is the CUDA-core parity mode");
    p.B = d.batch; p.T = d.n_tokens; p.C = d.channels; p.D = d.token_dim; p.heads = d.heads;
    p.inner = kDimHead * d.heads; p.M = d.batch * d.n_tokens; p.it = d.iteration; p.F = d.main_feat_dim; p.NP = d.n_out;
    size_t cur = 0;
    const size_t M = (size_t)p.M;
    int dim = d.token_dim, dmax = d.token_dim, ldh_max = 0;
    for (int l = 0; l < kDepth; ++l) {
        LayerPlan& L = p.L[l];
        L.last = (l == kDepth - 1);
        L.d = dim; L.hid = (dim * 3) / 4; L.out = L.last ? 3 : dim / 2; L.ldh = padp(L.hid);
        ldh_max = L.ldh > ldh_max ? L.ldh : ldh_max;
        L.X = take(cur, M * dim);
        L.Na = take(cur, M * padp(dim));     // GEMM operand: rows padded to 16 bytes (TMA), e.g. dim 98 of the token variant
        L.mean_a = take(cur, M); L.rstd_a = take(cur, M);
        L.QKV = take(cur, M * 3 * p.inner);
        L.P = take(cur, (size_t)p.B * p.heads * p.T * p.T);
        L.O = take(cur, M * p.inner);
        L.X1 = take(cur, M * dim);
        if (!L.last) { L.Nf = take(cur, M * padp(dim)); L.mean_f = take(cur, M); L.rstd_f = take(cur, M); }
        else { L.Nf = L.X1; L.mean_f = L.rstd_f = 0; }
        L.Z = take(cur, M * L.ldh);
        L.H = take(cur, M * L.ldh);
        L.ld_qkv = padp(dim); L.ld_out = p.inner; L.ld_fc1 = padp(dim); L.ld_fc2 = padp(L.hid);
        L.w_qkv = take(cur, (size_t)3 * p.inner * L.ld_qkv);
        L.w_out = take(cur, (size_t)dim * L.ld_out);
        if (!L.last) {
            L.w_fc1 = take(cur, (size_t)L.hid * L.ld_fc1);
            L.w_fc2 = take(cur, (size_t)L.out * L.ld_fc2);
            L.w_fc1s = L.w_fc1d = 0;
        } else {
            L.w_fc1 = L.w_fc2 = 0;
            L.w_fc1s = take(cur, (size_t)L.hid * 3 * padp(dim));
            L.w_fc1d = take(cur, (size_t)3 * padp(L.hid) * dim);
        }
        const int base = layer_base(l);
        L.p_na_w = base + L_NA_W; L.p_na_b = base + L_NA_B; L.p_qkv = base + L_QKV_W; L.p_out_w = base + L_OUT_W;
        L.p_out_b = base + L_OUT_B;
        if (!L.last) {
            L.p_nf_w = base + L_NF_W; L.p_nf_b = base + L_NF_B; L.p_fc1_w = base + L_FC1_W; L.p_fc1_b = base + L_FC1_B;
            L.p_fc2_w = base + L_FC2_W; L.p_fc2_b = base + L_FC2_B;
        } else {
            L.p_nf_w = L.p_nf_b = -1; L.p_fc1_w = base + LL_FC1_W; L.p_fc1_b = base + LL_FC1_B;
            L.p_fc2_w = base + LL_FC2_W; L.p_fc2_b = base + LL_FC2_B;
        }
        if (!L.last) dim /= 2;
    }
    p.feat_out = take(cur, M * 3);
    const int it = p.it > 0 ? p.it : 1, NP = p.NP > 0 ? p.NP : 1;
    p.states = take(cur, (size_t)p.B * it * NP);
    p.gsum = take(cur, (size_t)p.B * NP);
    p.gsteps = take(cur, (size_t)p.B * it * NP);
    p.dfeat = take(cur, M * 3);
    p.ones = take(cur, M * 3);
    p.g_pred = take(cur, (size_t)p.B * NP);
    p.hreg = take(cur, (size_t)p.B * NP);
    // gradient scratch: the fused train step sweeps the real cotangent and the path-length (ones) cotangent
    // through the backward together, stacked along the row dimension (2M rows)
    const size_t MS = d.pl_reg ? 2 * M : M;
    p.up2 = take(cur, MS * 3);
    p.dZ = take(cur, MS * ldh_max);
    p.dNf = take(cur, MS * dmax);
    p.dX1 = take(cur, MS * dmax);
    p.dO = take(cur, MS * p.inner);
    p.dQKV = take(cur, MS * 3 * p.inner);
    p.dNa = take(cur, MS * dmax);
    p.dNa2 = take(cur, MS * dmax);     // layers alternate: the side stream may still read a layer's dNa (LayerNorm parameter
                                       // gradients) while the next layer writes its own
    p.dX = take(cur, MS * dmax);
    p.dX16 = p.dX1_16 = 0;
    p.w_conv = take(cur, p.C > 0 ? conv_weight_prep_floats(p.C, p.T) : 64);
    p.dFv2 = take(cur, (p.C > 0 && d.precision != PREC_FP32) ? conv_split_floats(p.B, dmax, p.T) : 64);
    if (d.precision == PREC_BF16) {
        p.dX16 = take(cur, (MS * padh(dmax) + 1) / 2);
        p.dX1_16 = take(cur, (MS * padh(dmax) + 1) / 2);
    }
    p.cot[0] = HeadPlan::CotSet{p.dZ, p.dNf, p.dX1, p.dO, p.dQKV, p.dNa, p.dX, p.dX16, p.dX1_16};
    {
        HeadPlan::CotSet& c = p.cot[1];
        c.dZ = take(cur, MS * ldh_max); c.dNf = take(cur, MS * dmax); c.dX1 = take(cur, MS * dmax);
        c.dO = take(cur, MS * p.inner); c.dQKV = take(cur, MS * 3 * p.inner); c.dNa = p.dNa2; c.dX = take(cur, MS * dmax);
        c.dX16 = c.dX1_16 = 0;
        if (d.precision == PREC_BF16) {
            c.dX16 = take(cur, (MS * padh(dmax) + 1) / 2);
            c.dX1_16 = take(cur, (MS * padh(dmax) + 1) / 2);
        }
    }
    {
        const LayerPlan& LL = p.L[kDepth - 1];
        p.x1s = take(cur, M * 3 * padp(LL.d));
        p.dZs = take(cur, MS * 3 * padp(LL.hid));
    }
    p.dFv = take(cur, M * dmax);
    p.conv_scratch = take(cur, p.C > 0 ? conv_wgrad_scratch_floats(p.C, p.T) : 64);
    p.pl_scratch = take(cur, (size_t)p.B);
    p.total = cur;
    return 0;
}

__global__ void fill_kernel(float* p, float v, long long n) {
    pdl_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
    pdl_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
// one launch clears every parameter-gradient tensor: split-K GEMMs, column sums and LayerNorm affine gradients
// all combine partial results with reductions in L2, and a memset node per tensor would serialise the graph
struct ZeroJobs { float* ptr[40]; unsigned n[40]; int count; };
__global__ void zero_many_kernel(const ZeroJobs jobs) {
    pdl_sync();
    float* p = jobs.ptr[blockIdx.y];
    const unsigned n = jobs.n[blockIdx.y];
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        for (unsigned i = tid; i < (n >> 2); i += nthr) reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (unsigned i = (n & ~3u) + tid; i < n; i += nthr) p[i] = 0.f;
    } else {
        for (unsigned i = tid; i < n; i += nthr) p[i] = 0.f;
    }
}

__global__ void token_mean_kernel(const float* __restrict__ X, float* __restrict__ out, int n) {
    pdl_sync();
    // out[b, c] = mean_t X[b, t, c], c < 3  (hand_net.py:203)
    const int b = blockIdx.x, c = threadIdx.x;
    if (c >= 3) return;
    float s = 0.f;
    for (int t = 0; t < n; ++t) s += X[((long long)b * n + t) * 3 + c];
    out[b * 3 + c] = s / (float)n;
}

// Tail of EncoderTransformerCoarse.forward (hand_net.py:289-302): joints = mean template + transformer output, made
// relative to joint 1; camera = Linear(1027 -> 3)(cat(main_feat, mean_params[:3])), applied once.  One warp per sample.
__global__ void coarse_tail_kernel(const float* __restrict__ main_feat, const float* __restrict__ feat_out,
                                   const float* __restrict__ mean_params, const float* __restrict__ Wr /* [3, F + 3] */,
                                   const float* __restrict__ br, float* __restrict__ pred, int B, int F) {
    pdl_sync();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* mf = main_feat + (long long)b * F;
    float cam[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < F; k += 32) {
        const float v = mf[k];
#pragma unroll
        for (int j = 0; j < 3; ++j) cam[j] = fmaf(v, Wr[j * (F + 3) + k], cam[j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        cam[j] = warp_sum(cam[j]);
#pragma unroll
        for (int q = 0; q < 3; ++q) cam[j] = fmaf(mean_params[q], Wr[j * (F + 3) + F + q], cam[j]);   // pred_params[:, :3] = mean here
        cam[j] += br[j];
    }
    float* out = pred + (long long)b * 66;
    if (lane < 3) out[lane] = cam[lane];
    const float* fo = feat_out + (long long)b * 63;
    for (int e = lane; e < 63; e += 32) {
        const int c = e % 3;
        const float root = mean_params[3 + 3 + c] + fo[3 + c];                 // joint 1 before the subtraction
        out[3 + e] = (e / 3 == 1) ? 0.f : (mean_params[3 + e] + fo[e]) - root;
    }
}

// Weight views used by the GEMMs of one layer: the caller's fp32 tensors, or the rounded/padded copies.
struct LayerW {
    const float *qkv, *out, *fc1, *fc2;
    int ld_qkv, ld_out, ld_fc1, ld_fc2;
};
LayerW layer_weights(const HeadPlan& p, int l, const float* const* W, const float* ws, int prec) {
    const LayerPlan& L = p.L[l];
    LayerW w;
    const bool tc = prec != PREC_FP32, bf = prec == PREC_BF16;
    // tensor-core modes read the per-forward copies in the workspace: TF32-rounded fp32 (leading dimension padded to
    // 4) or bf16 (padded to 8, stored in the first half of the same slot)
    w.qkv = tc ? ws + L.w_qkv : W[L.p_qkv];   w.ld_qkv = bf ? padh(L.d) : tc ? L.ld_qkv : L.d;
    w.out = tc ? ws + L.w_out : W[L.p_out_w]; w.ld_out = p.inner;
    const bool tc_ff = tc && !L.last;         // last feed-forward stays fp32 on the caller's weights
    w.fc1 = tc_ff ? ws + L.w_fc1 : W[L.p_fc1_w]; w.ld_fc1 = !tc_ff ? L.d : bf ? padh(L.d) : L.ld_fc1;
    w.fc2 = tc_ff ? ws + L.w_fc2 : W[L.p_fc2_w]; w.ld_fc2 = !tc_ff ? L.hid : bf ? padh(L.hid) : L.ld_fc2;
    return w;
}

// one launch: TF32-round (and pad the leading dimension of) every weight a tensor-core GEMM reads
// (the conv weight stacks are prepared by launch_conv_weight_prep, conv_tc.cu)
int round_weights(const HeadPlan& p, const float* const* W, float* ws, int prec, cudaStream_t st) {
    RoundJobs jobs;
    int n = 0;
    const bool bf = prec == PREC_BF16;
    for (int l = 0; l < kDepth; ++l) {
        const LayerPlan& L = p.L[l];
        jobs.job[n++] = RoundJob{W[L.p_qkv], ws + L.w_qkv, 3 * p.inner, L.d, L.d, bf ? padh(L.d) : L.ld_qkv};
        jobs.job[n++] = RoundJob{W[L.p_out_w], ws + L.w_out, L.d, p.inner, p.inner, p.inner};
        if (!L.last) {
            jobs.job[n++] = RoundJob{W[L.p_fc1_w], ws + L.w_fc1, L.hid, L.d, L.d, bf ? padh(L.d) : L.ld_fc1};
            jobs.job[n++] = RoundJob{W[L.p_fc2_w], ws + L.w_fc2, L.out, L.hid, L.hid, bf ? padh(L.hid) : L.ld_fc2};
        }
    }
    for (int i = 0; i < n; ++i) jobs.job[i].to_bf16 = bf ? 1 : 0;
    jobs.n = n;
    if (n > 0) SCAT_PROPAGATE(launch_round_copy(jobs, st));
    // the last feed-forward's first Linear runs as an fp32-grade 3xTF32 product on the tensor core: split copies of its
    // weight, K-major [hi | hi | lo] for the forward and MN-major [hi; hi; lo] for the data gradient
    const LayerPlan& LL = p.L[kDepth - 1];
    SCAT_PROPAGATE(launch_split3(W[LL.p_fc1_w], LL.d, ws + LL.w_fc1s, LL.hid, LL.d, 1, st));
    SCAT_PROPAGATE(launch_split3(W[LL.p_fc1_w], LL.d, ws + LL.w_fc1d, LL.hid, LL.d, 2, st));
    return 0;
}

// ---- forward through the transformer (vision_transformer.py:97-101) --------------------------------
// Storage of the tensors that feed tensor-core GEMMs (Na, O, Nf, H; backward: dY, dZ, dX1, dQKV):
//   PREC_TF32: fp32, rounded to TF32-nearest by the producing kernel
//   PREC_BF16: bf16 in the same workspace slot (leading dimension padded to 8 elements); tensors that a
//              non-GEMM kernel also reads in fp32 (dX, dX1) get a separate bf16 shadow
// attn_variant: the wiring of models/vision_transformer_attn.py:104-113 (reg_transformer_coarse): attention on the RAW
// token matrix, LayerNorm on the attention branch's OUTPUT, then the residual:  x = LN(attn(x)) + x ; x = ff(x).
// The slot Na then holds the attention branch's output (the LayerNorm's input).  Forward only.
int transformer_forward(const HeadPlan& p, const float* const* W, float* ws, int prec, cudaStream_t st,
                        float* X0_override, bool attn_variant = false) {
    const int M = p.M;
    const bool tc = prec != PREC_FP32, bf = prec == PREC_BF16;
    const int omode = bf ? OUT_BF16 : tc ? OUT_TF32 : OUT_F32;
    for (int l = 0; l < kDepth; ++l) {
        const LayerPlan& L = p.L[l];
        const LayerW w = layer_weights(p, l, W, ws, prec);
        float* X = (l == 0 && X0_override) ? X0_override : ws + L.X;
        const int ld_n = bf ? padh(L.d) : padp(L.d);       // leading dimension of Na / Nf (16-byte rows for TMA)
        const int ld_h = bf ? padh(L.hid) : L.ldh;         // leading dimension of H as a GEMM operand
        GemmArgs g;
        if (!attn_variant) {
            // PreNorm + Attention + Residual (:18,:26,:59-79)
            SCAT_PROPAGATE(launch_layernorm_fwd(X, L.d, W[L.p_na_w], W[L.p_na_b], ws + L.Na, ld_n, ws + L.mean_a,
                                                ws + L.rstd_a, M, L.d, omode, st));
            g.A = ws + L.Na; g.sam = ld_n; g.sak = 1; g.B = w.qkv; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_qkv; g.sbk = 1; g.operand_bf16 = bf;
            g.C = ws + L.QKV; g.ldc = 3 * p.inner; g.M = M; g.N = 3 * p.inner; g.K = L.d; g.prerounded = tc;
            SCAT_PROPAGATE(launch_gemm(g, prec, st));
            SCAT_PROPAGATE(launch_attention_fwd(ws + L.QKV, ws + L.O, ws + L.P, p.B, p.T, p.heads, omode, st));
            g = GemmArgs();
            g.A = ws + L.O; g.sam = p.inner; g.sak = 1; g.B = w.out; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_out; g.sbk = 1; g.operand_bf16 = bf;
            g.C = ws + L.X1; g.ldc = L.d; g.M = M; g.N = L.d; g.K = p.inner; g.prerounded = tc;
            g.epilogue = EPI_BIAS_RESID; g.bias = W[L.p_out_b]; g.aux_in = X; g.ld_aux_in = L.d;
            SCAT_PROPAGATE(launch_gemm(g, prec, st));
        } else {
            // x1, attn = attention(x); x = pren(x1) + x (vision_transformer_attn.py:106-108): X is the caller's /
            // previous layer's fp32 tensor, so the tensor-core GEMM rounds it to TF32-nearest in shared memory itself
            g.A = X; g.sam = L.d; g.sak = 1; g.B = w.qkv; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_qkv; g.sbk = 1;
            g.C = ws + L.QKV; g.ldc = 3 * p.inner; g.M = M; g.N = 3 * p.inner; g.K = L.d; g.prerounded = 0;
            SCAT_PROPAGATE(launch_gemm(g, prec, st));
            SCAT_PROPAGATE(launch_attention_fwd(ws + L.QKV, ws + L.O, ws + L.P, p.B, p.T, p.heads, omode, st));
            g = GemmArgs();
            g.A = ws + L.O; g.sam = p.inner; g.sak = 1; g.B = w.out; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_out; g.sbk = 1;
            g.C = ws + L.Na; g.ldc = L.d; g.M = M; g.N = L.d; g.K = p.inner; g.prerounded = tc;
            g.epilogue = EPI_BIAS; g.bias = W[L.p_out_b];
            SCAT_PROPAGATE(launch_gemm(g, prec, st));
            SCAT_PROPAGATE(launch_layernorm_fwd(ws + L.Na, L.d, W[L.p_na_w], W[L.p_na_b], ws + L.X1, L.d, ws + L.mean_a,
                                                ws + L.rstd_a, M, L.d, OUT_F32, st, X, L.d));
        }
        // (PreNorm +) FeedForward, no residual (:44,:94 / :89)
        if (!L.last)
            SCAT_PROPAGATE(launch_layernorm_fwd(ws + L.X1, L.d, W[L.p_nf_w], W[L.p_nf_b], ws + L.Nf, ld_n,
                                                ws + L.mean_f, ws + L.rstd_f, M, L.d, omode, st));
        const int ffprec = L.last ? PREC_FP32 : prec;   // last FF stays fp32
        const bool fftc = ffprec != PREC_FP32, ffbf = ffprec == PREC_BF16;
        g = GemmArgs();
        g.A = ws + L.Nf; g.sam = L.last ? L.d : ld_n; g.sak = 1; g.B = w.fc1; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_fc1; g.sbk = 1; g.operand_bf16 = ffbf;
        g.M = M; g.N = L.hid; g.K = L.d; g.prerounded = fftc;
        if (ffbf) { g.C16 = ws + L.H; g.ldc16 = ld_h; }                         // H exists only as bf16
        else { g.C = ws + L.H; g.ldc = L.ldh; g.round_out = fftc; }
        // (the slot Z holds gelu'(z), not z: the only reader is the backward's dGELU epilogue, which then is a multiply)
        g.epilogue = EPI_BIAS_GELU; g.bias = W[L.p_fc1_b]; g.aux_out = ws + L.Z; g.ld_aux_out = L.ldh; g.gelu_saves_grad = g_save_dgelu;
        if (L.last && tc) {
            // fp32-grade on the tensor core: X1 split into [hi | lo | hi] against the weight's [hi | hi | lo], K = 3 d
            const int dp = padp(L.d);
            SCAT_PROPAGATE(launch_split3(ws + L.X1, L.d, ws + p.x1s, M, L.d, 0, st));
            g.A = ws + p.x1s; g.sam = 3 * dp; g.B = ws + L.w_fc1s; g.sbn = 3 * dp; g.K = 3 * dp; g.prerounded = 1;
            SCAT_PROPAGATE(launch_gemm_tc(g, PREC_TF32, st));
        } else {
            SCAT_PROPAGATE(L.last ? launch_gemm_exact(g, prec, st) : launch_gemm(g, ffprec, st));
        }
        float* Y = L.last ? ws + p.feat_out : ws + p.L[l + 1].X;
        g = GemmArgs();
        g.A = ws + L.H; g.sam = ffbf ? ld_h : L.ldh; g.sak = 1; g.B = w.fc2; g.b_static = 1; g.allow_wide = 1; g.sbn = w.ld_fc2; g.sbk = 1; g.operand_bf16 = ffbf;
        g.C = Y; g.ldc = L.out; g.M = M; g.N = L.out; g.K = L.hid; g.prerounded = fftc;
        g.epilogue = EPI_BIAS; g.bias = W[L.p_fc2_b];
        if (L.last && L.out == 3)     // three outputs per token: a warp per row instead of a 64 x 64 tile kernel
            SCAT_PROPAGATE(launch_ff_out3_fwd(ws + L.H, L.ldh, W[L.p_fc2_w], W[L.p_fc2_b], Y, M, L.hid, st));
        else
            SCAT_PROPAGATE(L.last ? launch_gemm_exact(g, prec, st) : launch_gemm(g, ffprec, st));
    }
    return 0;
}

// ---- reverse sweep: d/dX0 of <up, feat_out>, optionally with parameter gradients ---------------------
// up: [M,3] cotangent of the transformer output.  Result lands in ws + p.dX ([M, D]).  Uses the weight copies
// left in the workspace by transformer_forward.
// sweeps = 2: `up` holds two stacked cotangents [2M,3] (rows < M: the real one, rows >= M: the path-length ones);
// every dgrad-type kernel then runs once over 2M rows against the same saved activations, parameter gradients
// only see the first M rows.
int transformer_backward(const HeadPlan& p, const float* const* W, float* const* G /* null = dgrad only */, float* ws,
                         int prec, const float* up, cudaStream_t st, const float* X0_override, int sweeps = 1,
                         SideStream* sd = nullptr, int l_first = kDepth - 1, int l_last = 0, GradHook* hook = nullptr) {
    // parameter gradients run on the side stream `sg` (== st without one): each group is forked once its inputs exist.
    // Layer l works in cotangent set l & 1, so the side stream may lag a whole layer behind the critical chain: the main
    // stream only waits (at the top of layer l) for the side work of layer l + 2, whose buffers it is about to overwrite
    const cudaStream_t sg = (sd != nullptr && G != nullptr) ? sd->s : st;
    const cudaStream_t sb = (sd != nullptr && G != nullptr) ? sd->s2 : st;     // column sums, LayerNorm parameter gradients
    const int M = p.M;
    const int MR = M * sweeps;                       // rows of every cotangent tensor
    const int amod = sweeps > 1 ? M : 0;             // activation row = cotangent row % M
    const bool tc = prec != PREC_FP32, bf = prec == PREC_BF16;
    const int omode = bf ? OUT_BF16 : tc ? OUT_TF32 : OUT_F32;
    const float* dY = up;                            // fp32 cotangent of the layer output
    const float* dYg = up;                           // the copy GEMMs read: same tensor, or its bf16 shadow
    int ld_dYg = 3;
    if (l_first < kDepth - 1) {                      // resuming below the top layer: the cotangent is layer l_first+1's dX
        const int dn = p.L[l_first + 1].d;
        const HeadPlan::CotSet& prev = p.cot[(l_first + 1) & 1];
        dY = ws + prev.dX;
        dYg = bf ? ws + prev.dX16 : dY;
        ld_dYg = bf ? padh(dn) : dn;
    }
    cudaEvent_t layer_done[kDepth] = {nullptr, nullptr, nullptr}, dy_read[kDepth] = {nullptr, nullptr, nullptr};
    for (int l = l_first; l >= l_last; --l) {
        const LayerPlan& L = p.L[l];
        const HeadPlan::CotSet& c = p.cot[l & 1];
        if (G && l + 2 <= l_first) SCAT_PROPAGATE(side_wait(st, layer_done[l + 2]));
        const LayerW w = layer_weights(p, l, W, ws, prec);
        const float* X = (l == 0 && X0_override) ? X0_override : ws + L.X;
        const int ffprec = L.last ? PREC_FP32 : prec;
        const bool fftc = ffprec != PREC_FP32, ffbf = ffprec == PREC_BF16;
        const int ld_n = bf ? padh(L.d) : L.d;                 // dX1 / dX as GEMM operands
        const int ld_na = bf ? padh(L.d) : padp(L.d);          // saved Na / Nf (as the forward stored them)
        const int ld_h = ffbf ? padh(L.hid) : L.ldh;           // H / dZ as GEMM operands
        float* dZ = ws + c.dZ;
        GemmArgs g;
        if (G) {
            // dW2[out,hid] = dY^T H ; db2 = colsum(dY)
            SCAT_PROPAGATE(order_after(sd, st, sg));
            g.A = dYg; g.sam = 1; g.sak = ld_dYg; g.B = ws + L.H; g.b_static = 1; g.sbn = 1; g.sbk = ld_h; g.operand_bf16 = ffbf;
            g.C = G[L.p_fc2_w]; g.ldc = L.hid; g.M = L.out; g.N = L.hid; g.K = M; g.allow_split_k = 1; g.c_zeroed = 1; g.prerounded = fftc;
            SCAT_PROPAGATE((L.last ? launch_gemm_exact(g, prec, sg) : launch_gemm(g, ffprec, sg)));
            // db2: layer l + 1's attention-LayerNorm parameter kernel summed its dX (= this dY) already, unless this call
            // starts here
            if (l == l_first) {
                SCAT_PROPAGATE(order_after(sd, st, sb));
                SCAT_PROPAGATE(launch_colsum(dY, L.out, M, L.out, G[L.p_fc2_b], 1, sb));
            }
            SCAT_PROPAGATE(order_after(sd, sb, sg));
            SCAT_PROPAGATE(side_mark(sd, sg, &dy_read[l]));        // dY (the other set's dX) has been consumed on the side streams
        }
        // dZ = (dY W2) * gelu'(Z)
        g = GemmArgs();
        g.A = dYg; g.sam = ld_dYg; g.sak = 1; g.B = w.fc2; g.b_static = 1; g.sbn = 1; g.sbk = w.ld_fc2; g.operand_bf16 = ffbf;
        g.M = MR; g.N = L.hid; g.K = L.out; g.prerounded = fftc;
        if (ffbf) { g.C16 = dZ; g.ldc16 = ld_h; }
        else { g.C = dZ; g.ldc = L.ldh; g.round_out = fftc; }
        g.epilogue = EPI_DGELU; g.aux_in = ws + L.Z; g.ld_aux_in = L.ldh; g.aux_row_mod = amod; g.gelu_saves_grad = g_save_dgelu;
        g.allow_wide = g_exp_wide_bwd & 1;
        if (L.last && L.out == 3 && !ffbf)     // K = 3: elementwise
            SCAT_PROPAGATE(launch_ff_out3_bwd(dY, W[L.p_fc2_w], ws + L.Z, L.ldh, dZ, L.ldh, MR, L.hid, amod, tc ? ws + p.dZs : nullptr, st, /*z_is_grad=*/g_save_dgelu));
        else
            SCAT_PROPAGATE((L.last ? launch_gemm_exact(g, prec, st) : launch_gemm(g, ffprec, st)));
        if (G) {
            // dW1[hid,d] = dZ^T Nf ; db1 = colsum(dZ)
            SCAT_PROPAGATE(order_after(sd, st, sg));
            g = GemmArgs();
            g.A = dZ; g.sam = 1; g.sak = ld_h; g.B = ws + L.Nf; g.b_static = 1; g.sbn = 1; g.sbk = L.last ? L.d : ld_na; g.operand_bf16 = ffbf;
            g.C = G[L.p_fc1_w]; g.ldc = L.d; g.M = L.hid; g.N = L.d; g.K = M; g.allow_split_k = 1; g.c_zeroed = 1; g.prerounded = fftc;
            SCAT_PROPAGATE((L.last ? launch_gemm_exact(g, prec, sg) : launch_gemm(g, ffprec, sg)));
            SCAT_PROPAGATE(order_after(sd, st, sb));
            SCAT_PROPAGATE(launch_colsum(dZ, ld_h, M, L.hid, G[L.p_fc1_b], 1, sb, ffbf));
        }
        // dNf = dZ W1   (for the last layer this IS dX1 and feeds the tensor-core out-projection GEMMs)
        g = GemmArgs();
        g.A = dZ; g.sam = ld_h; g.sak = 1; g.B = w.fc1; g.b_static = 1; g.sbn = 1; g.sbk = w.ld_fc1; g.operand_bf16 = ffbf;
        g.C = ws + c.dNf; g.ldc = L.d; g.M = MR; g.N = L.d; g.K = L.hid; g.prerounded = fftc;
        if (L.last && bf) { g.C16 = ws + c.dX1_16; g.ldc16 = ld_n; }
        else g.round_out = (L.last && tc) ? 1 : 0;
        g.allow_wide = (g_exp_wide_bwd >> 1) & 1;
        if (L.last && tc && L.out == 3) {
            // fp32-grade on the tensor core: dZ's [hi | lo | hi] (written by the kernel above) against fc1.weight's [hi; hi; lo]
            const int hp = padp(L.hid);
            g.A = ws + p.dZs; g.sam = 3 * hp; g.B = ws + L.w_fc1d; g.sbk = L.d; g.K = 3 * hp; g.prerounded = 1;
            SCAT_PROPAGATE(launch_gemm_tc(g, PREC_TF32, st));
        } else {
            SCAT_PROPAGATE((L.last ? launch_gemm_exact(g, prec, st) : launch_gemm(g, ffprec, st)));
        }
        const float* dX1 = ws + c.dNf;
        if (!L.last) {
            // data gradient on the critical path; d gamma / d beta are column sums over the M real rows: side stream
            SCAT_PROPAGATE(launch_layernorm_bwd(ws + c.dNf, L.d, ws + L.X1, L.d, W[L.p_nf_w], ws + L.mean_f,
                                                ws + L.rstd_f, nullptr, 0, ws + c.dX1, L.d, nullptr, nullptr, MR, L.d,
                                                bf ? OUT_F32 : omode, st, amod, bf ? ws + c.dX1_16 : nullptr, ld_n));
            dX1 = ws + c.dX1;
        }
        const float* dX1g = bf ? ws + c.dX1_16 : dX1;           // what the GEMMs read
        if (G) {
            // dWo[d,inner] = dX1^T O ; dbo = colsum(dX1); and the feed-forward LayerNorm's d gamma / d beta (reads dNf)
            SCAT_PROPAGATE(order_after(sd, st, sg));
            SCAT_PROPAGATE(order_after(sd, st, sb));
            if (!L.last)     // ... with dbo = colsum(dX1) riding along
                SCAT_PROPAGATE(launch_layernorm_param_grads(ws + c.dNf, L.d, ws + L.X1, L.d, ws + L.mean_f, ws + L.rstd_f,
                                                            G[L.p_nf_w], G[L.p_nf_b], M, L.d, sb, dX1, L.d, G[L.p_out_b]));
            if (l == 0 && hook != nullptr && hook->fn != nullptr) {
                // layer 0's feed-forward half (parameters 7..12: its LayerNorm, fc1, fc2) is final once the kernels queued
                // so far on the two side streams have run: part 1 of the hook, ~60 us before the attention half
                SCAT_PROPAGATE(order_after(sd, sb, sg));
                SCAT_PROPAGATE(fire_hook(hook, sd, sg, 1));
            }
            g = GemmArgs();
            g.A = dX1g; g.sam = 1; g.sak = ld_n; g.B = ws + L.O; g.b_static = 1; g.sbn = 1; g.sbk = p.inner; g.operand_bf16 = bf;
            g.C = G[L.p_out_w]; g.ldc = p.inner; g.M = L.d; g.N = p.inner; g.K = M; g.allow_split_k = 1; g.c_zeroed = 1; g.prerounded = tc;
            SCAT_PROPAGATE(launch_gemm(g, prec, sg));
            if (L.last) SCAT_PROPAGATE(launch_colsum(dX1, L.d, M, L.d, G[L.p_out_b], 1, sb));
        }
        // dO = dX1 Wo
        g = GemmArgs();
        g.A = dX1g; g.sam = ld_n; g.sak = 1; g.B = w.out; g.b_static = 1; g.sbn = 1; g.sbk = w.ld_out; g.operand_bf16 = bf;
        g.C = ws + c.dO; g.ldc = p.inner; g.M = MR; g.N = p.inner; g.K = L.d; g.prerounded = tc;
        SCAT_PROPAGATE(launch_gemm(g, prec, st));
        SCAT_PROPAGATE(launch_attention_bwd(ws + L.QKV, ws + L.P, ws + c.dO, ws + c.dQKV, p.B * sweeps, p.T, p.heads, omode,
                                            st, sweeps > 1 ? p.B : 0));
        if (G) {
            // dWqkv[3inner,d] = dQKV^T Na
            SCAT_PROPAGATE(order_after(sd, st, sg));
            g = GemmArgs();
            g.A = ws + c.dQKV; g.sam = 1; g.sak = 3 * p.inner; g.B = ws + L.Na; g.b_static = 1; g.sbn = 1; g.sbk = ld_na; g.operand_bf16 = bf;
            g.C = G[L.p_qkv]; g.ldc = L.d; g.M = 3 * p.inner; g.N = L.d; g.K = M; g.allow_split_k = 1; g.c_zeroed = 1; g.prerounded = tc;
            SCAT_PROPAGATE(launch_gemm(g, prec, sg));
        }
        // dNa = dQKV Wqkv
        g = GemmArgs();
        float* dNa = ws + c.dNa;
        g.A = ws + c.dQKV; g.sam = 3 * p.inner; g.sak = 1; g.B = w.qkv; g.b_static = 1; g.sbn = 1; g.sbk = w.ld_qkv; g.operand_bf16 = bf;
        g.C = dNa; g.ldc = L.d; g.M = MR; g.N = L.d; g.K = 3 * p.inner; g.prerounded = tc;
        g.allow_wide = (g_exp_wide_bwd >> 2) & 1;
        SCAT_PROPAGATE(launch_gemm(g, prec, st));
        // dX = LN_a'(dNa) + dX1 (residual, vision_transformer.py:18); it is the dY of layer l-1's tensor-core GEMMs.  It
        // overwrites the dX of layer l + 2 = the dY that layer l + 1's first side-stream group read
        if (G && l + 1 <= l_first) SCAT_PROPAGATE(side_wait(st, dy_read[l + 1]));
        const bool feeds_gemm = tc && l > 0;
        SCAT_PROPAGATE(launch_layernorm_bwd(dNa, L.d, X, L.d, W[L.p_na_w], ws + L.mean_a, ws + L.rstd_a, dX1,
                                            L.d, ws + c.dX, L.d, nullptr, nullptr, MR, L.d,
                                            (feeds_gemm && !bf) ? OUT_TF32 : OUT_F32, st, amod,
                                            (feeds_gemm && bf) ? ws + c.dX16 : nullptr, ld_n));
        if (G) {
            // its parameter gradients read dNa and the saved X: side stream, joined two layers later or by the caller
            SCAT_PROPAGATE(order_after(sd, st, sb));
            // (and, when layer l - 1 follows in this call, its db2 = colsum of the dX just written)
            const bool fuse_db2 = l > l_last;
            SCAT_PROPAGATE(launch_layernorm_param_grads(dNa, L.d, X, L.d, ws + L.mean_a, ws + L.rstd_a, G[L.p_na_w], G[L.p_na_b],
                                                        M, L.d, sb, fuse_db2 ? ws + c.dX : nullptr, L.d,
                                                        fuse_db2 ? G[p.L[l - 1].p_fc2_b] : nullptr));
            SCAT_PROPAGATE(order_after(sd, sb, sg));
            SCAT_PROPAGATE(side_mark(sd, sg, &layer_done[l]));     // everything of layer l on the side streams is queued
            // `sg` now follows every writer of the gradients of layers >= l (and of the regressor, queued on it before):
            // layers 1.. are part 0 of the hook, the attention half of layer 0 (parameters 2..6) is part 2
            if (l == 1 || l == 0) SCAT_PROPAGATE(fire_hook(hook, sd, sg, l == 1 ? 0 : 2));
        }
        dY = ws + c.dX;
        dYg = bf ? ws + c.dX16 : dY;
        ld_dYg = bf ? ld_n : L.d;
    }
    return 0;
}

int zero_param_grads(const HeadPlan& p, float* const* G, cudaStream_t st) {
    ZeroJobs z;
    int n = 0;
    auto add = [&](int idx, size_t numel) { z.ptr[n] = G[idx]; z.n[n] = (unsigned)numel; ++n; };
    add(P_MASK_TOKEN, p.D);
    add(P_CONV_W, (size_t)p.T * p.C);
    for (int l = 0; l < kDepth; ++l) {
        const LayerPlan& L = p.L[l];
        add(L.p_na_w, L.d); add(L.p_na_b, L.d);
        add(L.p_qkv, (size_t)3 * p.inner * L.d);
        add(L.p_out_w, (size_t)L.d * p.inner); add(L.p_out_b, L.d);
        if (!L.last) { add(L.p_nf_w, L.d); add(L.p_nf_b, L.d); }
        add(L.p_fc1_w, (size_t)L.hid * L.d); add(L.p_fc1_b, L.hid);
        add(L.p_fc2_w, (size_t)L.out * L.hid); add(L.p_fc2_b, L.out);
    }
    add(P_REG_W, (size_t)p.NP * (p.F + p.NP)); add(P_REG_B, p.NP);
    z.count = n;
    SCAT_CHECK_CUDA(launch_k(zero_many_kernel, dim3(16, n), dim3(256), 0, st, z));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int check_ws(const HeadPlan& p, void* ws, size_t bytes) {
    SCAT_REQUIRE(ws != nullptr, kErrWorkspace, "workspace is null");
    SCAT_REQUIRE(bytes >= p.total * sizeof(float), kErrWorkspace, "workspace too small: %zu < %zu bytes", bytes,
                 p.total * sizeof(float));
    SCAT_REQUIRE(((uintptr_t)ws & 255) == 0, kErrWorkspace, "workspace must be 256-byte aligned");
    return 0;
}

int head_forward(const ScatHeadDesc& d, const float* const* W, const float* pe, const float* mean_params,
                 const int32_t* mask_idx, const void* x2, const float* main_feat, float* pred, float* fv, float* pl,
                 void* workspace, size_t ws_bytes, cudaStream_t st, bool defer_pl = false, bool skip_regressor = false) {
    HeadPlan p;
    SCAT_PROPAGATE(make_plan(d, p));
    SCAT_PROPAGATE(check_ws(p, workspace, ws_bytes));
    SCAT_REQUIRE(d.channels > 0 && d.n_out == 66 && d.n_tokens == 21, kErrUnsupported,
                 "head_forward: needs channels>0, n_out=66, n_tokens=21");
    SCAT_REQUIRE(W && x2 && main_feat && pred && fv && mean_params, kErrBadArg, "head_forward: null tensor");
    SCAT_REQUIRE(!d.pl_reg || pl, kErrBadArg, "head_forward: pl_reg set but pl_term is null");
    SCAT_REQUIRE(!d.pos_embed || pe, kErrBadArg, "head_forward: pos_embed set but pe is null");
    float* ws = (float*)workspace;
    // with pos_embed == 0 the reference's token matrix is a view of feat_visual (hand_net.py:364): alias it
    float* X0 = d.pos_embed ? ws + p.L[0].X : fv;
    const bool tc = d.precision != PREC_FP32;
    SideScope side_scope;
    SideStream* sd = side_scope.sd;
    const cudaStream_t sg = sd ? sd->s : st;
    // side stream: the iteration-invariant half of the regressor (needs only main_feat) and, below, the weight copies.
    // (They and the conv forward do not overlap on the GPU: whichever is dispatched first holds the SMs for 10-15 us.
    // Copies first is the better order -- they are needed by the first GEMM, 35 us into the step.)
    SCAT_PROPAGATE(order_after(sd, st, sg));
    SCAT_PROPAGATE(launch_regressor_hoist(main_feat, W[P_REG_W], W[P_REG_B], ws + p.hreg, p.B, p.F, p.NP, sg));
    if (tc) {
        // per-forward weight copies for the tensor cores: the conv weight stacks first, on the main stream; the
        // transformer's copies (TF32-rounded fp32 or bf16) overlap the conv kernel on the side stream
        SCAT_PROPAGATE(launch_conv_weight_prep(W[P_CONV_W], ws + p.w_conv, p.C, p.T, d.x2_dtype, st));
        SCAT_PROPAGATE(round_weights(p, W, ws, d.precision, sg));
        SCAT_PROPAGATE(launch_conv_pe_mask_fwd_tc(x2, d.x2_dtype, ws + p.w_conv, pe, W[P_MASK_TOKEN], mask_idx, d.n_masked,
                                                  d.pos_embed, fv, X0, p.B, p.C, p.D, p.T, st));
    } else {
        SCAT_PROPAGATE(launch_conv_pe_mask_fwd((const float*)x2, W[P_CONV_W], pe, W[P_MASK_TOKEN], mask_idx, d.n_masked,
                                               d.pos_embed, fv, X0, p.B, p.C, p.D, p.T, st));
    }
    SCAT_PROPAGATE(order_after(sd, sg, st));       // weight copies (and the hoisted regressor product) are in place
    SCAT_PROPAGATE(transformer_forward(p, W, ws, d.precision, st, d.pos_embed ? nullptr : fv));
    if (!skip_regressor)     // (the fused train step runs the regressor inside its tail kernel, see head_backward)
        SCAT_PROPAGATE(launch_regressor_fwd(main_feat, ws + p.feat_out, mean_params, W[P_REG_W], W[P_REG_B], pred,
                                            ws + p.states, ws + p.hreg, p.B, p.F, p.NP, p.it, 1, st, /*hoisted=*/1));
    if (d.pl_reg && !defer_pl) {
        // autograd.grad(sum(feat_out), feat_visual) (hand_net.py:396): dgrad-only sweep with a ones cotangent
        SCAT_CHECK_CUDA(launch_k(fill_kernel, dim3(64), dim3(256), 0, st, ws + p.ones, 1.0f, (long long)p.M * 3));
        SCAT_CHECK_LAUNCH();
        SCAT_PROPAGATE(transformer_backward(p, W, nullptr, ws, d.precision, ws + p.ones, st, d.pos_embed ? nullptr : fv));
        // masked tokens do not depend on feat_visual (zero rows) unless the overwrite aliased feat_visual itself
        SCAT_PROPAGATE(launch_mask_bwd(ws + p.dX, mask_idx, d.n_masked, d.pos_embed ? 0 : 1, pl, nullptr, p.B, p.T, p.D, st));
    }
    return 0;
}

// The fused train step hands its tail to head_backward: regressor forward, loss gradient and regressor backward are one
// kernel per sample (launch_regressor_train), the loss values are computed on the side stream.
struct TrainTail {
    const float* mean_params; const float* labels; int ld_labels; float w3d, w2d, grad_scale; float* pred; float* losses;
};

int head_backward(const ScatHeadDesc& d, const float* const* W, const int32_t* mask_idx, const void* x2,
                  const float* main_feat, const float* g_pred, const float* g_fv, float* const* G, void* x2_grad,
                  float* mf_grad, void* workspace, size_t ws_bytes, cudaStream_t st, const float* fv_alias,
                  float* pl_out = nullptr /* non-null: also sweep the path-length cotangent (stacked) into pl_out */,
                  int phase = -1 /* -1: everything; 0: down to transformer layer 1; 1: layer 0 and masking; 2: conv */,
                  const TrainTail* tail = nullptr, GradHook* hook = nullptr) {
    HeadPlan p;
    SCAT_PROPAGATE(make_plan(d, p));
    SCAT_PROPAGATE(check_ws(p, workspace, ws_bytes));
    SCAT_REQUIRE(W && G && x2 && main_feat && (g_pred || tail), kErrBadArg, "head_backward: null tensor");
    SCAT_REQUIRE(d.pos_embed || fv_alias, kErrBadArg, "head_backward: pos_embed==0 needs the forward's feat_visual");
    float* ws = (float*)workspace;
    SideScope side_scope;
    SideStream* sd = side_scope.sd;
    const cudaStream_t sg = sd ? sd->s : st;
    const int sweeps = pl_out ? 2 : 1;
    float* up = pl_out ? ws + p.up2 : ws + p.dfeat;        // [sweeps*M, 3]: real cotangent first
    cudaEvent_t loss_done = nullptr;
    if (phase <= 0) {
    // every parameter gradient is accumulated into (split-K / column-sum / LayerNorm reductions): clear them once,
    // on the side stream, which then also carries the regressor weight gradients
    SCAT_PROPAGATE(order_after(sd, st, sg));
    SCAT_PROPAGATE(zero_param_grads(p, G, sg));
    cudaEvent_t zeroed = nullptr;
    SCAT_PROPAGATE(side_mark(sd, sg, &zeroed));
    if (tail != nullptr) {
        // regressor forward + d loss / d pred + regressor backward (+ the ones cotangent of the stacked sweep): one kernel
        SCAT_PROPAGATE(launch_regressor_train(ws + p.feat_out, tail->mean_params, W[P_REG_W], ws + p.hreg, tail->labels,
                                              tail->ld_labels, tail->w3d, tail->w2d, tail->grad_scale, tail->pred, ws + p.states,
                                              up, pl_out ? up + (size_t)p.M * 3 : nullptr, ws + p.gsum, ws + p.gsteps, p.B, p.F,
                                              p.NP, p.it, st));
    } else {
        // regressor + root-relative backward
        SCAT_PROPAGATE(launch_regressor_bwd(g_pred, W[P_REG_W], up, mf_grad, ws + p.gsum, ws + p.gsteps, p.B, p.F,
                                            p.NP, p.it, 1, st, /*skip_main_feat_gemm=*/1));
        if (pl_out) {
            SCAT_CHECK_CUDA(launch_k(fill_kernel, dim3(64), dim3(256), 0, st, up + (size_t)p.M * 3, 1.0f, (long long)p.M * 3));
            SCAT_CHECK_LAUNCH();
        }
    }
    {
        SCAT_PROPAGATE(order_after(sd, st, sg));      // gsum / gsteps (and pred) are ready
        if (tail != nullptr)                          // the loss values need a batch reduction: off the critical path
        {
            SCAT_PROPAGATE(launch_proj_loss(tail->pred, tail->labels, tail->ld_labels, nullptr, p.T * p.D, p.T, tail->w3d,
                                            tail->w2d, tail->grad_scale, tail->losses, nullptr, ws + p.pl_scratch, p.B, sg));
            SCAT_PROPAGATE(side_mark(sd, sg, &loss_done));
        }
        // d main_feat = gsum Wr[:, :F] (nothing downstream reads it), dWr[:, :F] = gsum^T main_feat, dWr[:, F:] = sum over
        // samples and steps of g_step (x) state, d br = column sums of gsum: one launch
        SCAT_PROPAGATE(launch_regressor_param_grads(ws + p.gsum, ws + p.gsteps, ws + p.states, main_feat, W[P_REG_W], mf_grad,
                                                    G[P_REG_W], G[P_REG_B], p.B, p.F, p.NP, p.it, sg));
        SCAT_PROPAGATE(side_wait(st, zeroed));        // the main stream reduces into the gradients from here on
    }
    }
    // the gradients of layers 2, 1 and of the regressor are final after phase 0: a data-parallel caller can start
    // reducing that part of the bucket while phase 1 runs; layer 0 (the bulk of the bucket) is final after phase 1 and
    // its reduction overlaps the conv backward of phase 2
    if (phase != 2) {
    const int l_first = phase == 1 ? 0 : kDepth - 1, l_last = phase == 0 ? 1 : 0;
    SCAT_PROPAGATE(transformer_backward(p, W, G, ws, d.precision, up, st, d.pos_embed ? nullptr : fv_alias, sweeps, sd,
                                        l_first, l_last, hook));
    // the side stream still carries the last layer's LayerNorm parameter gradients: a phase joins them before it returns
    // (its part of the gradient bucket is final then); the single call joins at the very end, behind the conv passes
    if (phase == 0) return join_side(sd, st);
    // through masking / positional encoding into the conv output.  Tensor-core precisions without an external
    // feat_visual cotangent: masking, d mask_token and the split operand of the conv passes are ONE pass (phase 2)
    const bool fused_prep = d.precision != PREC_FP32 && g_fv == nullptr;
    if (!fused_prep)
        SCAT_PROPAGATE(launch_mask_bwd(ws + p.dX, mask_idx, d.n_masked, 0, ws + p.dFv, G[P_MASK_TOKEN], p.B, p.T, p.D, st, 1));
    if (pl_out) {
        // second half of the stacked sweep is d(sum feat_out)/d feat_visual (hand_net.py:396); nothing on the chain reads
        // it, so it (and, in the fused step, loss += 10 * l_pl, losses[3] = l_pl, train.py:178-183,201) runs on the second
        // side stream under the conv passes
        const cudaStream_t sp = (sd && tail != nullptr) ? sd->s2 : st;
        SCAT_PROPAGATE(order_after(sd, st, sp));
        SCAT_PROPAGATE(launch_mask_bwd(ws + p.dX + (size_t)p.M * p.D, mask_idx, d.n_masked, d.pos_embed ? 0 : 1, pl_out,
                                       nullptr, p.B, p.T, p.D, sp));
        if (tail != nullptr) {
            SCAT_PROPAGATE(side_wait(sp, loss_done));     // the loss kernel wrote losses[0..2] (and used the same scratch)
            SCAT_PROPAGATE(launch_pl_loss_add(pl_out, p.T * p.D, p.T, tail->losses, ws + p.pl_scratch, p.B, sp));
        }
    }
    if (g_fv != nullptr) {
        SCAT_CHECK_CUDA(launch_k(add_inplace_kernel, dim3(148 * 4), dim3(256), 0, st, ws + p.dFv, g_fv, (long long)p.M * p.D));
        SCAT_CHECK_LAUNCH();
    }
    if (phase == 1 && fused_prep)    // phased issue: the mask-token gradient belongs to the part of the bucket phase 1 completes
        SCAT_PROPAGATE(launch_conv_bwd_prep(ws + p.dX, mask_idx, d.n_masked, ws + p.dFv2, G[P_MASK_TOKEN], p.B, p.T, p.D,
                                            d.x2_dtype, st));
    if (phase == 1) return join_side(sd, st);
    }
    if (d.precision != PREC_FP32) {
        if (g_fv != nullptr)          // dFv = masked dX + external cotangent was formed above: split it (nothing left to mask)
            SCAT_PROPAGATE(launch_conv_bwd_prep(ws + p.dFv, nullptr, 0, ws + p.dFv2, nullptr, p.B, p.T, p.D, d.x2_dtype, st));
        else if (phase != 2)          // (phase 2 of a phased issue: phase 1 already ran the fused pass)
            SCAT_PROPAGATE(launch_conv_bwd_prep(ws + p.dX, mask_idx, d.n_masked, ws + p.dFv2, G[P_MASK_TOKEN], p.B, p.T, p.D,
                                                d.x2_dtype, st));
        // two persistent one-CTA-per-SM streams (x2 in, x2.grad out): back to back on the main stream
        SCAT_PROPAGATE(launch_conv_wgrad_tc(ws + p.dFv2, x2, d.x2_dtype, G[P_CONV_W], p.B, p.C, p.D, p.T, st));   // G was zeroed above
        SCAT_PROPAGATE(fire_hook(hook, sd, st, 3));      // mask token + conv weight are final: exchanged under the dgrad stream
        if (x2_grad != nullptr)
            SCAT_PROPAGATE(launch_conv_dgrad_tc(ws + p.dFv2, ws + p.w_conv, d.x2_dtype, x2_grad, p.B, p.C, p.D, p.T, st));
    } else {
        if (x2_grad != nullptr)
            SCAT_PROPAGATE(launch_conv_dgrad(ws + p.dFv, W[P_CONV_W], (float*)x2_grad, p.B, p.C, p.D, p.T, st));
        SCAT_PROPAGATE(launch_conv_wgrad(ws + p.dFv, (const float*)x2, G[P_CONV_W], ws + p.conv_scratch, p.B, p.C, p.D, p.T, st));
        SCAT_PROPAGATE(fire_hook(hook, sd, st, 3));
    }
    if (hook != nullptr && hook->used && sd != nullptr) SCAT_PROPAGATE(order_after(sd, sd->sx, st));
    return join_side(sd, st);
}

}  // namespace
}  // namespace scat

// =============================================================================================
// C ABI
// =============================================================================================
using namespace scat;

extern "C" {

int scat_abi_version(void) { return SCAT_B200_ABI_VERSION; }
const char* scat_last_error_string(void) { return last_error(); }
uint64_t scat_launch_count(void) { return g_launch_count; }

size_t scat_head_workspace_bytes(const ScatHeadDesc* desc) {
    if (!desc) return 0;
    HeadPlan p;
    if (make_plan(*desc, p) != 0) return 0;
    return p.total * sizeof(float);
}

int scat_head_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe, const float* mean_params,
                      const int32_t* mask_idx, const void* x2, const float* main_feat, float* pred_params,
                      float* feat_visual, float* pl_term, void* workspace, size_t workspace_bytes, void* stream) {
    SCAT_REQUIRE(desc, kErrBadArg, "desc is null");
    return head_forward(*desc, params, pe, mean_params, mask_idx, x2, main_feat, pred_params, feat_visual, pl_term,
                        workspace, workspace_bytes, (cudaStream_t)stream);
}

int scat_head_backward(const ScatHeadDesc* desc, const float* const* params, const int32_t* mask_idx, const void* x2,
                       const float* main_feat, const float* feat_visual, const float* grad_pred,
                       const float* grad_feat_visual, float* const* grads, void* x2_grad, float* main_feat_grad,
                       void* workspace, size_t workspace_bytes, void* stream) {
    SCAT_REQUIRE(desc, kErrBadArg, "desc is null");
    return head_backward(*desc, params, mask_idx, x2, main_feat, grad_pred, grad_feat_visual, grads, x2_grad,
                         main_feat_grad, workspace, workspace_bytes, (cudaStream_t)stream, feat_visual);
}

int scat_proj_loss(int32_t batch, int32_t n_tokens, int32_t token_dim, const float* pred_params, const float* labels,
                   int32_t ld_labels, const float* pl_term, float l_weight_3d, float l_weight_2d, float grad_scale,
                   float* losses, float* grad_pred, float* scratch, void* stream) {
    return launch_proj_loss(pred_params, labels, ld_labels, pl_term, n_tokens * token_dim, n_tokens, l_weight_3d,
                            l_weight_2d, grad_scale, losses, grad_pred, scratch, batch, (cudaStream_t)stream);
}

static int head_train_step_impl(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                                const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                                const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d, float grad_scale,
                                float* pred_params, float* feat_visual, float* pl_term, float* losses, float* const* grads,
                                void* x2_grad, float* main_feat_grad, void* workspace, size_t workspace_bytes, void* stream,
                                int phase, GradHook* hook = nullptr) {
    SCAT_REQUIRE(desc, kErrBadArg, "desc is null");
    SCAT_REQUIRE(phase >= -1 && phase <= 2, kErrBadArg, "train_step: phase %d", phase);
    cudaStream_t st = (cudaStream_t)stream;
    // The path-length VJP is a dgrad-only sweep of the same graph as the backward and its result carries no
    // gradient (hand_net.py:396, train.py:201), so it rides along with the real cotangent: one stacked sweep.
    const bool pl = desc->pl_reg != 0;
    HeadPlan p;
    SCAT_PROPAGATE(make_plan(*desc, p));
    float* ws = (float*)workspace;
    // regressor forward, loss gradient and regressor backward as one kernel per sample whenever the shapes are the head's
    const bool fused_tail = desc->iteration >= 1 && desc->n_out == 66 && (ld_labels == 105 || ld_labels == 166);
    const TrainTail tail{mean_params, labels, ld_labels, l_weight_3d, l_weight_2d, grad_scale, pred_params, losses};
    if (phase <= 0) {
        SCAT_REQUIRE(labels && losses && pred_params, kErrBadArg, "train_step: labels / losses / pred_params is null");
        SCAT_PROPAGATE(head_forward(*desc, params, pe, mean_params, mask_idx, x2, main_feat, pred_params, feat_visual,
                                    pl_term, workspace, workspace_bytes, st, /*defer_pl=*/true, /*skip_regressor=*/fused_tail));
        if (!fused_tail)
            SCAT_PROPAGATE(launch_proj_loss(pred_params, labels, ld_labels, nullptr, p.T * p.D, p.T, l_weight_3d, l_weight_2d,
                                            grad_scale, losses, ws + p.g_pred, ws + p.pl_scratch, p.B, st));
    }
    SCAT_PROPAGATE(head_backward(*desc, params, mask_idx, x2, main_feat, ws + p.g_pred, nullptr, grads, x2_grad,
                                 main_feat_grad, workspace, workspace_bytes, st, feat_visual, pl ? pl_term : nullptr, phase,
                                 fused_tail ? &tail : nullptr, hook));
    if (pl && !fused_tail && (phase == -1 || phase == 1))       // loss += 10 * l_pl, losses[3] = l_pl (train.py:178-183,201)
        SCAT_PROPAGATE(launch_pl_loss_add(pl_term, p.T * p.D, p.T, losses, ws + p.pl_scratch, p.B, st));
    return 0;
}

int scat_head_train_step(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                         const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                         const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d, float grad_scale,
                         float* pred_params, float* feat_visual, float* pl_term, float* losses, float* const* grads,
                         void* x2_grad, float* main_feat_grad, void* workspace, size_t workspace_bytes, void* stream) {
    return head_train_step_impl(desc, params, pe, mean_params, mask_idx, x2, main_feat, labels, ld_labels, l_weight_3d,
                                l_weight_2d, grad_scale, pred_params, feat_visual, pl_term, losses, grads, x2_grad,
                                main_feat_grad, workspace, workspace_bytes, stream, -1);
}

int scat_head_train_step_phase(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                               const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                               const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d,
                               float grad_scale, float* pred_params, float* feat_visual, float* pl_term, float* losses,
                               float* const* grads, void* x2_grad, float* main_feat_grad, void* workspace,
                               size_t workspace_bytes, void* stream, int32_t phase) {
    return head_train_step_impl(desc, params, pe, mean_params, mask_idx, x2, main_feat, labels, ld_labels, l_weight_3d,
                                l_weight_2d, grad_scale, pred_params, feat_visual, pl_term, losses, grads, x2_grad,
                                main_feat_grad, workspace, workspace_bytes, stream, phase);
}

int scat_head_train_step_hooked(const ScatHeadDesc* desc, const float* const* params, const float* pe,
                                const float* mean_params, const int32_t* mask_idx, const void* x2, const float* main_feat,
                                const float* labels, int32_t ld_labels, float l_weight_3d, float l_weight_2d,
                                float grad_scale, float* pred_params, float* feat_visual, float* pl_term, float* losses,
                                float* const* grads, void* x2_grad, float* main_feat_grad, void* workspace,
                                size_t workspace_bytes, scat_grads_ready_fn ready, void* ready_user, void* stream) {
    GradHook hook;
    hook.fn = ready; hook.user = ready_user;
    return head_train_step_impl(desc, params, pe, mean_params, mask_idx, x2, main_feat, labels, ld_labels, l_weight_3d,
                                l_weight_2d, grad_scale, pred_params, feat_visual, pl_term, losses, grads, x2_grad,
                                main_feat_grad, workspace, workspace_bytes, stream, -1, ready ? &hook : nullptr);
}

int scat_tokens_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe, const int32_t* mask_idx,
                        const float* tokens, float* out, float* mean, void* workspace, size_t workspace_bytes,
                        void* stream) {
    SCAT_REQUIRE(desc && params && tokens && out, kErrBadArg, "tokens_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    HeadPlan p;
    SCAT_PROPAGATE(make_plan(*desc, p));
    SCAT_PROPAGATE(check_ws(p, workspace, workspace_bytes));
    float* ws = (float*)workspace;
    if (desc->precision != PREC_FP32) SCAT_PROPAGATE(round_weights(p, params, ws, desc->precision, st));
    SCAT_PROPAGATE(launch_pe_mask_tokens(tokens, pe, params[P_MASK_TOKEN], mask_idx, desc->n_masked, desc->pos_embed,
                                         ws + p.L[0].X, p.B, p.T, p.D, st));
    SCAT_PROPAGATE(transformer_forward(p, params, ws, desc->precision, st, nullptr));
    SCAT_CHECK_CUDA(cudaMemcpyAsync(out, ws + p.feat_out, (size_t)p.M * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (mean != nullptr) {
        SCAT_CHECK_CUDA(launch_k(token_mean_kernel, dim3(p.B), dim3(32), 0, st, ws + p.feat_out, mean, p.T));
        SCAT_CHECK_LAUNCH();
    }
    return 0;
}

int scat_coarse_forward(const ScatHeadDesc* desc, const float* const* params, const float* pe, const float* mean_params,
                        const int32_t* mask_idx, const void* x2, const float* main_feat, float* pred_params,
                        float* feat_visual, float* attn, void* workspace, size_t workspace_bytes, void* stream) {
    SCAT_REQUIRE(desc && params && x2 && main_feat && pred_params && feat_visual && attn && mean_params, kErrBadArg,
                 "coarse_forward: null argument");
    const ScatHeadDesc& d = *desc;
    cudaStream_t st = (cudaStream_t)stream;
    HeadPlan p;
    SCAT_PROPAGATE(make_plan(d, p));
    SCAT_PROPAGATE(check_ws(p, workspace, workspace_bytes));
    SCAT_REQUIRE(d.channels > 0 && d.n_out == 66 && d.n_tokens == 21, kErrUnsupported,
                 "coarse_forward: needs channels>0, n_out=66, n_tokens=21");
    SCAT_REQUIRE(!d.pl_reg, kErrUnsupported, "coarse_forward: inference path, the path-length term needs a backward (pl_reg must be 0)");
    SCAT_REQUIRE(d.precision == PREC_FP32 || d.precision == PREC_TF32, kErrUnsupported, "coarse_forward: precision fp32 or tf32");
    SCAT_REQUIRE(!d.pos_embed || pe, kErrBadArg, "coarse_forward: pos_embed set but pe is null");
    float* ws = (float*)workspace;
    float* X0 = d.pos_embed ? ws + p.L[0].X : feat_visual;       // same aliasing as hand_net.py:364,373 (here :268-269,280)
    if (d.precision != PREC_FP32) {
        SCAT_PROPAGATE(launch_conv_weight_prep(params[P_CONV_W], ws + p.w_conv, p.C, p.T, d.x2_dtype, st));
        SCAT_PROPAGATE(round_weights(p, params, ws, d.precision, st));
        SCAT_PROPAGATE(launch_conv_pe_mask_fwd_tc(x2, d.x2_dtype, ws + p.w_conv, pe, params[P_MASK_TOKEN], mask_idx, d.n_masked,
                                                  d.pos_embed, feat_visual, X0, p.B, p.C, p.D, p.T, st));
    } else {
        SCAT_PROPAGATE(launch_conv_pe_mask_fwd((const float*)x2, params[P_CONV_W], pe, params[P_MASK_TOKEN], mask_idx, d.n_masked,
                                               d.pos_embed, feat_visual, X0, p.B, p.C, p.D, p.T, st));
    }
    SCAT_PROPAGATE(transformer_forward(p, params, ws, d.precision, st, d.pos_embed ? nullptr : feat_visual, /*attn_variant=*/true));
    // the attention maps of the LAST layer are the second output (vision_transformer_attn.py:113)
    SCAT_CHECK_CUDA(cudaMemcpyAsync(attn, ws + p.L[kDepth - 1].P, (size_t)p.B * p.heads * p.T * p.T * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
    SCAT_CHECK_CUDA(launch_k(coarse_tail_kernel, dim3(ceil_div(p.B, 4)), dim3(128), 0, st, main_feat, (const float*)(ws + p.feat_out),
                             mean_params, params[P_REG_W], params[P_REG_B], pred_params, p.B, p.F));
    SCAT_CHECK_LAUNCH();
    return 0;
}

// ---- single operators ----------------------------------------------------------------------------
int scat_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk, float* C, int32_t ldc,
              int32_t M, int32_t N, int32_t K, int32_t epilogue, const float* bias, const float* aux_in,
              int32_t ld_aux_in, float* aux_out, int32_t ld_aux_out, int32_t accumulate, int32_t precision,
              void* stream) {
    GemmArgs g;
    g.A = A; g.sam = sam; g.sak = sak; g.B = B; g.sbn = sbn; g.sbk = sbk; g.C = C; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.epilogue = epilogue; g.bias = bias; g.aux_in = aux_in; g.ld_aux_in = ld_aux_in;
    g.aux_out = aux_out; g.ld_aux_out = ld_aux_out; g.accumulate = accumulate;
    g.prerounded = (precision & SCAT_PREC_FLAG_PREROUNDED) ? 1 : 0;
    g.allow_wide = 1;
    if (precision & SCAT_PREC_FLAG_SPLIT_K) { g.allow_split_k = 1; g.c_zeroed = 1; }
    precision &= ~(SCAT_PREC_FLAG_PREROUNDED | SCAT_PREC_FLAG_SPLIT_K);
    if (precision == PREC_FP32) return launch_gemm_simt(g, (cudaStream_t)stream);
    if (precision == SCAT_PREC_TF32X3) return launch_gemm_mma3(g, (cudaStream_t)stream);
    SCAT_REQUIRE(gemm_tc_supported(g), kErrUnsupported,
                 "scat_gemm: operand layout not expressible as TMA tensor maps (16-byte strides, unit inner stride)");
    return launch_gemm_tc(g, precision, (cudaStream_t)stream);   // operands rounded to TF32-nearest inside the kernel
}

void scat_debug_gemm_timeline(void* dev_int64x8) { gemm_tc_set_debug_buffer((long long*)dev_int64x8); }

int scat_gemm_bf16(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk, float* C,
                   int32_t ldc, void* C16, int32_t ldc16, int32_t M, int32_t N, int32_t K, int32_t epilogue,
                   const float* bias, const float* aux_in, int32_t ld_aux_in, float* aux_out, int32_t ld_aux_out,
                   int32_t split_k, void* stream) {
    GemmArgs g;
    g.A = A; g.sam = sam; g.sak = sak; g.B = B; g.sbn = sbn; g.sbk = sbk; g.operand_bf16 = 1;
    g.C = C; g.ldc = ldc; g.C16 = C16; g.ldc16 = ldc16;
    g.M = M; g.N = N; g.K = K; g.epilogue = epilogue; g.bias = bias; g.aux_in = aux_in; g.ld_aux_in = ld_aux_in;
    g.aux_out = aux_out; g.ld_aux_out = ld_aux_out;
    g.allow_split_k = split_k ? 1 : 0; g.c_zeroed = split_k ? 1 : 0;
    g.allow_wide = 1;
    SCAT_REQUIRE(gemm_tc_supported(g), kErrUnsupported,
                 "scat_gemm_bf16: operand layout not expressible as TMA tensor maps (16-byte strides, unit inner stride)");
    return launch_gemm_tc(g, PREC_BF16, (cudaStream_t)stream);
}

int scat_conv_pe_mask_fwd(const float* x2, const float* conv_w, const float* pe, const float* mask_token,
                          const int32_t* mask_idx, int32_t n_masked, int32_t pos_embed, float* feat_visual,
                          float* tokens_out, int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens,
                          void* stream) {
    return launch_conv_pe_mask_fwd(x2, conv_w, pe, mask_token, mask_idx, n_masked, pos_embed, feat_visual, tokens_out,
                                   batch, channels, hw, n_tokens, (cudaStream_t)stream);
}

// ---- tensor-core front end (conv_tc.cu) as single operators: what the head runs in TF32 / BF16 mode ----
size_t scat_conv_tc_scratch_floats(int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens) {
    // [weight stacks][split d tokens]
    return round_up(conv_weight_prep_floats(channels, n_tokens), 64) + conv_split_floats(batch, hw, n_tokens);
}

int scat_conv_pe_mask_fwd_tc(const void* x2, int32_t x2_dtype, const float* conv_w, const float* pe, const float* mask_token,
                             const int32_t* mask_idx, int32_t n_masked, int32_t pos_embed, float* feat_visual,
                             float* tokens_out, float* scratch, int32_t batch, int32_t channels, int32_t hw,
                             int32_t n_tokens, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SCAT_REQUIRE(scratch && x2 && conv_w && feat_visual && tokens_out, kErrBadArg, "conv_fwd_tc: null argument");
    SCAT_REQUIRE(x2_dtype == SCAT_DTYPE_F32 || x2_dtype == SCAT_DTYPE_BF16, kErrBadArg, "conv_fwd_tc: x2_dtype %d", x2_dtype);
    SCAT_PROPAGATE(launch_conv_weight_prep(conv_w, scratch, channels, n_tokens, x2_dtype, st));
    return launch_conv_pe_mask_fwd_tc(x2, x2_dtype, scratch, pe, mask_token, mask_idx, n_masked, pos_embed, feat_visual,
                                      tokens_out, batch, channels, hw, n_tokens, st);
}

int scat_conv_bwd_tc(const float* d_tokens, const void* x2, int32_t x2_dtype, const float* conv_w, const int32_t* mask_idx,
                     int32_t n_masked, void* x2_grad, float* conv_w_grad, float* mask_token_grad, float* scratch,
                     int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SCAT_REQUIRE(scratch && d_tokens && x2 && conv_w && conv_w_grad, kErrBadArg, "conv_bwd_tc: null argument");
    SCAT_REQUIRE(x2_dtype == SCAT_DTYPE_F32 || x2_dtype == SCAT_DTYPE_BF16, kErrBadArg, "conv_bwd_tc: x2_dtype %d", x2_dtype);
    float* w_prep = scratch;
    float* dsplit = scratch + round_up(conv_weight_prep_floats(channels, n_tokens), 64);
    SCAT_PROPAGATE(launch_conv_weight_prep(conv_w, w_prep, channels, n_tokens, x2_dtype, st));
    if (mask_token_grad) SCAT_CHECK_CUDA(cudaMemsetAsync(mask_token_grad, 0, (size_t)hw * sizeof(float), st));
    SCAT_PROPAGATE(launch_conv_bwd_prep(d_tokens, mask_idx, n_masked, dsplit, mask_token_grad, batch, n_tokens, hw, x2_dtype, st));
    SCAT_CHECK_CUDA(cudaMemsetAsync(conv_w_grad, 0, (size_t)n_tokens * channels * sizeof(float), st));
    SCAT_PROPAGATE(launch_conv_wgrad_tc(dsplit, x2, x2_dtype, conv_w_grad, batch, channels, hw, n_tokens, st));
    if (x2_grad) SCAT_PROPAGATE(launch_conv_dgrad_tc(dsplit, w_prep, x2_dtype, x2_grad, batch, channels, hw, n_tokens, st));
    return 0;
}

size_t scat_conv_bwd_scratch_floats(int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens) {
    return conv_wgrad_scratch_floats(channels, n_tokens) + (size_t)batch * n_tokens * hw;
}

int scat_conv_bwd(const float* d_tokens, const float* x2, const float* conv_w, const int32_t* mask_idx,
                  int32_t n_masked, float* x2_grad, float* conv_w_grad, float* mask_token_grad, float* scratch,
                  int32_t batch, int32_t channels, int32_t hw, int32_t n_tokens, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SCAT_REQUIRE(scratch, kErrBadArg, "conv_bwd: scratch is null");
    float* dFv = scratch + conv_wgrad_scratch_floats(channels, n_tokens);
    SCAT_PROPAGATE(launch_mask_bwd(d_tokens, mask_idx, n_masked, 0, dFv, mask_token_grad, batch, n_tokens, hw, st));
    if (x2_grad) SCAT_PROPAGATE(launch_conv_dgrad(dFv, conv_w, x2_grad, batch, channels, hw, n_tokens, st));
    return launch_conv_wgrad(dFv, x2, conv_w_grad, scratch, batch, channels, hw, n_tokens, st);
}

int scat_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                       int32_t rows, int32_t dim, void* stream) {
    return launch_layernorm_fwd(x, dim, gamma, beta, y, dim, mean, rstd, rows, dim, 0, (cudaStream_t)stream);
}

int scat_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* resid, float* dx, float* dgamma, float* dbeta, int32_t rows, int32_t dim,
                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dgamma) {
        SCAT_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, dim * sizeof(float), st));
        SCAT_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, dim * sizeof(float), st));
    }
    return launch_layernorm_bwd(dy, dim, x, dim, gamma, mean, rstd, resid, dim, dx, dim, dgamma, dbeta, rows, dim, 0, st);
}

int scat_attention_fwd(const float* qkv, float* o, float* p, int32_t batch, int32_t n, int32_t heads, void* stream) {
    return launch_attention_fwd(qkv, o, p, batch, n, heads, 0, (cudaStream_t)stream);
}

int scat_attention_bwd(const float* qkv, const float* p, const float* d_o, float* d_qkv, int32_t batch, int32_t n,
                       int32_t heads, void* stream) {
    return launch_attention_bwd(qkv, p, d_o, d_qkv, batch, n, heads, 0, (cudaStream_t)stream);
}

int scat_attention_fwd_tc(const float* qkv, float* o, float* p, int32_t batch, int32_t n, int32_t heads, void* stream) {
    if (n == 128)      // tcgen05 kernel of the token variant (inference: p is left untouched); OUT_TF32 = its in-head output mode
        return launch_attention_tc128_fwd(qkv, o, batch, n, heads, OUT_F32, (cudaStream_t)stream);
    return launch_attention_mma_fwd(qkv, o, p, batch, n, heads, OUT_F32, (cudaStream_t)stream);
}

int scat_attention_bwd_tc(const float* qkv, const float* p, const float* d_o, float* d_qkv, int32_t batch, int32_t n,
                          int32_t heads, void* stream) {
    return launch_attention_mma_bwd(qkv, p, d_o, d_qkv, batch, n, heads, OUT_F32, (cudaStream_t)stream);
}

int scat_regressor_fwd(const float* main_feat, const float* feat_out, const float* mean_params, const float* w,
                       const float* b, float* pred, float* states, float* scratch, int32_t batch, int32_t feat_dim,
                       int32_t n_out, int32_t iteration, int32_t root_relative, void* stream) {
    return launch_regressor_fwd(main_feat, feat_out, mean_params, w, b, pred, states, scratch, batch, feat_dim, n_out,
                                iteration, root_relative, (cudaStream_t)stream);
}

}  // extern "C"
