// tcgen05 / TMEM / TMA GEMM for sm_100a: C[M,N] = epilogue(A * B^T), fp32 accumulation in tensor memory.
//   operands fp32 in HBM  -> tcgen05.mma kind::tf32 (K = 8 per instruction, 32 k per 128-byte swizzle row)
//   operands bf16 in HBM  -> tcgen05.mma kind::f16  (K = 16 per instruction, 64 k per 128-byte swizzle row)
//
// Replaces the cuBLAS calls behind the reference's nn.Linear layers (vision_transformer.py:33-35,53-55) and
// their autograd backward (dgrad / wgrad).  One kernel covers all three because both operands may be
// K-major (row stride, unit k stride) or MN-major (unit row stride, k stride): the majorness only changes the
// TMA box issue pattern, the UMMA shared-memory descriptors and two bits of the instruction descriptor, all
// of them compile-time here.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, 4 per stage)
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b -> registers -> bias / GELU / residual -> global fp32 and/or bf16)
//
// Tile: BM = 128 rows (TMEM lanes) x BN in {64,128} columns x one 128-byte swizzle row of k per stage.  Two CTAs
// are resident per SM (<= 97 KB of shared memory and <= 128 TMEM columns each) so that one CTA's epilogue and
// prologue overlap the other's main loop; the head's GEMMs are small (M = 2016) and one tile per CTA.
// Out-of-range rows / columns / k are zero-filled by TMA, stores are bounds-checked.
//
// The producer and MMA loops are executed by whole warps with elect.sync around the issuing instructions and
// incremental stage / phase / descriptor arithmetic: a single-lane `if (lane == 0)` loop makes the compiler wrap
// every uniform-datapath instruction (UTMALDG, UTCHMMA, UTCBAR) in an election loop, ~150 dependent instructions
// per k-block, which bounded the first version of this kernel at ~0.45 us per k-block.
#include "kernels.h"
#include "tc_ptx.cuh"

namespace scat {
namespace {

constexpr int BM = 128;
constexpr int TC_THREADS = 192;

// ---------------------------------------------------------------------------------------------
// element traits.  UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) |
// SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64).
//   K-major, 128B swizzle (layout 2): 8-row groups 1024 B apart (SBO); a k-slice of one MMA is +32 B inside the row
//   MN-major fp32 (layout 1 = SWIZZLE_128B_BASE32B, "the only smem layout for mn-major tf32"): TMA mode
//     SWIZZLE_128B_ATOM_32B, boxes of 32 mn x 32 k; 32-wide mn groups one box apart (LBO), 4-k-row atoms 512 B apart
//     (SBO), a k-slice of 8 = +1024 B
//   MN-major bf16 (layout 2): boxes of 64 mn x 64 k; 64-wide mn groups one box apart (LBO), 8-k-row atoms 1024 B
//     apart (SBO), a k-slice of 16 = +2048 B
// ---------------------------------------------------------------------------------------------
template <bool BF16>
struct Elem {
    static constexpr int BYTES = BF16 ? 2 : 4;
    static constexpr int BK = 128 / BYTES;            // k per stage (one swizzle row)
    static constexpr int UK = 32 / BYTES;             // k per MMA
    static constexpr int MN_BOX = 128 / BYTES;        // mn elements per 128-byte row of an MN-major box
    static constexpr uint32_t FMT = BF16 ? 1u : 2u;   // instruction-descriptor operand format: BF16 = 1, TF32 = 2
    static constexpr uint32_t MN_LAYOUT = BF16 ? 2u : 1u;
    static constexpr uint32_t MN_SBO = BF16 ? 1024u : 512u;
    static constexpr uint32_t MN_KSTEP = BF16 ? 2048u : 1024u;
    static constexpr uint32_t MN_LBO = BK * 128;      // one box
};

struct TcParams {
    int M, N, K;
    float* C; int ldc;                  // fp32 output (may be null when only the bf16 copy is wanted)
    __nv_bfloat16* C16; int ldc16;      // bf16 copy of the output (null = none): it feeds another bf16 GEMM
    int epilogue;
    const float* bias;
    const float* aux_in; int ld_aux_in; int aux_row_mod;
    float* aux_out; int ld_aux_out;
    int accumulate;
    int round_out;                  // 1: store C rounded to TF32-nearest (it feeds another tensor-core GEMM)
    int round_operands;             // 1: round fp32 operands to TF32 (nearest) in shared memory before the MMA
    int stages;                     // depth of the operand ring (SmemLayout::STAGES or STAGES_DEEP)
    int gelu_saves_grad;            // GemmArgs::gelu_saves_grad
    int exp_skip_gelu;              // timing experiment only (SCAT_EXP_SKIP_GELU=1): GELU / dGELU math replaced by a copy
    int b_static;                   // 1: B does not depend on the preceding kernel (a weight): prefetched before the PDL wait
    int kb_per_split;               // k-blocks per gridDim.z slice (split-K: weight gradients, K = B*21 rows)
    int atomic_out;                 // 1: C += tile with red.global.add (split-K slices combine in L2; C pre-zeroed)
    // batched launches (blockIdx.z = batch index instead of a split-K slice): per-batch TMA coordinate offsets of the
    // operands (rows / k, in elements of the respective tensor-map dimension) and element offsets of the outputs
    int batched;
    int a_row_z, a_k_z, b_row_z, b_k_z;
    long long c_z, aux_out_z;
    const int32_t* mask_idx; int n_masked;   // EPI_PE_MASK: rows (tokens) replaced by bias[n] (the mask token)
    int tiles_m, tiles_n, tiles_z;  // tile space walked by the CTAs (n fastest); z = split-K slice or batch index
    int persistent;                 // a CTA may own several tiles: double-buffered accumulators, dedicated scratch
    int vec_ok;                     // every pointer / leading dimension the epilogue touches allows 16-byte accesses
    long long* dbg;                 // optional: CTA (0,0,0) records clock64() at 8 milestones (tools/gemm_timeline.py)
};
__device__ __forceinline__ void dbg_mark(const TcParams& p, int slot) {
    if (p.dbg != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) p.dbg[slot] = clock64();
}

template <int BN>
struct SmemLayout {
    // ring depth is a launch parameter: 96 KB per CTA (two CTAs per SM) when the launch has more CTAs than SMs, 192 KB
    // when every CTA has an SM to itself -- one CTA's fill rate is ring bytes / TMA latency (profiles/r1_gemm_timeline.txt:
    // 96 KB in flight sustain ~52 B/cycle), so a lone CTA needs the deeper ring to keep its tensor core fed
    // Tiles wider than 128 (BN = 192, 256) only ever run one CTA per SM with the deep ring (single-wave launches).
    static constexpr int STAGES = BN == 64 ? 4 : 3;
    static constexpr int STAGES_DEEP = BN == 64 ? 8 : BN == 128 ? 6 : BN == 192 ? 5 : 4;    // 192 / 192 / 200 / 192 KB
    static constexpr int A_BYTES = BM * 128;              // 16 KB
    static constexpr int B_BYTES = BN * 128;              // 8 / 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 512;                 // 3 x 8 stage barriers + 4 accumulator barriers + tmem slot
    static constexpr int bar_off(int stages) { return stages * STAGE_BYTES; }
    static constexpr int total(int stages, bool persistent) {       // persistent CTAs: + 4 warps x 32x32 floats of epilogue scratch
        return bar_off(stages) + BAR_BYTES + (persistent ? 4 * 32 * 32 * 4 : 0);
    }
};

// ---------------------------------------------------------------------------------------------
// Epilogue of one 128 x BN tile, executed by the four epilogue warps (warp q owns TMEM lanes / tile rows 32q..32q+31).
//
// tcgen05.ld hands lane l the accumulator ROW l of the warp's quadrant; stored from there every global access would
// touch 32 different rows (32 L1 wavefronts per instruction, half-used sectors).  Each warp therefore bounces its
// 32x32 chunk through a private padded scratch in the (now idle) operand ring and continues in a coalesced mapping:
// 8 lanes x float4 cover 128 contiguous bytes of one row, 4 rows per instruction.  Each epilogue warp sits alone on
// its SM sub-partition, so instruction latency is fully exposed: addresses are carried incrementally, flags are
// tested once per chunk (not per element), and all global reads of a chunk (bias, residual / saved pre-activation
// or old C) are issued before the TMEM load so that they overlap it and each other.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_tile_impl(const TcParams& p, uint32_t tmem_lane_base, float* scratch, int row_base,
                                                   int n0, int lane, int BN, int zb, bool mark) {
    // scratch: 32 rows x 32 floats per warp; the float4 chunk c of row r lives at chunk c ^ (r & 7), which keeps both
    // the row-wise writes and the 4-rows-per-instruction reads free of bank conflicts without padding
    const int rsub = lane >> 3, csub = (lane & 7) * 4;
    const int epi = p.epilogue;
    const bool has_aux = epi == EPI_BIAS_RESID || epi == EPI_RESID || epi == EPI_DGELU || (epi == EPI_PE_MASK && p.aux_in != nullptr);
    const bool has_bias = epi == EPI_BIAS || epi == EPI_BIAS_RESID || epi == EPI_BIAS_GELU || (epi == EPI_PE_MASK && p.n_masked > 0);
    const bool rmw = p.accumulate && !p.atomic_out && p.C != nullptr && !has_aux;
    // rows this lane touches: row_base + rsub + 4i, i < 8
    const int row0 = row_base + rsub;
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) valid |= (row0 + 4 * i < p.M ? 1u : 0u) << i;
#define SCAT_ROW_OK(i) ((valid >> (i)) & 1u)
    // the common case gets out early below: plain fp32 store of the accumulator, nothing else to do
    const bool plain = epi == EPI_NONE && !p.round_out && p.C16 == nullptr && !p.atomic_out &&
                       !p.accumulate && p.C != nullptr;
    const long long c_step = 4LL * p.ldc, h_step = 4LL * p.ldc16, z_step = 4LL * p.ld_aux_out, a_step = 4LL * p.ld_aux_in;
    float* cptr = p.C ? p.C + zb * p.c_z + (long long)row0 * p.ldc + n0 + csub : nullptr;
    __nv_bfloat16* hptr = p.C16 ? p.C16 + (long long)row0 * p.ldc16 + n0 + csub : nullptr;
    float* zptr = p.aux_out ? p.aux_out + zb * p.aux_out_z + (long long)row0 * p.ld_aux_out + n0 + csub : nullptr;
    // residual / saved pre-activation rows; with stacked cotangents (aux_row_mod) row m reads activation row m % mod
    const bool aux_linear = p.aux_row_mod <= 0 || row_base + 32 <= p.aux_row_mod;
    const float* aptr = p.aux_in ? p.aux_in + (long long)row0 * p.ld_aux_in + n0 + csub : nullptr;
    // EPI_PE_MASK (conv front end, hand_net.py:366-373): rows are tokens; bit i = this lane's row row0 + 4i is masked
    uint32_t masked = 0;
    if (epi == EPI_PE_MASK) {
        for (int k = 0; k < p.n_masked; ++k) {
            const int t = __ldg(p.mask_idx + k);
#pragma unroll
            for (int i = 0; i < 8; ++i) masked |= (t == row0 + 4 * i ? 1u : 0u) << i;
        }
    }
    const bool has_pe = epi == EPI_PE_MASK && p.aux_in != nullptr;
    float* srow = scratch + lane * 32;                          // TMEM-row mapping: this lane's row of the chunk
    const int cc = lane & 7;                                    // coalesced mapping: chunk cc of rows rsub + 4i
    const float* sread0 = scratch + rsub * 32 + 4 * (cc ^ rsub);            // rows rsub + 8j     ((row & 7) = rsub)
    const float* sread1 = scratch + (rsub + 4) * 32 + 4 * (cc ^ (rsub + 4)); // rows rsub + 4 + 8j ((row & 7) = rsub + 4)
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
        const int n = n0 + c + csub;                 // this lane's 4 columns
        if (n0 + c >= p.N) break;                    // warp-uniform: nothing left in this tile row
        const bool vec = p.vec_ok && n + 4 <= p.N;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ext[8];                               // residual / saved pre-activation, or old C (accumulate)
        if (vec) {
            if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            if (has_aux) {
                if (aux_linear) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (SCAT_ROW_OK(i)) ext[i] = __ldg(reinterpret_cast<const float4*>(aptr + i * a_step));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = row0 + 4 * i;
                        if (SCAT_ROW_OK(i))
                            ext[i] = __ldg(reinterpret_cast<const float4*>(p.aux_in + (long long)(row % p.aux_row_mod) * p.ld_aux_in + n));
                    }
                }
            } else if (rmw) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (SCAT_ROW_OK(i)) ext[i] = *reinterpret_cast<const float4*>(cptr + i * c_step);
            }
        }
        float v[32];
        __syncwarp();                                // tcgen05.ld is warp-collective; scratch reads of the last chunk done
        tmem_ld32(tmem_lane_base + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + 4 * (j ^ (lane & 7))) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (vec) {
            float4 acc[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[2 * j] = *reinterpret_cast<const float4*>(sread0 + j * 8 * 32);
                acc[2 * j + 1] = *reinterpret_cast<const float4*>(sread1 + j * 8 * 32);
            }
            if (plain) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (SCAT_ROW_OK(i)) *reinterpret_cast<float4*>(cptr + i * c_step) = acc[i];
                cptr += 32;
                continue;
            }
            // ---- math phase (registers only, except the saved pre-activation of BIAS_GELU) ----
            if (epi == EPI_PE_MASK) {
                // C <- conv output (masked rows overwritten only when the token matrix aliases it: no aux_out);
                // aux_out <- token matrix: mask token on masked rows, else conv output (+ positional encoding)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (!SCAT_ROW_OK(i)) continue;
                    const bool mk = (masked >> i) & 1;
                    float4 tok = acc[i];
                    if (mk) tok = b4;
                    else if (has_pe) { tok.x += ext[i].x; tok.y += ext[i].y; tok.z += ext[i].z; tok.w += ext[i].w; }
                    if (zptr != nullptr) *reinterpret_cast<float4*>(zptr + i * z_step) = tok;
                    else acc[i] = tok;
                }
            } else if (has_bias) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { acc[i].x += b4.x; acc[i].y += b4.y; acc[i].z += b4.z; acc[i].w += b4.w; }
            }
            if (epi == EPI_BIAS_RESID || epi == EPI_RESID || rmw) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (SCAT_ROW_OK(i)) { acc[i].x += ext[i].x; acc[i].y += ext[i].y; acc[i].z += ext[i].z; acc[i].w += ext[i].w; }
            } else if (epi == EPI_DGELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (SCAT_ROW_OK(i) && !p.exp_skip_gelu) {
                        if (p.gelu_saves_grad) {       // the forward stored gelu'(z)
                            acc[i].x *= ext[i].x; acc[i].y *= ext[i].y; acc[i].z *= ext[i].z; acc[i].w *= ext[i].w;
                        } else {
                            acc[i].x *= gelu_erf_grad(ext[i].x); acc[i].y *= gelu_erf_grad(ext[i].y);
                            acc[i].z *= gelu_erf_grad(ext[i].z); acc[i].w *= gelu_erf_grad(ext[i].w);
                        }
                    }
            } else if (epi == EPI_BIAS_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (p.gelu_saves_grad && !p.exp_skip_gelu) {
                        float4 dg;
                        acc[i].x = gelu_erf_both(acc[i].x, dg.x); acc[i].y = gelu_erf_both(acc[i].y, dg.y);
                        acc[i].z = gelu_erf_both(acc[i].z, dg.z); acc[i].w = gelu_erf_both(acc[i].w, dg.w);
                        if (SCAT_ROW_OK(i)) *reinterpret_cast<float4*>(zptr + i * z_step) = dg;
                        continue;
                    }
                    if (SCAT_ROW_OK(i)) *reinterpret_cast<float4*>(zptr + i * z_step) = acc[i];
                    if (p.exp_skip_gelu) continue;
                    acc[i].x = gelu_erf(acc[i].x); acc[i].y = gelu_erf(acc[i].y);
                    acc[i].z = gelu_erf(acc[i].z); acc[i].w = gelu_erf(acc[i].w);
                }
            }
            if (p.round_out) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i].x = round_tf32(acc[i].x); acc[i].y = round_tf32(acc[i].y);
                    acc[i].z = round_tf32(acc[i].z); acc[i].w = round_tf32(acc[i].w);
                }
            }
            // ---- store phase ----
            if (hptr != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(acc[i].x, acc[i].y), hi = __floats2bfloat162_rn(acc[i].z, acc[i].w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                    if (SCAT_ROW_OK(i)) *reinterpret_cast<uint2*>(hptr + i * h_step) = pk;
                }
            }
            if (cptr != nullptr) {
                if (p.atomic_out) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (SCAT_ROW_OK(i)) red_add_v4(cptr + i * c_step, acc[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (SCAT_ROW_OK(i)) *reinterpret_cast<float4*>(cptr + i * c_step) = acc[i];
                }
            }
        } else if (n < p.N) {
            // ragged / unaligned edge: scalar, rolled (rare: N tails and odd leading dimensions)
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int row = row0 + 4 * i;
                if (row >= p.M) break;
                const float4 o = *reinterpret_cast<const float4*>(scratch + (4 * i + rsub) * 32 + 4 * (cc ^ ((4 * i + rsub) & 7)));
                const float ov[4] = {o.x, o.y, o.z, o.w};
                const long long ar = p.aux_row_mod > 0 ? row % p.aux_row_mod : row;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ne = n + e;
                    if (ne >= p.N) break;
                    float t = ov[e];
                    if (epi == EPI_PE_MASK) {
                        bool mk = false;
                        for (int k = 0; k < p.n_masked; ++k) mk |= (p.mask_idx[k] == row);
                        const float tok = mk ? p.bias[ne] : (p.aux_in ? t + p.aux_in[ar * p.ld_aux_in + ne] : t);
                        if (p.aux_out) p.aux_out[zb * p.aux_out_z + (long long)row * p.ld_aux_out + ne] = tok;
                        else t = tok;
                    }
                    switch (epi) {
                        case EPI_BIAS: t += p.bias[ne]; break;
                        case EPI_BIAS_RESID: t += p.bias[ne] + p.aux_in[ar * p.ld_aux_in + ne]; break;
                        case EPI_BIAS_GELU: {
                            t += p.bias[ne];
                            float dg = t;
                            if (p.gelu_saves_grad) t = gelu_erf_both(t, dg);
                            else t = gelu_erf(t);
                            p.aux_out[(long long)row * p.ld_aux_out + ne] = dg;
                        } break;
                        case EPI_DGELU: {
                            const float z = p.aux_in[ar * p.ld_aux_in + ne];
                            t *= p.gelu_saves_grad ? z : gelu_erf_grad(z);
                        } break;
                        case EPI_RESID: t += p.aux_in[ar * p.ld_aux_in + ne]; break;
                        default: break;
                    }
                    if (p.round_out) t = round_tf32(t);
                    if (p.C16) p.C16[(long long)row * p.ldc16 + ne] = __float2bfloat16_rn(t);
                    if (p.C) {
                        float* cp = p.C + zb * p.c_z + (long long)row * p.ldc + ne;
                        if (p.atomic_out) atomicAdd(cp, t);
                        else *cp = (p.accumulate && !has_aux) ? *cp + t : t;
                    }
                }
            }
        }
        if (cptr != nullptr) cptr += 32;
        if (hptr != nullptr) hptr += 32;
        if (zptr != nullptr) zptr += 32;
        if (aptr != nullptr) aptr += 32;
    }
#undef SCAT_ROW_OK
}

__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t tmem_lane_base, float* scratch, int row_base,
                                              int n0, int lane, int BN, int zb, bool mark) {
    // (Compile-time specialisation of this function per (epilogue, output mode) was tried and measured: as inlined
    // code it spills at the 168-register cap of two CTAs per SM, as out-of-line functions the by-reference parameter
    // block costs every thread a 688-byte local copy, +1.2 us per launch.  The runtime-flag version stays.)
    if (row_base >= p.M) return;     // this warp's 32 rows are all outside M (the conv front end uses 21 of 128 rows)
    epilogue_tile_impl(p, tmem_lane_base, scratch, row_base, n0, lane, BN, zb, mark);
}

template <bool BF16, int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    using L = SmemLayout<BN>;
    using E = Elem<BF16>;
    const int STAGES = p.stages;
    constexpr int BK = E::BK;
    // no static shared memory in this kernel: the dynamic window starts at the CTA's shared base, which is 1024-byte
    // aligned (128B swizzle atoms need it); checked below rather than paid for with a kilobyte of slack
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const int bar_off = STAGES * L::STAGE_BYTES;
    const uint32_t full_bar = smem_base + bar_off;             // STAGES x 8 B each
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t conv_bar = empty_bar + 8 * STAGES;
    const uint32_t tfull_bar = conv_bar + 8 * STAGES;          // 2: accumulator buffer complete (MMA -> epilogue)
    const uint32_t tempty_bar = tfull_bar + 16;                // 2: accumulator buffer drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + bar_off + 8 * (3 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) dbg_mark(p, 0);
    if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();
    const int kb_all = (p.K + BK - 1) / BK;
    // tile space, n fastest: [z][m tile][n tile]; a CTA walks it with stride gridDim.x (one tile each unless persistent)
    // One tile per CTA (the usual case): a 3-D grid, the tile is the block index and nothing is divided.  Persistent:
    // a 1-D grid of resident CTAs strides through the linear tile index.
    const int tiles_n = p.tiles_n, tiles_mn = p.tiles_n * p.tiles_m;
    const bool persistent = p.persistent != 0;
    const int total_tiles = persistent ? tiles_mn * p.tiles_z : 1;
    const int tile_first = persistent ? (int)blockIdx.x : 0, tile_stride = persistent ? (int)gridDim.x : 1;
    auto decode_tile = [&](int tile, int& z, int& m0, int& n0) {
        if (persistent) {
            z = tile / tiles_mn;
            const int mn = tile - z * tiles_mn, mt = mn / tiles_n;
            m0 = mt * BM; n0 = (mn - mt * tiles_n) * BN;
        } else {
            z = blockIdx.z; m0 = blockIdx.y * BM; n0 = blockIdx.x * BN;
        }
    };
    // (allocations are powers of two >= 32 columns; the host never makes a wide-tile launch persistent)
    const uint32_t tmem_cols = BN > 128 ? 256u : (persistent ? 2u * BN : (uint32_t)BN);

    if (warp == 0 && lane == 0) {
        // fetch the two TMA descriptors while the barriers are set up (they are kernel parameters, not data of
        // the preceding kernel): the first bulk load otherwise pays the descriptor miss on top of the L2 latency
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
            mbar_init(conv_bar + 8 * s, TC_THREADS - 64);     // every epilogue/converter thread arrives
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar + 8 * i, 1);
            mbar_init(tempty_bar + 8 * i, 4);                 // one elected lane of each epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // B operands the caller declares independent of the preceding kernel (weights): the first ring-full of their stages
    // is requested BEFORE the dependency wait, so the L2 / HBM latency of the weight tile hides under that kernel's tail
    int b_prefetched = 0;
    if (warp == 0 && p.b_static && tile_first < total_tiles) {
        int z, m0, n0;
        decode_tile(tile_first, z, m0, n0);
        const int zb = p.batched ? z : 0;
        const int kb_begin = p.batched ? 0 : z * p.kb_per_split;
        const int num_kb = min(kb_all, kb_begin + p.kb_per_split) - kb_begin;
        const int bn0 = n0 + zb * p.b_row_z, bk_off = zb * p.b_k_z;
        b_prefetched = min(STAGES, num_kb);
        if (elect_one()) {
            for (int kb = 0; kb < b_prefetched; ++kb) {
                const uint32_t fb = full_bar + 8 * kb;
                const uint32_t sb = smem_base + kb * L::STAGE_BYTES + L::A_BYTES;
                const int k0 = (kb_begin + kb) * BK;
                mbar_expect_tx(fb, L::STAGE_BYTES);           // A's bytes arrive after the wait below
                if (!B_MN) {
                    tma_load_2d(&tmB, fb, sb, k0 + bk_off, bn0);
                } else {
#pragma unroll
                    for (int i = 0; i < BN / E::MN_BOX; ++i)
                        tma_load_2d(&tmB, fb, sb + i * E::MN_LBO, bn0 + E::MN_BOX * i, k0 + bk_off);
                }
            }
        }
        __syncwarp();
    }
    // everything above (barrier init, TMEM allocation) overlapped the previous kernel's tail; operands and the
    // output buffer may only be touched once that kernel has completed
    pdl_sync();
    if (threadIdx.x == 0) dbg_mark(p, 1);

    if (warp == 0) {
        // ===== TMA producer: runs ahead of the MMA warp across tile boundaries =====
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_stride) {
            int z, m0, n0;
            decode_tile(tile, z, m0, n0);
            const int zb = p.batched ? z : 0;
            const int kb_begin = p.batched ? 0 : z * p.kb_per_split;
            const int num_kb = min(kb_all, kb_begin + p.kb_per_split) - kb_begin;      // host guarantees >= 1
            int k0 = kb_begin * BK;
            const int am0 = m0 + zb * p.a_row_z, ak_off = zb * p.a_k_z, bn0 = n0 + zb * p.b_row_z, bk_off = zb * p.b_k_z;
            for (int kb = 0; kb < num_kb; ++kb) {
                const bool b_done = tile == tile_first && kb < b_prefetched;      // (first pass of the ring: slots are free)
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                if (elect_one()) {
                    const uint32_t fb = full_bar + 8 * s;
                    if (!b_done) mbar_expect_tx(fb, L::STAGE_BYTES);
                    const uint32_t sa = smem_base + s * L::STAGE_BYTES;
                    const uint32_t sb = sa + L::A_BYTES;
                    if (!A_MN) {
                        tma_load_2d(&tmA, fb, sa, k0 + ak_off, am0);                              // box {BK k, 128 rows}
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / E::MN_BOX; ++i)
                            tma_load_2d(&tmA, fb, sa + i * E::MN_LBO, am0 + E::MN_BOX * i, k0 + ak_off);    // box {MN_BOX rows, BK k}
                    }
                    if (b_done) {
                    } else if (!B_MN) {
                        tma_load_2d(&tmB, fb, sb, k0 + bk_off, bn0);                              // box {BK k, BN rows}
                    } else {
#pragma unroll
                        for (int i = 0; i < BN / E::MN_BOX; ++i)
                            tma_load_2d(&tmB, fb, sb + i * E::MN_LBO, bn0 + E::MN_BOX * i, k0 + bk_off);
                    }
                }
                __syncwarp();
                k0 += BK;
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format at [7,10)/[10,13), a/b major bits 15/16,
        // n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
        constexpr uint32_t idesc = (1u << 4) | (E::FMT << 7) | (E::FMT << 10) | ((A_MN ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        constexpr uint32_t a_lbo = A_MN ? E::MN_LBO : 16u, a_sbo = A_MN ? E::MN_SBO : 1024u, a_lay = A_MN ? E::MN_LAYOUT : 2u;
        constexpr uint32_t b_lbo = B_MN ? E::MN_LBO : 16u, b_sbo = B_MN ? E::MN_SBO : 1024u, b_lay = B_MN ? E::MN_LAYOUT : 2u;
        constexpr uint32_t a_hi = (a_sbo >> 4) | (1u << 14) | (a_lay << 29);
        constexpr uint32_t b_hi = (b_sbo >> 4) | (1u << 14) | (b_lay << 29);
        constexpr uint32_t a_kstep = (A_MN ? E::MN_KSTEP : 32u) >> 4, b_kstep = (B_MN ? E::MN_KSTEP : 32u) >> 4;
        const uint32_t a_lo0 = (smem_base >> 4) | ((a_lbo >> 4) << 16);
        const uint32_t b_lo0 = ((smem_base + L::A_BYTES) >> 4) | ((b_lbo >> 4) << 16);
        const uint32_t ready_bar = p.round_operands ? conv_bar : full_bar;
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_stride, ++it) {
            int z, m0_unused, n0_unused;
            decode_tile(tile, z, m0_unused, n0_unused);
            const int kb_begin = p.batched ? 0 : z * p.kb_per_split;
            const int num_kb = min(kb_all, kb_begin + p.kb_per_split) - kb_begin;
            const int as = it & 1;                                   // accumulator buffer
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(tempty_bar + 8 * as, aph ^ 1);                 // the epilogue has drained this buffer (2 tiles ago)
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(ready_bar + 8 * s, ph);
                tc_fence_after();
                if (it == 0 && kb == 0 && lane == 0) dbg_mark(p, 2);
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + s * (L::STAGE_BYTES >> 4);
                    const uint32_t b_lo = b_lo0 + s * (L::STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < BK / E::UK; ++k)
                        umma<BF16>(tmem_d, a_lo + k * a_kstep, a_hi, b_lo + k * b_kstep, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar + 8 * s);                  // smem slot is free once these MMAs have read it
                    if (kb == num_kb - 1) umma_commit(tfull_bar + 8 * as);   // accumulator complete
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===== operand conditioning during the main loop (fp32 operands only), then the epilogue =====
        // The tensor core TRUNCATES fp32 operands to TF32 (measured: -6.6e-4 mean bias on positive data, 2.6x the
        // error of round-to-nearest).  These four otherwise idle warps round every landed stage to TF32-nearest
        // in place (cvt.rna.tf32.f32) and hand it to the MMA warp through conv_bar.
        const int q = warp & 3;                      // a warp may only touch TMEM lanes [32*(warp%4), +32)
        // transposition scratch of the epilogue: the idle operand ring when this CTA owns a single tile, a dedicated
        // region behind the barriers when it is persistent (the ring is then busy with the next tile)
        float* scratch = reinterpret_cast<float*>(smem + (persistent ? bar_off + L::BAR_BYTES : 0)) + q * 32 * 32;
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_stride, ++it) {
            int z, m0, n0;
            decode_tile(tile, z, m0, n0);
            const int zb = p.batched ? z : 0;
            if (!BF16 && p.round_operands) {
                const int kb_begin = p.batched ? 0 : z * p.kb_per_split;
                const int num_kb = min(kb_all, kb_begin + p.kb_per_split) - kb_begin;
                const int ctid = threadIdx.x - 64;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar + 8 * s, ph);
                    uint4* st = reinterpret_cast<uint4*>(smem + s * L::STAGE_BYTES);
#pragma unroll 4
                    for (int i = ctid; i < L::STAGE_BYTES / 16; i += TC_THREADS - 64) {
                        uint4 v = st[i];
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.x) : "f"(__uint_as_float(v.x)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.y) : "f"(__uint_as_float(v.y)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.z) : "f"(__uint_as_float(v.z)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.w) : "f"(__uint_as_float(v.w)));
                        st[i] = v;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (UMMA) reads
                    mbar_arrive(conv_bar + 8 * s);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
            // ===== epilogue: TMEM -> registers -> shared (transpose) -> registers -> global =====
            const int as = it & 1;
            mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);          // every MMA of this tile has completed
            tc_fence_after();
            if (it == 0 && threadIdx.x == 64) dbg_mark(p, 3);
            epilogue_tile(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN), scratch, m0 + q * 32, n0, lane, BN, zb,
                          it == 0 && threadIdx.x == 64);
            tc_fence_before();                                       // TMEM reads ordered before the hand-back
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
        }
    }
    if (threadIdx.x == 64) dbg_mark(p, 7);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

long long* g_gemm_dbg = nullptr;
// A/B switch for the measurement in DESIGN.md: SCAT_GEMM_SHALLOW=1 keeps the 96 KB ring on every launch
const bool g_no_b_prefetch = [] { const char* e = getenv("SCAT_GEMM_NO_B_PREFETCH"); return e != nullptr && e[0] == '1'; }();
const bool g_exp_skip_gelu = [] { const char* e = getenv("SCAT_EXP_SKIP_GELU"); return e != nullptr && e[0] == '1'; }();
const bool g_no_wide_tiles = [] { const char* e = getenv("SCAT_GEMM_NO_WIDE"); return e != nullptr && e[0] == '1'; }();
const bool g_shallow_ring = [] { const char* e = getenv("SCAT_GEMM_SHALLOW"); return e != nullptr && e[0] == '1'; }();

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
// unit stride in one direction, 16-byte aligned base and leading stride
bool operand_ok(const void* p, long long s_row, long long s_k, int elem_bytes) {
    const long long q = 16 / elem_bytes;
    if (reinterpret_cast<uintptr_t>(p) & 15) return false;
    if (s_k == 1) return (s_row % q == 0) && s_row >= 1;                  // K-major
    if (s_row == 1) return (s_k % q == 0) && s_k >= 1;                    // MN-major
    return false;
}

template <bool BF16, int BN, bool A_MN, bool B_MN>
int launch_variant(const GemmArgs& g, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    using E = Elem<BF16>;
    auto kern = gemm_tc_kernel<BF16, BN, A_MN, B_MN>;
    SCAT_ENSURE_SMEM(kern, L::total(L::STAGES_DEEP, false));
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle mn_sw = BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    // extents of the tensor maps: one problem, or (batched) the whole stack of problems along rows / k
    const int nb1 = g.batch > 1 ? g.batch - 1 : 0;
    const long long a_rows = g.M + (long long)nb1 * g.a_row_z, a_k = g.K + (long long)nb1 * g.a_k_z;
    const long long b_rows = g.N + (long long)nb1 * g.b_row_z, b_k = g.K + (long long)nb1 * g.b_k_z;
    if (!A_MN) SCAT_PROPAGATE(make_map(&tmA, g.A, E::BYTES, a_k, a_rows, g.sam, E::BK, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    else SCAT_PROPAGATE(make_map(&tmA, g.A, E::BYTES, a_rows, a_k, g.sak, E::MN_BOX, E::BK, mn_sw));
    if (!B_MN) SCAT_PROPAGATE(make_map(&tmB, g.B, E::BYTES, b_k, b_rows, g.sbn, E::BK, BN, CU_TENSOR_MAP_SWIZZLE_128B));
    else SCAT_PROPAGATE(make_map(&tmB, g.B, E::BYTES, b_rows, b_k, g.sbk, E::MN_BOX, E::BK, mn_sw));
    TcParams p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.C = g.C; p.ldc = g.ldc;
    p.C16 = reinterpret_cast<__nv_bfloat16*>(g.C16); p.ldc16 = g.ldc16;
    p.epilogue = g.epilogue; p.bias = g.bias; p.aux_in = g.aux_in; p.ld_aux_in = g.ld_aux_in; p.aux_row_mod = g.aux_row_mod; p.aux_out = g.aux_out;
    p.ld_aux_out = g.ld_aux_out; p.accumulate = g.accumulate;
    p.round_operands = (BF16 || g.prerounded) ? 0 : 1;
    p.round_out = g.round_out;
    p.b_static = (g.b_static && (BF16 || g.prerounded) && !g_no_b_prefetch) ? 1 : 0;     // (the in-kernel rounding pass owns un-rounded stages)
    p.batched = g.batch > 1 ? 1 : 0;
    p.a_row_z = g.a_row_z; p.a_k_z = g.a_k_z; p.b_row_z = g.b_row_z; p.b_k_z = g.b_k_z; p.c_z = g.c_z; p.aux_out_z = g.aux_out_z;
    p.mask_idx = g.mask_idx; p.n_masked = g.n_masked;
    p.dbg = g_gemm_dbg;
    p.exp_skip_gelu = g_exp_skip_gelu ? 1 : 0;
    p.gelu_saves_grad = g.gelu_saves_grad;
    auto al16 = [](const void* q, long long ld) { return q == nullptr || (((uintptr_t)q & 15) == 0 && (ld & 3) == 0); };
    p.vec_ok = al16(g.C, g.ldc) && (g.C16 == nullptr || (((uintptr_t)g.C16 & 7) == 0 && (g.ldc16 & 3) == 0)) &&
               al16(g.aux_in, g.ld_aux_in) && al16(g.aux_out, g.ld_aux_out) && al16(g.bias, 0);
    SCAT_REQUIRE(!(g.accumulate && (g.epilogue == EPI_BIAS_RESID || g.epilogue == EPI_RESID || g.epilogue == EPI_DGELU)),
                 kErrUnsupported, "gemm_tc: accumulate cannot be combined with an epilogue that reads aux_in");
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), 1);
    // split-K (weight gradients: few output tiles, K = B*21 rows): slice the k-blocks over gridDim.z so that the
    // launch covers the 148 SMs twice; slices combine with vector reductions in L2 (C is pre-zeroed by the caller)
    const int kb_all = ceil_div(g.K, E::BK);
    int splits = 1;
    const int tiles = (int)(grid.x * grid.y);
    if (g.allow_split_k && g.epilogue == EPI_NONE && !g.round_out && !g.C16 && tiles <= 148 && kb_all >= 8)
        splits = max(1, min(kb_all / 4, 296 / tiles));
    if (g.batch > 1) splits = 1;
    p.kb_per_split = ceil_div(kb_all, splits);
    grid.z = g.batch > 1 ? g.batch : ceil_div(kb_all, p.kb_per_split);
    p.atomic_out = (g.batch > 1 ? g.batch_accumulate : grid.z > 1) ? 1 : 0;
    // 1-D grid over the tile space; more tiles than resident CTA slots (2 per SM): persistent CTAs with
    // double-buffered accumulators stride through it, so that loads and epilogues of successive tiles overlap
    p.tiles_n = (int)grid.x; p.tiles_m = (int)grid.y; p.tiles_z = (int)grid.z;
    const int total_tiles = p.tiles_n * p.tiles_m * p.tiles_z;
    constexpr int kSlots = 2 * 148;
    p.persistent = total_tiles > kSlots ? 1 : 0;
    if (p.persistent) grid = dim3(kSlots, 1, 1);
    p.stages = (BN > 128 || (total_tiles <= 148 && !g_shallow_ring)) ? L::STAGES_DEEP : L::STAGES;
    SCAT_REQUIRE(BN <= 128 || (total_tiles <= 148 && !p.persistent), kErrUnsupported, "gemm_tc: wide tiles need a single-wave launch");
    if (p.atomic_out && !g.accumulate && !g.c_zeroed)
        SCAT_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, stream));
    SCAT_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(TC_THREADS), L::total(p.stages, p.persistent != 0), stream, tmA, tmB, p));
    SCAT_CHECK_LAUNCH();
    return 0;
}

template <bool BF16, int BN>
int launch_major(const GemmArgs& g, cudaStream_t stream) {
    const bool a_mn = g.sak != 1, b_mn = g.sbk != 1;
    if (!a_mn && !b_mn) return launch_variant<BF16, BN, false, false>(g, stream);
    if (!a_mn && b_mn) return launch_variant<BF16, BN, false, true>(g, stream);
    if (a_mn && !b_mn) return launch_variant<BF16, BN, true, false>(g, stream);
    return launch_variant<BF16, BN, true, true>(g, stream);
}

}  // namespace

void gemm_tc_set_debug_buffer(long long* dev8) { g_gemm_dbg = dev8; }

bool gemm_tc_supported(const GemmArgs& g) {
    const int eb = g.operand_bf16 ? 2 : 4;
    if (g.M < 1 || g.N < 8 || g.K < 8) return false;
    if (!operand_ok(g.A, g.sam, g.sak, eb) || !operand_ok(g.B, g.sbn, g.sbk, eb)) return false;
    return get_encode_fn() != nullptr;
}

int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream) {
    SCAT_REQUIRE(precision == PREC_TF32 || precision == PREC_BF16, kErrBadArg, "gemm_tc: precision %d", precision);
    const int eb = g.operand_bf16 ? 2 : 4;
    SCAT_REQUIRE(operand_ok(g.A, g.sam, g.sak, eb) && operand_ok(g.B, g.sbn, g.sbk, eb), kErrUnsupported,
                 "gemm_tc: operands need a unit stride, 16-byte aligned base and 16-byte multiple leading stride");
    SCAT_REQUIRE(g.C != nullptr || g.C16 != nullptr, kErrBadArg, "gemm_tc: no output");
    // The main loop is bound by the shared-memory fill (one SM takes ~64 B per cycle from L2, measured with
    // tools/gemm_timeline.py: 500 cycles per 32 KB k-block whether one deep-ring CTA or two shallow ones share the SM), so
    // pick the tile width that minimises what the busiest SM has to stage, plus the (serial) epilogue of its last tile:
    //   cycles(BN) = waves x k-blocks x (BM + BN) x 128 B / 64 B + (BN / 32) x 450,   waves = ceil(tiles / 148).
    // 192- and 256-wide tiles only when the launch then fits one wave (one CTA per SM, deep ring) and needs no split-K:
    // M = 2016 x N = 1536 (every qkv projection) is 192 tiles of 128 = 1.3 waves but 128 tiles of 192; the stacked
    // 4032-row dgrads with N = 784 / 588 are 224 / 160 tiles of 128 but 128 tiles of 256 / 192.
    const int tm = ceil_div(g.M, BM);
    const int kb = ceil_div(g.K, g.operand_bf16 ? 64 : 32);
    auto cycles = [&](int bn) {
        const long long tiles = (long long)tm * ceil_div(g.N, bn);
        return ceil_div((int)tiles, 148) * (long long)kb * (BM + bn) * 2 + (bn / 32) * 450LL;
    };
    int bn = 64;
    if (g.force_bn) {
        bn = g.force_bn;
    } else if (g.N > 64) {
        bn = cycles(128) <= cycles(64) ? 128 : 64;
        const bool wide_ok = g.allow_wide && !g_no_wide_tiles && g.batch <= 1 && !g.allow_split_k;
        for (int cand : {192, 256})
            if (wide_ok && g.N > cand - 64 && (long long)tm * ceil_div(g.N, cand) <= 148 && cycles(cand) < cycles(bn)) bn = cand;
    }
    SCAT_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, kErrBadArg, "gemm_tc: tile width %d", bn);
#define SCAT_TC_DISPATCH(BF)                                               \
    switch (bn) {                                                          \
        case 64: return launch_major<BF, 64>(g, stream);                   \
        case 128: return launch_major<BF, 128>(g, stream);                 \
        case 192: return launch_major<BF, 192>(g, stream);                 \
        default: return launch_major<BF, 256>(g, stream);                  \
    }
    if (g.operand_bf16) { SCAT_TC_DISPATCH(true) }
    SCAT_TC_DISPATCH(false)
#undef SCAT_TC_DISPATCH
}

}  // namespace scat
