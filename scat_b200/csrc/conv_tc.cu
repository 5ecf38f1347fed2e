// The 1x1-conv front end of the head (hand_net.py:329,363-373) and its backward as three persistent tcgen05 kernels.
//
// All three passes are thin contractions against a 21-token weight (one dimension is 21, the other 512 channels or
// 784 pixels), so they are HBM streams of the backbone seam tensor x2[B,512,28,28] (154 MB fp32 at B = 96) or of
// x2.grad, with a tensor-core contraction riding along.  Each kernel is one CTA per SM:
//
//   warp 0      TMA producer over a deep shared-memory ring (8 x 16 KB of x2 in flight per SM)
//   warp 1      tcgen05.mma issuer, accumulators in TMEM (double-buffered where a tile has an epilogue)
//   warps 2-5   epilogue: tcgen05.ld -> registers -> coalesced global stores / L2 reductions / TMA stores
//   warps 6-9   fp32 seam only: round every landed x2 stage to TF32-NEAREST in shared memory (cvt.rna) before the
//               tensor core reads it -- tcgen05 kind::tf32 would otherwise TRUNCATE the caller's fp32 data
//
// The big tensor is always the A operand (M = 128 pixels or channels on the TMEM lanes), the 21-token operand is B:
//
//   forward   Fv[b][t, px]   = sum_c  W[t, c]  x2[b][c, px]      A = x2 tile, MN-major (px contiguous), B = W resident
//   wgrad     dW[t, c]      += sum_px dFv[b][t, px] x2[b][c, px]  A = x2 tile, K-major  (px = k),         B = dFv tile
//   dgrad     dx2[b][c, px]  = sum_t  W[t, c]  dFv[b][t, px]      A = dFv tile, MN-major, B = W^T resident, D -> TMA store
//
// Seam dtype (ScatHeadDesc.x2_dtype):
//   fp32  kind::tf32; x2 rounded to TF32-nearest in shared memory, W rounded by conv_weight_prep; the backward uses the
//         exact TF32 hi/lo split of the small d-token tensor ([hi; lo; hi] against [Wh; Wh; Wl]): fp32-grade x2.grad
//   bf16  kind::f16; x2 is exact on the tensor core, W and the d tokens enter as 2-3 bf16 split terms stacked along N
//         or K, so products carry >= 16 mantissa bits and only the bf16 I/O rounds; x2.grad is written as bf16
#include "kernels.h"
#include "tc_ptx.cuh"

namespace scat {
namespace {

constexpr int kT = 21;                 // tokens = joints (hand_net.py:328-329); the split stacks assume 3 * 21 <= 64
constexpr int CT_THREADS = 320;
constexpr int A_STAGE = 16384;         // bytes of one x2 stage: 128 x 32 fp32 or 128 x 64 bf16
constexpr int CT_STAGES = 8;

template <bool BF16>
struct CElem {
    static constexpr int BYTES = BF16 ? 2 : 4;
    static constexpr int BK = 128 / BYTES;                     // k per stage = one 128-byte swizzle row: 32 fp32 / 64 bf16
    static constexpr int UK = 32 / BYTES;                      // k per tcgen05.mma
    static constexpr int MN_BOX = 128 / BYTES;                 // mn elements per row of an MN-major box
    static constexpr int NROWS = BF16 ? 64 : 32;               // rows of the token operand (3 x 21 split terms / 21)
    static constexpr uint32_t FMT = BF16 ? 1u : 2u;            // instruction descriptor operand format
    static constexpr uint32_t MN_LAYOUT = BF16 ? 2u : 1u;      // SWIZZLE_128B / SWIZZLE_128B_BASE32B (see gemm_tc.cu)
    static constexpr uint32_t MN_SBO = BF16 ? 1024u : 512u;
    static constexpr uint32_t MN_KSTEP = BF16 ? 2048u : 1024u; // bytes between UMMA k-slices inside a box
    static constexpr uint32_t BOX_BYTES = BK * 128;            // one MN-major box (BK k rows x 128 B) = LBO between mn groups
    static constexpr int TOK_BYTES = NROWS * 128;              // one K-major k-block of the token operand
};

__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return (addr >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo, uint32_t layout) { return (sbo >> 4) | (1u << 14) | (layout << 29); }
template <bool BF16>
__device__ __forceinline__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn, int n) {
    return (1u << 4) | (CElem<BF16>::FMT << 7) | (CElem<BF16>::FMT << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// in-place TF32-nearest rounding of `bytes` of shared memory by the 128 converter threads
__device__ __forceinline__ void round_stage(uint8_t* stage, int bytes, int ctid) {
    uint4* st = reinterpret_cast<uint4*>(stage);
#pragma unroll 4
    for (int i = ctid; i < bytes / 16; i += 128) {
        uint4 v = st[i];
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.x) : "f"(__uint_as_float(v.x)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.y) : "f"(__uint_as_float(v.y)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.z) : "f"(__uint_as_float(v.z)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.w) : "f"(__uint_as_float(v.w)));
        st[i] = v;
    }
    fence_async_smem();
}

// =================================================================================================================
// forward: feat_visual = W . x2 (hand_net.py:363), token matrix = feat_visual + pe (:367) with masked tokens replaced
// by mask_token (:369-373).  Work unit = TPU consecutive 128-pixel tiles of one sample (all 512 channels).
// =================================================================================================================
struct ConvFwdParams {
    int B, C, HW;
    int units_full;              // full units per sample
    int tail_px0, tail_tiles;    // the ragged rest of a sample's pixels (784 = 3 x 256 + 16): its own, short unit
    int n_units;
    const float* pe;             // [T, HW] or null
    const float* mask_token;     // [HW]
    const int32_t* mask_idx;
    int n_masked;
    float* fv;                   // [B, T, HW]
    float* X0;                   // [B, T, HW]; == fv: the token matrix aliases feat_visual (pos_embed off)
};

template <int TPU>
__device__ __forceinline__ void fwd_decode(const ConvFwdParams& p, int v, int& b, int& px0, int& nt) {
    const int nfull = p.B * p.units_full;            // full units first, the short tail units fill the end of the schedule
    if (v < nfull) {
        b = v / p.units_full;
        px0 = (v - b * p.units_full) * (128 * TPU);
        nt = TPU;
    } else {
        b = v - nfull;
        px0 = p.tail_px0;
        nt = p.tail_tiles;
    }
}

template <bool BF16, int TPU>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvFwdParams p) {
    using E = CElem<BF16>;
    constexpr int NR = E::NROWS;
    constexpr int W_BYTES = 65536;                    // resident weight: (C / BK) k-blocks x NROWS rows x 128 B (C <= 512)
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase, ring = sbase + W_BYTES;
    const uint32_t full_bar = ring + CT_STAGES * A_STAGE, empty_bar = full_bar + 8 * CT_STAGES, conv_bar = empty_bar + 8 * CT_STAGES;
    const uint32_t w_bar = conv_bar + 8 * CT_STAGES, tfull_bar = w_bar + 8, tempty_bar = tfull_bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + W_BYTES + CT_STAGES * A_STAGE + 8 * (3 * CT_STAGES + 5));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = p.C / E::BK;
    if (threadIdx.x == 0) {
        if ((sbase & 1023u) != 0) __trap();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        for (int s = 0; s < CT_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
            mbar_init(conv_bar + 8 * s, 128);
        }
        mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar + 8 * i, 1);
            mbar_init(tempty_bar + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = 2 * TPU * NR;      // 2 accumulator sets x TPU tiles x NR columns (64 .. 256)
    if (warp == 1) tmem_alloc_n(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        // ===== producer =====
        if (elect_one()) {
            mbar_expect_tx(w_bar, (uint32_t)(KB * E::TOK_BYTES));
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmW, w_bar, sW + kb * E::TOK_BYTES, kb * E::BK, 0);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        for (int v = blockIdx.x; v < p.n_units; v += gridDim.x) {
            int b, px0, nt;
            fwd_decode<TPU>(p, v, b, px0, nt);
            for (int kb = 0; kb < KB; ++kb) {
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(empty_bar + 8 * s, ph ^ 1);
                    if (elect_one()) {
                        const int px = px0 + t * 128;
                        // pixel boxes that start beyond the row are skipped: their accumulator rows are never stored
                        const int nbox = min(128 / E::MN_BOX, (p.HW - px + E::MN_BOX - 1) / E::MN_BOX);
                        const uint32_t fb = full_bar + 8 * s, dst = ring + s * A_STAGE;
                        mbar_expect_tx(fb, (uint32_t)nbox * E::BOX_BYTES);
                        for (int i = 0; i < nbox; ++i)
                            tma_load_2d(&tmX, fb, dst + i * E::BOX_BYTES, px + i * E::MN_BOX, b * p.C + kb * E::BK);
                    }
                    __syncwarp();
                    if (++s == CT_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 px, NR] += x2 tile [128 px, BK ch] . W[NR, BK ch]^T =====
        constexpr uint32_t idesc = make_idesc<BF16>(true, false, NR);
        constexpr uint32_t a_hi = desc_hi(E::MN_SBO, E::MN_LAYOUT), b_hi = desc_hi(1024u, 2u);
        mbar_wait(w_bar, 0);
        tc_fence_after();
        int s = 0, it = 0;
        uint32_t ph = 0;
        const uint32_t ready_bar = BF16 ? full_bar : conv_bar;
        for (int v = blockIdx.x; v < p.n_units; v += gridDim.x, ++it) {
            int b, px0, nt;
            fwd_decode<TPU>(p, v, b, px0, nt);
            const int as = it & 1;
            mbar_wait(tempty_bar + 8 * as, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb) {
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(ready_bar + 8 * s, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo = desc_lo(ring + s * A_STAGE, E::BOX_BYTES);
                        const uint32_t b_lo = desc_lo(sW + kb * E::TOK_BYTES, 16u);
                        const uint32_t tmem_d = tmem_base + (uint32_t)((as * TPU + t) * NR);
#pragma unroll
                        for (int k = 0; k < E::BK / E::UK; ++k)
                            umma<BF16>(tmem_d, a_lo + k * (E::MN_KSTEP >> 4), a_hi, b_lo + k * 2, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(empty_bar + 8 * s);
                        if (kb == KB - 1 && t == nt - 1) umma_commit(tfull_bar + 8 * as);
                    }
                    __syncwarp();
                    if (++s == CT_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp < 6) {
        // ===== epilogue: lane = pixel, registers = tokens; every store is one 128-byte row segment per warp =====
        const int q = warp & 3;
        uint32_t maskbits = 0;
        for (int k = 0; k < p.n_masked; ++k) maskbits |= 1u << __ldg(p.mask_idx + k);
        const bool alias = p.X0 == p.fv;
        int it = 0;
        for (int v = blockIdx.x; v < p.n_units; v += gridDim.x, ++it) {
            int b, px0, nt;
            fwd_decode<TPU>(p, v, b, px0, nt);
            const int as = it & 1;
            // everything the epilogue reads from memory is independent of the accumulator: fetch it while the MMAs of
            // this unit are still running (a load -> add -> store chain per token would otherwise expose 21 L2 round
            // trips per tile on the critical path of a CTA that only owns two or three units)
            float pe_r[TPU][kT], mtok[TPU];
#pragma unroll
            for (int t = 0; t < TPU; ++t) {
                const int px = px0 + t * 128 + q * 32 + lane;
                const bool ok = t < nt && px < p.HW;
                mtok[t] = (ok && maskbits) ? __ldg(p.mask_token + px) : 0.f;
#pragma unroll
                for (int j = 0; j < kT; ++j) pe_r[t][j] = (ok && p.pe) ? __ldg(p.pe + j * p.HW + px) : 0.f;
            }
            mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int t = 0; t < TPU; ++t) {
                if (t >= nt || px0 + t * 128 + q * 32 >= p.HW) break;   // warp-uniform: this warp's 32 pixels are beyond the row
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * TPU + t) * NR);
                float c[NR];
                tmem_ld32_nowait(taddr, c);
                if (BF16) tmem_ld32_nowait(taddr + 32, c + (BF16 ? 32 : 0));
                tmem_ld_wait();
                const int px = px0 + t * 128 + q * 32 + lane;
                if (px < p.HW) {
                    float* fvp = p.fv + (long long)b * kT * p.HW + px;
                    float* xp = p.X0 + (long long)b * kT * p.HW + px;
#pragma unroll
                    for (int j = 0; j < kT; ++j) {
                        // bf16 seam: columns j, 21 + j, 42 + j hold x2 . (W1, W2, W3), smallest terms first
                        const float val = BF16 ? (c[(2 * kT + j) % NR] + c[(kT + j) % NR]) + c[j] : c[j];
                        const bool mk = (maskbits >> j) & 1u;
                        const float tok = mk ? mtok[t] : val + pe_r[t][j];   // pe_r is 0 without positional encoding
                        if (alias) {
                            fvp[j * p.HW] = tok;                    // hand_net.py:364,373: the overwrite lands in feat_visual
                        } else {
                            fvp[j * p.HW] = val;
                            xp[j * p.HW] = tok;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
        }
    } else if (!BF16) {
        // ===== converters: TF32-nearest rounding of every landed x2 stage =====
        const int ctid = threadIdx.x - 192;
        int s = 0;
        uint32_t ph = 0;
        for (int v = blockIdx.x; v < p.n_units; v += gridDim.x) {
            int b, px0, nt;
            fwd_decode<TPU>(p, v, b, px0, nt);
            for (int i = 0; i < KB * nt; ++i) {
                mbar_wait(full_bar + 8 * s, ph);
                round_stage(smem + W_BYTES + s * A_STAGE, A_STAGE, ctid);
                mbar_arrive(conv_bar + 8 * s);
                if (++s == CT_STAGES) { s = 0; ph ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =================================================================================================================
// weight gradient: dW[t, c] += sum over samples and pixels of dFv[b][t, px] * x2[b][c, px].
// The (channel tile, sample, pixel block) space is cut into equal contiguous ranges, one per CTA; a CTA keeps its
// accumulator in TMEM across samples and reduces it into dW (L2 reductions) when its channel tile changes or its range
// ends -- at most twice.
// =================================================================================================================
struct ConvWgradParams {
    int B, C, HW;
    int kbw;               // pixel blocks per (sample, channel tile)
    int d_rows, d_off;     // rows per sample of the split d-token tensor, first row of the window this pass reads
    long long total;       // (C / 128) * B * kbw
    float* dW;             // [T, C], zero on entry
};

template <bool BF16>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD, const ConvWgradParams p) {
    using E = CElem<BF16>;
    constexpr int NR = E::NROWS;
    constexpr int STAGE = A_STAGE + E::TOK_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t full_bar = sbase + CT_STAGES * STAGE, empty_bar = full_bar + 8 * CT_STAGES, conv_bar = empty_bar + 8 * CT_STAGES;
    const uint32_t tfull_bar = conv_bar + 8 * CT_STAGES, tempty_bar = tfull_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + CT_STAGES * STAGE + 8 * (3 * CT_STAGES + 2));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        if ((sbase & 1023u) != 0) __trap();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmD)) : "memory");
        for (int s = 0; s < CT_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
            mbar_init(conv_bar + 8 * s, 128);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = NR;
    if (warp == 1) tmem_alloc_n(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    const long long lo = p.total * blockIdx.x / gridDim.x, hi = p.total * (blockIdx.x + 1) / gridDim.x;
    const long long per_ct = (long long)p.B * p.kbw;
    const int ct_lo = (int)(lo / per_ct);
    const int r_lo = (int)(lo - ct_lo * per_ct);
    const int b_lo = r_lo / p.kbw, kb_lo = r_lo - b_lo * p.kbw;

    if (warp == 0) {
        int s = 0, ct = ct_lo, b = b_lo, kb = kb_lo;
        uint32_t ph = 0;
        for (long long idx = lo; idx < hi; ++idx) {
            mbar_wait(empty_bar + 8 * s, ph ^ 1);
            if (elect_one()) {
                const uint32_t fb = full_bar + 8 * s, dst = sbase + s * STAGE;
                mbar_expect_tx(fb, STAGE);
                tma_load_2d(&tmX, fb, dst, kb * E::BK, b * p.C + ct * 128);                     // box {BK px, 128 channels}
                tma_load_2d(&tmD, fb, dst + A_STAGE, kb * E::BK, b * p.d_rows + p.d_off);      // box {BK px, NR token rows}
            }
            __syncwarp();
            if (++s == CT_STAGES) { s = 0; ph ^= 1; }
            if (++kb == p.kbw) { kb = 0; if (++b == p.B) { b = 0; ++ct; } }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<BF16>(false, false, NR);
        constexpr uint32_t k_hi = desc_hi(1024u, 2u);
        int s = 0, seg = 0;
        uint32_t ph = 0;
        const uint32_t ready_bar = BF16 ? full_bar : conv_bar;
        long long in_ct = (long long)r_lo;                         // position inside the current channel tile
        bool first = true;
        for (long long idx = lo; idx < hi; ++idx) {
            if (first) {
                mbar_wait(tempty_bar, (uint32_t)(seg & 1) ^ 1);    // the previous segment's accumulator has been drained
                tc_fence_after();
            }
            mbar_wait(ready_bar + 8 * s, ph);
            tc_fence_after();
            const bool last = (idx + 1 == hi) || (in_ct + 1 == per_ct);
            if (elect_one()) {
                const uint32_t a_lo = desc_lo(sbase + s * STAGE, 16u), b_lo = desc_lo(sbase + s * STAGE + A_STAGE, 16u);
#pragma unroll
                for (int k = 0; k < E::BK / E::UK; ++k)
                    umma<BF16>(tmem_base, a_lo + k * 2, k_hi, b_lo + k * 2, k_hi, idesc, (first && k == 0) ? 0u : 1u);
                umma_commit(empty_bar + 8 * s);
                if (last) umma_commit(tfull_bar);
            }
            __syncwarp();
            if (++s == CT_STAGES) { s = 0; ph ^= 1; }
            first = false;
            if (last) { first = true; ++seg; }
            if (++in_ct == per_ct) in_ct = 0;
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        int seg = 0, ct = ct_lo;
        long long pos = lo;
        while (pos < hi) {
            const long long seg_end = min(hi, (long long)(ct + 1) * per_ct);
            mbar_wait(tfull_bar, (uint32_t)(seg & 1));
            tc_fence_after();
            float c[NR];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
            tmem_ld32_nowait(taddr, c);
            if (BF16) tmem_ld32_nowait(taddr + 32, c + (BF16 ? 32 : 0));
            tmem_ld_wait();
            float* dst = p.dW + ct * 128 + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < kT; ++j) {
                const float val = BF16 ? (c[(2 * kT + j) % NR] + c[(kT + j) % NR]) + c[j] : c[j];
                atomicAdd(dst + j * p.C, val);                      // RED.ADD.F32, 128 contiguous bytes per warp
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            pos = seg_end;
            ++ct;
            ++seg;
        }
    } else if (!BF16) {
        const int ctid = threadIdx.x - 192;
        int s = 0;
        uint32_t ph = 0;
        for (long long idx = lo; idx < hi; ++idx) {
            mbar_wait(full_bar + 8 * s, ph);
            round_stage(smem + s * STAGE, A_STAGE, ctid);          // the d-token operand is already TF32-exact (hi part)
            mbar_arrive(conv_bar + 8 * s);
            if (++s == CT_STAGES) { s = 0; ph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =================================================================================================================
// data gradient: x2.grad[b][c, px] = sum_j Wstack[j, c] * dsplit[b][j, px], K = 64 stacked split terms.  A CTA owns
// one half of the channels (its 256 x 64 slice of W^T stays in shared memory) and walks 128-pixel tiles; the 128 x 256
// accumulator leaves through per-warp shared-memory boxes and TMA stores (whole 128-byte lines, rows clipped at HW).
// =================================================================================================================
struct ConvDgradParams {
    int B, C, HW;
    int tiles_full, tail_px0, has_tail;
    int n_units;           // per channel half
    int d_rows;            // rows per sample of the split d-token tensor (window starts at row 0)
    int n_half;            // C / 256
};
constexpr int DG_STAGES = 3;
constexpr int DG_N = 256;

__device__ __forceinline__ void dgrad_decode(const ConvDgradParams& p, int v, int& b, int& px0) {
    const int nfull = p.B * p.tiles_full;
    if (v < nfull) { b = v / p.tiles_full; px0 = (v - b * p.tiles_full) * 128; }
    else { b = v - nfull; px0 = p.tail_px0; }
}

constexpr int DG_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)

template <bool BF16>
__global__ void __launch_bounds__(DG_THREADS, 1)
conv_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmO, const ConvDgradParams p) {
    using E = CElem<BF16>;
    constexpr int KH = 64 / E::BK;                              // k-halves: 2 (fp32, 32 k per box) / 1 (bf16)
    constexpr int A_KH = (128 / E::MN_BOX) * E::BOX_BYTES;      // bytes of one k-half of the A tile (16 KB)
    constexpr int A_BYTES = KH * A_KH;                          // 32 KB fp32 / 16 KB bf16
    constexpr int W_KH = (DG_N / E::MN_BOX) * E::BOX_BYTES;     // 32 KB
    constexpr int W_BYTES = KH * W_KH;                          // 64 KB fp32 / 32 KB bf16
    constexpr int OUT_BYTES = BF16 ? 2 : 4;
    constexpr int STG = 32 * 32 * OUT_BYTES;                    // one staging box: 32 channel rows x 32 px
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase, ring = sbase + W_BYTES, stg = ring + DG_STAGES * A_BYTES;
    const uint32_t full_bar = stg + 8 * 2 * STG, empty_bar = full_bar + 8 * DG_STAGES, w_bar = empty_bar + 8 * DG_STAGES;
    const uint32_t tfull_bar = w_bar + 8, tempty_bar = tfull_bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + W_BYTES + DG_STAGES * A_BYTES + 8 * 2 * STG + 8 * (2 * DG_STAGES + 5));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = blockIdx.x % p.n_half, cta = blockIdx.x / p.n_half, n_cta = gridDim.x / p.n_half;
    if (threadIdx.x == 0) {
        if ((sbase & 1023u) != 0) __trap();
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmD)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
        for (int s = 0; s < DG_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        mbar_init(w_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar + 8 * i, 1);
            mbar_init(tempty_bar + 8 * i, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr uint32_t TMEM_COLS = 2 * DG_N;
    if (warp == 1) tmem_alloc_n(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(w_bar, W_BYTES);
            for (int kh = 0; kh < KH; ++kh)
                for (int i = 0; i < DG_N / E::MN_BOX; ++i)          // box {MN_BOX channels, BK stacked-token rows}
                    tma_load_2d(&tmW, w_bar, sW + kh * W_KH + i * E::BOX_BYTES, half * DG_N + i * E::MN_BOX, kh * E::BK);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        for (int v = cta; v < p.n_units; v += n_cta) {
            int b, px0;
            dgrad_decode(p, v, b, px0);
            mbar_wait(empty_bar + 8 * s, ph ^ 1);
            if (elect_one()) {
                const int nbox = min(128 / E::MN_BOX, (p.HW - px0 + E::MN_BOX - 1) / E::MN_BOX);
                const uint32_t fb = full_bar + 8 * s, dst = ring + s * A_BYTES;
                mbar_expect_tx(fb, (uint32_t)(KH * nbox) * E::BOX_BYTES);
                for (int kh = 0; kh < KH; ++kh)
                    for (int i = 0; i < nbox; ++i)                  // box {MN_BOX px, BK stacked-token rows}
                        tma_load_2d(&tmD, fb, dst + kh * A_KH + i * E::BOX_BYTES, px0 + i * E::MN_BOX, b * p.d_rows + kh * E::BK);
            }
            __syncwarp();
            if (++s == DG_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc<BF16>(true, true, DG_N);
        constexpr uint32_t mn_hi = desc_hi(E::MN_SBO, E::MN_LAYOUT);
        constexpr int KPH = E::BK / E::UK;                          // UMMA k-slices per k-half (4)
        mbar_wait(w_bar, 0);
        tc_fence_after();
        int s = 0, it = 0;
        uint32_t ph = 0;
        for (int v = cta; v < p.n_units; v += n_cta, ++it) {
            const int as = it & 1;
            mbar_wait(tempty_bar + 8 * as, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            mbar_wait(full_bar + 8 * s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * DG_N);
#pragma unroll
                for (int k = 0; k < 64 / E::UK; ++k) {
                    const int kh = k / KPH, kk = k % KPH;
                    const uint32_t a_lo = desc_lo(ring + s * A_BYTES + kh * A_KH + kk * E::MN_KSTEP, E::BOX_BYTES);
                    const uint32_t b_lo = desc_lo(sW + kh * W_KH + kk * E::MN_KSTEP, E::BOX_BYTES);
                    umma<BF16>(tmem_d, a_lo, mn_hi, b_lo, mn_hi, idesc, k != 0 ? 1u : 0u);
                }
                umma_commit(empty_bar + 8 * s);
                umma_commit(tfull_bar + 8 * as);
            }
            __syncwarp();
            if (++s == DG_STAGES) { s = 0; ph ^= 1; }
        }
    } else {
        // ===== epilogue: lane = pixel, 32-channel column chunks -> [32 ch][32 px] box in shared memory -> TMA store.
        // Two warps share each TMEM lane quadrant and take alternate chunks: the chain tcgen05.ld -> 32 st.shared ->
        // fence -> TMA store is latency bound per warp, and this kernel is one long epilogue =====
        const int q = warp & 3, part = (warp - 2) >> 2, ew = warp - 2;
        const uint32_t my_stg = stg + ew * 2 * STG;
        uint8_t* my_stg_ptr = smem + W_BYTES + DG_STAGES * A_BYTES + ew * 2 * STG;
        int it = 0, nstore = 0;
        for (int v = cta; v < p.n_units; v += n_cta, ++it) {
            int b, px0;
            dgrad_decode(p, v, b, px0);
            const int as = it & 1;
            mbar_wait(tfull_bar + 8 * as, (it >> 1) & 1);
            tc_fence_after();
            const int pxw = px0 + q * 32;
            if (pxw < p.HW) {
                for (int cc = part; cc < DG_N / 32; cc += 2, ++nstore) {
                    const int buf = nstore & 1;
                    if (lane == 0) bulk_wait_read<1>();             // the store that last read this buffer has drained it
                    __syncwarp();
                    float c[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * DG_N + cc * 32), c);
                    if (BF16) {
                        __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(my_stg_ptr + buf * STG) + lane;
#pragma unroll
                        for (int j = 0; j < 32; ++j) sp[j * 32] = __float2bfloat16_rn(c[j]);
                    } else {
                        float* sp = reinterpret_cast<float*>(my_stg_ptr + buf * STG) + lane;
#pragma unroll
                        for (int j = 0; j < 32; ++j) sp[j * 32] = c[j];
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmO, my_stg + buf * STG, pxw, b * p.C + half * DG_N + cc * 32);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
        }
        if (lane == 0) bulk_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =================================================================================================================
// operand preparation (both small): the conv weight stacks, once per forward, and the split d-token tensor (masking fused)
// =================================================================================================================
// fp32 seam: dst = fp32 [3T, C] = [Wh; Wh; Wl]  (Wh = TF32-nearest(W), Wl = TF32-nearest(W - Wh)); the forward reads
//            the first T rows, the data gradient all of them against [hi; lo; hi]
// bf16 seam: dst = bf16 [3T, C] = [W1; W2; W3] (forward: three-term split, 24 mantissa bits) followed by
//            bf16 [3T, C] = [W1; W2; W1] (data gradient, against [d1; d1; d2])
__global__ void conv_weight_prep_kernel(const float* __restrict__ W, void* __restrict__ dst, int n, int bf16) {
    pdl_sync();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float w = __ldg(W + i);
        if (!bf16) {
            float* d = reinterpret_cast<float*>(dst);
            const float h = round_tf32(w);
            d[i] = h; d[n + i] = h; d[2 * n + i] = round_tf32(w - h);
        } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst);
            const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
            const float r1 = w - __bfloat162float(w1);
            const __nv_bfloat16 w2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 w3 = __float2bfloat16_rn(r1 - __bfloat162float(w2));
            d[i] = w1; d[n + i] = w2; d[2 * n + i] = w3;
            d[3 * n + i] = w1; d[4 * n + i] = w2; d[5 * n + i] = w1;
        }
    }
}

// Operand preparation of the backward, one pass over the cotangent of the token matrix dX0 [B,T,HW] (fp32):
//   * rows of masked tokens do not depend on the conv output (hand_net.py:373): they are zeroed for the conv passes and
//     summed over the batch into d mask_token (vector reductions in L2; d mask_token zero on entry) instead;
//   * fp32 seam: dsplit fp32 [B,3T,HW] = [hi; lo; hi], hi = TF32-nearest(d), lo = TF32-nearest(d - hi): exact on the
//     tensor core, against [Wh; Wh; Wl] the data gradient is fp32-grade; the weight gradient reads hi;
//   * bf16 seam: dsplit bf16 [B,4T,HW] = [d1; d1; d2; d3] (three-term bf16 split): the data gradient reads the window of
//     64 rows at row 0 ([d1; d1; d2] against [W1; W2; W1]), the weight gradient the window at row T ([d1; d2; d3]).
template <bool BF16>
__global__ void conv_bwd_prep_kernel(const float* __restrict__ dX0, const int32_t* __restrict__ mask_idx, int n_masked,
                                     void* __restrict__ dsplit, float* __restrict__ d_mask_token, int B, int T, int HW) {
    pdl_sync();
    uint32_t maskbits = 0;
    for (int k = 0; k < n_masked; ++k) maskbits |= 1u << __ldg(mask_idx + k);
    const long long n4 = (long long)B * T * (HW >> 2);
    const int row4 = HW >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long bt = i / row4;
        const int qd = (int)(i - bt * row4);
        const long long b = bt / T;
        const int t = (int)(bt - b * T);
        float4 v = __ldg(reinterpret_cast<const float4*>(dX0) + i);
        if ((maskbits >> t) & 1u) {
            if (d_mask_token != nullptr)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d_mask_token + 4 * qd), "f"(v.x), "f"(v.y),
                             "f"(v.z), "f"(v.w) : "memory");
            v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const long long step = (long long)T * row4;
        if (!BF16) {
            const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
            const float4 lo = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(dsplit) + ((b * 3 * T + t) * (long long)HW)) + qd;
            dst[0] = hi;
            dst[step] = lo;
            dst[2 * step] = hi;
        } else {
            const float in[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 d1[4], d2[4], d3[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                d1[e] = __float2bfloat16_rn(in[e]);
                const float r1 = in[e] - __bfloat162float(d1[e]);
                d2[e] = __float2bfloat16_rn(r1);
                d3[e] = __float2bfloat16_rn(r1 - __bfloat162float(d2[e]));
            }
            uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dsplit) + ((b * 4 * T + t) * (long long)HW)) + qd;
            dst[0] = *reinterpret_cast<const uint2*>(d1);
            dst[step] = *reinterpret_cast<const uint2*>(d1);
            dst[2 * step] = *reinterpret_cast<const uint2*>(d2);
            dst[3 * step] = *reinterpret_cast<const uint2*>(d3);
        }
    }
}

int sm_count() {
    static int cached[16] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

template <bool BF16>
int conv_fwd_tc_impl(const void* x2, const void* Wprep, const float* pe, const float* mask_token, const int32_t* mask_idx,
                     int n_masked, int pos_embed, float* fv, float* X0, int B, int C, int HW, cudaStream_t stream) {
    using E = CElem<BF16>;
    constexpr int TPU = 2;
    CUtensorMap tmX, tmW;
    SCAT_PROPAGATE(make_map(&tmX, x2, E::BYTES, HW, (long long)B * C, HW, E::MN_BOX, E::BK,
                            BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    // K-major weight: fp32 = the first T rows of [Wh; Wh; Wl]; bf16 = [W1; W2; W3].  Rows beyond the extent are zero-filled
    SCAT_PROPAGATE(make_map(&tmW, Wprep, E::BYTES, C, BF16 ? 3 * kT : kT, C, E::BK, E::NROWS, CU_TENSOR_MAP_SWIZZLE_128B));
    ConvFwdParams p;
    p.B = B; p.C = C; p.HW = HW;
    p.units_full = HW / (128 * TPU);
    p.tail_px0 = p.units_full * 128 * TPU;
    p.tail_tiles = (HW - p.tail_px0 + 127) / 128;
    p.n_units = B * p.units_full + (p.tail_tiles ? B : 0);
    p.pe = pos_embed ? pe : nullptr; p.mask_token = mask_token; p.mask_idx = mask_idx; p.n_masked = n_masked;
    p.fv = fv; p.X0 = X0;
    auto kern = conv_fwd_tc_kernel<BF16, TPU>;
    constexpr int SMEM = 65536 + CT_STAGES * A_STAGE + 256;
    SCAT_ENSURE_SMEM(kern, SMEM);
    const int grid = min(sm_count(), p.n_units);
    SCAT_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(CT_THREADS), SMEM, stream, tmX, tmW, p));
    SCAT_CHECK_LAUNCH();
    return 0;
}

template <bool BF16>
int conv_wgrad_tc_impl(const void* dsplit, const void* x2, float* dW, int B, int C, int HW, cudaStream_t stream) {
    using E = CElem<BF16>;
    CUtensorMap tmX, tmD;
    const int d_rows = BF16 ? 4 * kT : 3 * kT;
    SCAT_PROPAGATE(make_map(&tmX, x2, E::BYTES, HW, (long long)B * C, HW, E::BK, 128, CU_TENSOR_MAP_SWIZZLE_128B));
    SCAT_PROPAGATE(make_map(&tmD, dsplit, E::BYTES, HW, (long long)B * d_rows, HW, E::BK, E::NROWS, CU_TENSOR_MAP_SWIZZLE_128B));
    ConvWgradParams p;
    p.B = B; p.C = C; p.HW = HW;
    p.kbw = (HW + E::BK - 1) / E::BK;
    p.d_rows = d_rows; p.d_off = BF16 ? kT : 0;
    p.total = (long long)(C / 128) * B * p.kbw;
    p.dW = dW;
    auto kern = conv_wgrad_tc_kernel<BF16>;
    constexpr int SMEM = CT_STAGES * (A_STAGE + E::TOK_BYTES) + 256;
    SCAT_ENSURE_SMEM(kern, SMEM);
    const int grid = (int)min((long long)sm_count(), p.total);
    SCAT_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(CT_THREADS), SMEM, stream, tmX, tmD, p));
    SCAT_CHECK_LAUNCH();
    return 0;
}

template <bool BF16>
int conv_dgrad_tc_impl(const void* dsplit, const void* Wstack, void* x2_grad, int B, int C, int HW, cudaStream_t stream) {
    using E = CElem<BF16>;
    CUtensorMap tmD, tmW, tmO;
    const int d_rows = BF16 ? 4 * kT : 3 * kT;
    const CUtensorMapSwizzle mn_sw = BF16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    SCAT_PROPAGATE(make_map(&tmD, dsplit, E::BYTES, HW, (long long)B * d_rows, HW, E::MN_BOX, E::BK, mn_sw));
    SCAT_PROPAGATE(make_map(&tmW, Wstack, E::BYTES, C, 3 * kT, C, E::MN_BOX, E::BK, mn_sw));   // row 63 of the K window: zero-filled
    SCAT_PROPAGATE(make_map(&tmO, x2_grad, BF16 ? 2 : 4, HW, (long long)B * C, HW, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
    ConvDgradParams p;
    p.B = B; p.C = C; p.HW = HW;
    p.tiles_full = HW / 128; p.tail_px0 = p.tiles_full * 128; p.has_tail = HW > p.tail_px0 ? 1 : 0;
    p.n_units = B * (p.tiles_full + p.has_tail);
    p.d_rows = d_rows; p.n_half = C / DG_N;
    auto kern = conv_dgrad_tc_kernel<BF16>;
    constexpr int KH = 64 / E::BK;
    constexpr int SMEM = KH * (DG_N / E::MN_BOX) * E::BOX_BYTES + DG_STAGES * KH * (128 / E::MN_BOX) * E::BOX_BYTES +
                         8 * 2 * 32 * 32 * (BF16 ? 2 : 4) + 256;
    SCAT_ENSURE_SMEM(kern, SMEM);
    int grid = min(sm_count(), p.n_units * p.n_half);
    grid -= grid % p.n_half;
    SCAT_CHECK_CUDA(launch_k(kern, dim3(grid), dim3(DG_THREADS), SMEM, stream, tmD, tmW, tmO, p));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int check_conv_shape(int B, int C, int HW, int T, const char* what) {
    SCAT_REQUIRE(T == kT, kErrUnsupported, "%s: built for T = 21 tokens (got %d)", what, T);
    SCAT_REQUIRE(B > 0 && C >= 256 && C <= 512 && C % 256 == 0, kErrUnsupported, "%s: channels must be 256 or 512 (got %d)", what, C);
    SCAT_REQUIRE(HW % 8 == 0 && HW >= 64, kErrUnsupported, "%s: pixels per map must be a multiple of 8 (got %d)", what, HW);
    SCAT_REQUIRE(get_encode_fn() != nullptr, kErrUnsupported, "%s: cuTensorMapEncodeTiled entry point not available", what);
    return 0;
}

}  // namespace

size_t conv_weight_prep_floats(int C, int T) { return (size_t)3 * T * C; }          // fp32 [3T,C], or 2 x bf16 [3T,C]
size_t conv_split_floats(int B, int HW, int T) { return (size_t)3 * B * T * HW; }    // fp32 [B,3T,HW], or bf16 [B,4T,HW]

int launch_conv_weight_prep(const float* Wc, void* dst, int C, int T, int x2_bf16, cudaStream_t stream) {
    const int n = T * C;
    SCAT_CHECK_CUDA(launch_k(conv_weight_prep_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, stream, Wc, dst, n, x2_bf16));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_conv_bwd_prep(const float* dX0, const int32_t* mask_idx, int n_masked, void* dsplit, float* d_mask_token, int B,
                         int T, int HW, int x2_bf16, cudaStream_t stream) {
    SCAT_REQUIRE(HW % 4 == 0 && T <= 32, kErrUnsupported, "conv bwd prep: HW%%4, T<=32");
    SCAT_REQUIRE(n_masked == 0 || mask_idx != nullptr, kErrBadArg, "conv bwd prep: mask_idx is null");
    const long long n4 = (long long)B * T * (HW / 4);
    const int grid = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    if (x2_bf16)
        SCAT_CHECK_CUDA(launch_k(conv_bwd_prep_kernel<true>, dim3(grid), dim3(256), 0, stream, dX0, mask_idx, n_masked, dsplit,
                                 d_mask_token, B, T, HW));
    else
        SCAT_CHECK_CUDA(launch_k(conv_bwd_prep_kernel<false>, dim3(grid), dim3(256), 0, stream, dX0, mask_idx, n_masked, dsplit,
                                 d_mask_token, B, T, HW));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_conv_pe_mask_fwd_tc(const void* x2, int x2_bf16, const void* Wprep, const float* pe, const float* mask_token,
                               const int32_t* mask_idx, int n_masked, int pos_embed, float* feat_visual, float* X0,
                               int B, int C, int HW, int T, cudaStream_t stream) {
    SCAT_PROPAGATE(check_conv_shape(B, C, HW, T, "conv fwd (tc)"));
    SCAT_REQUIRE(n_masked == 0 || (mask_idx != nullptr && mask_token != nullptr), kErrBadArg, "conv fwd (tc): mask_idx / mask_token is null");
    SCAT_REQUIRE(!pos_embed || pe != nullptr, kErrBadArg, "conv fwd (tc): pe is null");
    return x2_bf16 ? conv_fwd_tc_impl<true>(x2, Wprep, pe, mask_token, mask_idx, n_masked, pos_embed, feat_visual, X0, B, C, HW, stream)
                   : conv_fwd_tc_impl<false>(x2, Wprep, pe, mask_token, mask_idx, n_masked, pos_embed, feat_visual, X0, B, C, HW, stream);
}

int launch_conv_wgrad_tc(const void* dsplit, const void* x2, int x2_bf16, float* dWc, int B, int C, int HW, int T,
                         cudaStream_t stream) {
    SCAT_PROPAGATE(check_conv_shape(B, C, HW, T, "conv wgrad (tc)"));
    return x2_bf16 ? conv_wgrad_tc_impl<true>(dsplit, x2, dWc, B, C, HW, stream)
                   : conv_wgrad_tc_impl<false>(dsplit, x2, dWc, B, C, HW, stream);
}

int launch_conv_dgrad_tc(const void* dsplit, const void* Wprep, int x2_bf16, void* x2_grad, int B, int C, int HW, int T,
                         cudaStream_t stream) {
    SCAT_PROPAGATE(check_conv_shape(B, C, HW, T, "conv dgrad (tc)"));
    // bf16: the data-gradient stack [W1; W2; W1] sits behind the forward's [W1; W2; W3]
    const void* Wstack = x2_bf16 ? (const void*)(reinterpret_cast<const __nv_bfloat16*>(Wprep) + (size_t)3 * T * C) : Wprep;
    return x2_bf16 ? conv_dgrad_tc_impl<true>(dsplit, Wstack, x2_grad, B, C, HW, stream)
                   : conv_dgrad_tc_impl<false>(dsplit, Wstack, x2_grad, B, C, HW, stream);
}

}  // namespace scat
