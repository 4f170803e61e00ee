// ff_score_tc: the Blackwell-native scorer (GB200_PREC_BF16_TC and GB200_PREC_F16X3_TC).
//
// Two operand precisions share one kernel template:
//   BF16   bf16 operands, fp32 accumulate: 1 MMA per K step, tanh.approx -- the fastest path (~8e-3 typical error on yhat);
//   F16X3  fp32-grade: every fp32 operand is split into hi + lo fp16 halves (x = hi + lo + O(2^-22 x)) while it is
//          staged, and A_hi.B_hi + A_hi.B_lo + A_lo.B_hi is accumulated in fp32 in tensor memory (3 MMAs per K step,
//          the tensor pipe has the room); activations through ex2/rcp instead of tanh.approx.  Product error < 2^-20,
//          results agree with the fp32 FMA path to ~1e-6.  fp16 overflows at 65 504: hidden activations must be
//          bounded (tanh / sigmoid), and the scaled INPUT row is multiplied by a per-row power of two s (1 unless
//          the row's largest magnitude exceeds 2^15; the ones column carries s, the first accumulator is
//          multiplied by 1/s before its activation -- exact), so any finite input keeps its 22 bits.
//
//   MinMax-scale -> 2E+1 chained Dense layers on tcgen05 tensor cores -> anomaly columns,
// one 128-row tile of one Machine at a time per warpgroup, everything between the HBM read of the
// tile and the HBM write of the result columns stays on chip.
//
// Structure (one persistent CTA per SM, NWG warpgroups = NWG independent "tile engines"):
//   * the Machine's operand image (bf16 weights pre-packed in the tcgen05 canonical K-major
//     no-swizzle layout, the bias as one extra K row) is brought in once per CTA per Machine with ONE 1-D bulk
//     async copy (cp.async.bulk -> UBLKCP, mbarrier complete_tx);
//   * each warpgroup: bulk-copies its [128 x T] fp32 tile (a tile of consecutive rows of a
//     row-major matrix is one contiguous run in HBM), builds the bf16 A operand in shared memory
//     (thread = row = TMEM lane), then for every layer one elected thread issues Kp/16
//     tcgen05.mma (M=128, N=Np, K=16, fp32 accumulate in TMEM) + tcgen05.commit; all 128
//     threads pull the accumulator back with tcgen05.ld (32x32b.x16) (bias already added by the
//     GEMM through a ones column of A), apply the
//     activation and write the next layer's A operand;
//   * final epilogue: |yhat - y| is written IN PLACE over the y tile, the row totals are reduced in
//     registers, and the tile goes back to HBM as full-row contiguous runs (tag-anomaly-unscaled,
//     and, scaled per column on the way out, tag-anomaly-scaled and anomaly-confidence); yhat then
//     overwrites the tile and is copied out the same way.
// Algorithmic HBM traffic per row: 4*T_in read + 4*(3*T_out + 2) written (+4*(T_out+1) with
// thresholds); weights are amortised over ~780 tiles per Machine.
//
// Replaces models.py:289-300 + diff.py:336-444.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace {

constexpr int TILE = 128;
constexpr int WG_THREADS = 128;
constexpr int MAX_WG = 5;

// ---------------------------------------------------------------- packed operand image
struct PackLayout {
    int n_layers;
    int Kp[GB200_MAX_LAYERS], Np[GB200_MAX_LAYERS];
    int w_off[GB200_MAX_LAYERS];     // byte offset of layer l's B operand (the hi half for F16X3)
    int w_lo[GB200_MAX_LAYERS];      // F16X3: byte offset of the lo half
    int total_bytes;                 // multiple of 16
    int max_Kp, max_Np;
};

PackLayout make_layout(const gb200_ff_arch* a, int prec = GB200_PREC_BF16_TC) {
    PackLayout p{};
    p.n_layers = a->n_layers;
    int off = 0;
    for (int l = 0; l < a->n_layers; ++l) {
        p.Kp[l] = gb_round_up(a->widths[l] + 1, 16);      // +1: the bias row (A carries a ones column)
        p.Np[l] = gb_round_up(a->widths[l + 1], 16);
        p.w_off[l] = off; off += p.Kp[l] * p.Np[l] * 2;
        if (prec == GB200_PREC_F16X3_TC) { p.w_lo[l] = off; off += p.Kp[l] * p.Np[l] * 2; }
        if (p.Kp[l] > p.max_Kp) p.max_Kp = p.Kp[l];
        if (p.Np[l] > p.max_Np) p.max_Np = p.Np[l];
    }
    p.total_bytes = gb_round_up(off, 16);
    return p;
}

template <int PREC>
__global__ void pack_kernel(gb200_ff_arch arch, PackLayout lay, int64_t n_params,
                            const float* __restrict__ params, uint8_t* __restrict__ packed) {
    const int m = blockIdx.x;
    const float* P = params + (size_t)m * n_params;
    uint8_t* out = packed + (size_t)m * lay.total_bytes;
    int64_t go = 0;
    for (int l = 0; l < arch.n_layers; ++l) {
        const int win = arch.widths[l], wout = arch.widths[l + 1];
        const int Kp = lay.Kp[l], Np = lay.Np[l];
        __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(out + lay.w_off[l]);
        __half* Bh = reinterpret_cast<__half*>(out + lay.w_off[l]);
        __half* Bl = reinterpret_cast<__half*>(out + lay.w_lo[l]);
        // canonical K-major, no swizzle: element (n, k) at
        //   (k/8) * (Np/8)*64 + (n/8)*64 + (n%8)*8 + (k%8)      [bf16 elements]
        for (int i = threadIdx.x; i < Kp * Np; i += blockDim.x) {
            const int k = i / Np, n = i - k * Np;
            float w = 0.0f;
            if (n < wout) {
                if (k < win) w = P[go + (int64_t)k * wout + n];
                else if (k == win) w = P[go + (int64_t)win * wout + n];      // bias row, multiplied by A's ones column
            }
            const int at = (k >> 3) * (Np >> 3) * 64 + (n >> 3) * 64 + (n & 7) * 8 + (k & 7);
            if (PREC == GB200_PREC_F16X3_TC) {
                const __half hi = __float2half_rn(w);
                Bh[at] = hi; Bl[at] = __float2half_rn(w - __half2float(hi));
            } else {
                B[at] = __float2bfloat16_rn(w);
            }
        }
        go += (int64_t)win * wout + wout;
    }
}

using namespace gbptx;

// ACC: the fp32-grade path cannot use tanh.approx (2^-11 relative error): tanh(z) = 1 - 2 / (1 + e^(2z))
// through ex2.approx + rcp.approx, absolute error ~2e-7, exact limits at +-inf, NaN propagates.
template <int ACT, bool ACC = false>
__device__ __forceinline__ float act_t(float z) {
    if (ACT == GB200_ACT_TANH) {
        if (ACC) return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * z));
        float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(z)); return r;
    }
    if (ACT == GB200_ACT_RELU) return fmaxf(z, 0.0f);
    if (ACT == GB200_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-z));
    if (ACT == GB200_ACT_ELU) return z > 0.0f ? z : __expf(z) - 1.0f;
    if (ACT == GB200_ACT_SOFTPLUS) return z > 15.0f ? z : __logf(1.0f + __expf(z));
    return z;
}

// dispatch a per-layer activation code to a compile-time template argument (once per layer, not per element)
#define GB_DISPATCH_ACT(code, CALL)                                    \
    switch (code) {                                                    \
        case GB200_ACT_TANH:     { constexpr int ACT = GB200_ACT_TANH; CALL; break; }     \
        case GB200_ACT_RELU:     { constexpr int ACT = GB200_ACT_RELU; CALL; break; }     \
        case GB200_ACT_SIGMOID:  { constexpr int ACT = GB200_ACT_SIGMOID; CALL; break; }  \
        case GB200_ACT_ELU:      { constexpr int ACT = GB200_ACT_ELU; CALL; break; }      \
        case GB200_ACT_SOFTPLUS: { constexpr int ACT = GB200_ACT_SOFTPLUS; CALL; break; } \
        default:                 { constexpr int ACT = GB200_ACT_LINEAR; CALL; break; }   \
    }

struct TcArgs {
    gb200_ff_arch arch;
    PackLayout lay;
    const int64_t* row_lo; const int64_t* row_hi; const int32_t* tile_off;
    int n_machines, tiles_total;
    const uint8_t* packed;
    const float* in_scale; const float* in_min; const float* err_scale;
    const float* feat_thr; const float* agg_thr;
    const float* x; const float* y;
    float* model_out; float* tag_scaled; float* tag_unscaled;
    float* total_scaled; float* total_unscaled; float* conf; float* total_conf;
    int nwg;                 // warpgroups per CTA
    int tmem_cols_wg;        // accumulator columns per warpgroup
    int tmem_cols_total;     // power of two >= 32
    int xtile_bytes;         // 128*T_in*4 rounded to 128
    int ytile_bytes;         // 0 when y aliases x
    int tmem_a_off;          // column offset of the 16-bit A operand inside a warpgroup's TMEM slice
    int tmem_a_lo;           // F16X3: column distance from the hi half of A to its lo half
    int vp;                  // padded length of each per-Machine vector (multiple of 16 floats)
};

__device__ __forceinline__ int find_machine(const int32_t* tile_off, int n, int tile) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (tile_off[mid] <= tile) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// fp32 pair -> fp16 hi pair + fp16 lo pair (x = hi + lo + O(2^-22 |x|))
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 activations -> 4 TMEM columns of the A operand (this thread's lane = its row); F16X3: 4 hi + 4 lo columns
template <int PREC>
__device__ __forceinline__ void store_a8(uint32_t taddr, int lo_delta, const float* v) {
    if (PREC == GB200_PREC_F16X3_TC) {
        uint32_t hi[4], lo[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) split_f16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        tmem_st4(taddr, hi); tmem_st4(taddr + lo_delta, lo);
    } else {
        uint32_t pk[4] = {pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])};
        tmem_st4(taddr, pk);
    }
}
template <int PREC>
__device__ __forceinline__ void store_a16(uint32_t taddr, int lo_delta, const float* v) {
    if (PREC == GB200_PREC_F16X3_TC) {
        uint32_t hi[8], lo[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) split_f16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        tmem_st8(taddr, hi); tmem_st8(taddr + lo_delta, lo);
    } else {
        uint32_t pk[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        tmem_st8(taddr, pk);
    }
}
// F16X3: power-of-two scale that brings a row whose largest magnitude is `amax` below 2^15 (1 when it already is)
__device__ __forceinline__ void row_pow2_scale(float amax, float& s, float& inv_s) {
    s = 1.0f; inv_s = 1.0f;
    if (amax > 32768.0f) {                                     // false for NaN
        int e = (__float_as_int(amax) >> 23) - 127;            // amax < 2^(e+1)
        e = min(e, 127);
        s = __int_as_float((127 + 14 - e) << 23);              // amax * s < 2^15
        inv_s = __int_as_float((127 - 14 + e) << 23);
    }
}
__device__ __forceinline__ void load16_bcast(const float* p, float* o) {       // 16-byte aligned broadcast loads
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
        o[4 * q] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
    }
}

// Copy one warp's rows (cnt floats, contiguous in smem and in HBM) out, optionally scaled per
// column on the way (MODE 1).  `pat` is the per-column factor repeated periodically
// (pat[i] = vec[i % T], i < T + 4) so the 4 factors of a float4 are one contiguous read.
template <int MODE>
__device__ __forceinline__ void warp_copy_out(float* __restrict__ dst, const float* __restrict__ src, int cnt,
                                              int T, const float* __restrict__ pat, int lane, bool vec16) {
    if (vec16) {
        int col = (lane * 4) % T;
        const int step = 128 % T;
        const bool even = (T & 1) == 0;
        for (int o = lane * 4; o < cnt; o += 128) {
            float4 v = *reinterpret_cast<const float4*>(src + o);
            if (MODE != 0) {
                if (even) {
                    const float2 p0 = *reinterpret_cast<const float2*>(pat + col), p1 = *reinterpret_cast<const float2*>(pat + col + 2);
                    v.x *= p0.x; v.y *= p0.y; v.z *= p1.x; v.w *= p1.y;
                } else {
                    v.x *= pat[col]; v.y *= pat[col + 1]; v.z *= pat[col + 2]; v.w *= pat[col + 3];
                }
                col += step; if (col >= T) col -= T;
            }
            __stcs(reinterpret_cast<float4*>(dst + o), v);
        }
    } else {
        for (int o = lane; o < cnt; o += 32) {
            float v = src[o];
            if (MODE == 1) v *= pat[o % T];
            dst[o] = v;
        }
    }
}

// hidden layer: accumulator (TMEM, bias already inside the GEMM) -> activation -> bf16 A operand of
// the next layer, written straight back to TENSOR MEMORY (tcgen05.st): activations never touch
// shared memory.  Column `wout` of the next A is the ones column that carries the next bias.
template <int ACT, int PREC, bool SCALED>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_lane, int n_chunks, int wout, int kp_next,
                                                uint32_t tmem_a_lane, int lo_delta, float inv_s) {
    constexpr bool ACC = PREC == GB200_PREC_F16X3_TC;
    const int c_one = wout >> 4, j_one = wout & 15;
    for (int c = 0; c < n_chunks; ++c) {
        float v[16];
        tmem_ld16(tmem_lane + c * 16, v);
        #pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act_t<ACT, ACC>(SCALED ? v[j] * inv_s : v[j]);
        if (c == c_one) {
            #pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (j == j_one) ? 1.0f : v[j];
        }
        store_a16<PREC>(tmem_a_lane + c * 8, lo_delta, v);
    }
    if (kp_next > n_chunks * 16) {                     // wout is a multiple of 16: the ones column opens a new chunk
        float v[16];
        #pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (j == j_one) ? 1.0f : 0.0f;
        store_a16<PREC>(tmem_a_lane + n_chunks * 8, lo_delta, v);
    }
    tmem_wait_st();
}

// final layer, pass 1: d = |yhat - y| written in place over the y tile; returns the two row sums
template <int ACT, bool EVEN, bool ACC>
__device__ __forceinline__ void final_pass1(uint32_t tmem_lane,
                                            const float* __restrict__ v_es, float* __restrict__ yrow,
                                            int T_out, float& su_out, float& ss_out) {
    float su = 0.0f, ss = 0.0f;
    const int full = T_out >> 4;
    for (int c = 0; c < full; ++c) {
        float v[16], e[16], yv[16];
        tmem_ld16(tmem_lane + c * 16, v);
        load16_bcast(v_es + c * 16, e);
        if (EVEN) {
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 t = *reinterpret_cast<const float2*>(yrow + c * 16 + 2 * j);
                yv[2 * j] = t.x; yv[2 * j + 1] = t.y;
            }
        } else {
            #pragma unroll
            for (int j = 0; j < 16; ++j) yv[j] = yrow[c * 16 + j];
        }
        #pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float d = fabsf(act_t<ACT, ACC>(v[j]) - yv[j]);
            const float s = d * e[j];
            su = fmaf(d, d, su); ss = fmaf(s, s, ss);
            yv[j] = d;
        }
        if (EVEN) {
            #pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float2*>(yrow + c * 16 + 2 * j) = make_float2(yv[2 * j], yv[2 * j + 1]);
        } else {
            #pragma unroll
            for (int j = 0; j < 16; ++j) yrow[c * 16 + j] = yv[j];
        }
    }
    if (T_out & 15) {                                   // ragged last chunk
        const int n0 = full * 16;
        float v[16];
        tmem_ld16(tmem_lane + n0, v);
        #pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = n0 + j;
            if (n < T_out) {
                const float d = fabsf(act_t<ACT, ACC>(v[j]) - yrow[n]);
                const float s = d * v_es[n];
                su = fmaf(d, d, su); ss = fmaf(s, s, ss);
                yrow[n] = d;
            }
        }
    }
    su_out = su; ss_out = ss;
}

// final layer, pass 2: yhat overwrites the tile
template <int ACT, bool EVEN, bool ACC>
__device__ __forceinline__ void final_pass2(uint32_t tmem_lane, float* __restrict__ yrow, int T_out) {
    const int full = T_out >> 4;
    for (int c = 0; c < full; ++c) {
        float v[16];
        tmem_ld16(tmem_lane + c * 16, v);
        #pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act_t<ACT, ACC>(v[j]);
        if (EVEN) {
            #pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float2*>(yrow + c * 16 + 2 * j) = make_float2(v[2 * j], v[2 * j + 1]);
        } else {
            #pragma unroll
            for (int j = 0; j < 16; ++j) yrow[c * 16 + j] = v[j];
        }
    }
    if (T_out & 15) {
        const int n0 = full * 16;
        float v[16];
        tmem_ld16(tmem_lane + n0, v);
        #pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < T_out) yrow[n0 + j] = act_t<ACT, ACC>(v[j]);
    }
}

template <int PREC>
__global__ void __launch_bounds__(MAX_WG * WG_THREADS, 1)
ff_score_tc_kernel(const __grid_constant__ TcArgs a) {
    constexpr bool ACC = PREC == GB200_PREC_F16X3_TC;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ __align__(8) uint64_t x_bar[MAX_WG];
    __shared__ __align__(8) uint64_t mma_bar[MAX_WG];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_seg_m, s_seg_end;

    const int tid = threadIdx.x;
    const int wg = tid / WG_THREADS;
    const int wtid = tid - wg * WG_THREADS;          // row within the tile == TMEM lane
    const int warp_in_wg = wtid >> 5, lane = tid & 31;
    const int L = a.arch.n_layers;
    const int T_in = a.arch.widths[0], T_out = a.arch.widths[L];
    const bool y_sep = a.ytile_bytes != 0;
    const float inv_T = 1.0f / (float)T_out;

    // ---- shared memory carve-up
    uint8_t* w_img = smem;                                                    // operand image
    float* vecs = reinterpret_cast<float*>(smem + a.lay.total_bytes);         // per-Machine vectors, padded to vp
    float* v_scale = vecs; float* v_min = vecs + a.vp; float* v_es = vecs + 2 * a.vp;
    float* pat_es = vecs + 3 * a.vp; float* pat_ift = vecs + 4 * a.vp + 16;          // periodic copies, T_out + 4 long
    uint8_t* wg_base = smem + a.lay.total_bytes + (5 * a.vp + 32) * 4
                     + (size_t)wg * (a.xtile_bytes + a.ytile_bytes);
    float* xbuf = reinterpret_cast<float*>(wg_base);
    float* ybuf = y_sep ? reinterpret_cast<float*>(wg_base + a.xtile_bytes) : xbuf;

    if (tid == 0) {
        mbar_init(&w_bar, 1);
        for (int g = 0; g < a.nwg; ++g) { mbar_init(&x_bar[g], 1); mbar_init(&mma_bar[g], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(&s_tmem_base, a.tmem_cols_total);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = s_tmem_base + (uint32_t)(wg * a.tmem_cols_wg);          // column offset
    const uint32_t tmem_lane = tmem_acc + ((uint32_t)(warp_in_wg * 32) << 16);        // this warp's lanes
    const uint32_t tmem_a = tmem_acc + (uint32_t)a.tmem_a_off;                        // 16-bit A operand columns
    const int lo_delta = a.tmem_a_lo;
    const uint32_t tmem_a_lane = tmem_a + ((uint32_t)(warp_in_wg * 32) << 16);

    const int per_cta = (a.tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta;
    const int t_end = min(t_begin + per_cta, a.tiles_total);
    uint32_t w_phase = 0, x_phase = 0, mma_phase = 0;

    int seg_begin = t_begin;
    while (seg_begin < t_end) {
        // ---- segment = run of this CTA's tiles that belong to one Machine
        if (tid == 0) {
            const int m = find_machine(a.tile_off, a.n_machines, seg_begin);
            s_seg_m = m; s_seg_end = min(t_end, a.tile_off[m + 1]);
            mbar_expect_tx(&w_bar, (uint32_t)a.lay.total_bytes);
            bulk_g2s(w_img, a.packed + (size_t)m * a.lay.total_bytes, (uint32_t)a.lay.total_bytes, &w_bar);
        }
        __syncthreads();
        const int m = s_seg_m, seg_end = s_seg_end;
        for (int i = tid; i < a.vp; i += blockDim.x) {
            v_scale[i] = (i < T_in) ? (a.in_scale ? a.in_scale[(size_t)m * T_in + i] : 1.0f) : 0.0f;
            v_min[i] = (i < T_in && a.in_min) ? a.in_min[(size_t)m * T_in + i] : 0.0f;
            v_es[i] = (i < T_out) ? (a.err_scale ? fabsf(a.err_scale[(size_t)m * T_out + i]) : 1.0f) : 0.0f;
        }
        for (int i = tid; i < T_out + 4; i += blockDim.x) {
            const int c = i % T_out;
            pat_es[i] = a.err_scale ? fabsf(a.err_scale[(size_t)m * T_out + c]) : 1.0f;
            pat_ift[i] = a.feat_thr ? 1.0f / a.feat_thr[(size_t)m * T_out + c] : 0.0f;
        }
        mbar_wait(&w_bar, w_phase); w_phase ^= 1;
        __syncthreads();
        const int64_t m_row0 = a.row_lo[m], m_row1 = a.row_hi[m];
        const float inv_agg = a.agg_thr ? 1.0f / a.agg_thr[m] : 1.0f;

        for (int tile = seg_begin + wg; tile < seg_end; tile += a.nwg) {
            const int64_t row0 = m_row0 + (int64_t)(tile - a.tile_off[m]) * TILE;
            const int nrows = (int)min((int64_t)TILE, m_row1 - row0);
            // ---- tile load: one contiguous run of nrows*T floats
            const float* xsrc = a.x + row0 * T_in;
            const float* ysrc = y_sep ? a.y + row0 * T_out : xsrc;
            const uint32_t xbytes = (uint32_t)nrows * T_in * 4, ybytes = (uint32_t)nrows * T_out * 4;
            const bool x16 = ((reinterpret_cast<uintptr_t>(xsrc) | xbytes) & 15) == 0;
            const bool y16 = !y_sep || ((reinterpret_cast<uintptr_t>(ysrc) | ybytes) & 15) == 0;
            if (x16 && y16) {
                if (wtid == 0) {
                    mbar_expect_tx(&x_bar[wg], xbytes + (y_sep ? ybytes : 0));
                    bulk_g2s(xbuf, xsrc, xbytes, &x_bar[wg]);
                    if (y_sep) bulk_g2s(ybuf, ysrc, ybytes, &x_bar[wg]);
                }
                mbar_wait(&x_bar[wg], x_phase); x_phase ^= 1;
            } else {
                for (int i = wtid; i < nrows * T_in; i += WG_THREADS) xbuf[i] = xsrc[i];
                if (y_sep) for (int i = wtid; i < nrows * T_out; i += WG_THREADS) ybuf[i] = ysrc[i];
                named_bar_sync(1 + wg, WG_THREADS);
            }
            // ---- A operand of layer 0: bf16(x*scale+min), written to tensor memory.  Rows past nrows
            // hold stale data: rows never mix inside a GEMM and those rows are never stored.
            float row_s = 1.0f, row_inv_s = 1.0f;
            {
                const int Kp = a.lay.Kp[0];
                const float* xr = xbuf + wtid * T_in;
                const int full = T_in >> 3;
                if (ACC) {                       // fp16 operands: bring the row into range (see the header note)
                    float amax = 0.0f;
                    for (int k = 0; k < T_in; ++k) amax = fmaxf(amax, fabsf(fmaf(xr[k], v_scale[k], v_min[k])));
                    row_pow2_scale(amax, row_s, row_inv_s);
                }
                if ((T_in & 1) == 0) {
                    for (int c = 0; c < full; ++c) {
                        float v[8];
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 t = *reinterpret_cast<const float2*>(xr + c * 8 + 2 * j);
                            v[2 * j] = t.x; v[2 * j + 1] = t.y;
                        }
                        const float4 s0 = *reinterpret_cast<const float4*>(v_scale + c * 8), s1 = *reinterpret_cast<const float4*>(v_scale + c * 8 + 4);
                        const float4 m0 = *reinterpret_cast<const float4*>(v_min + c * 8), m1 = *reinterpret_cast<const float4*>(v_min + c * 8 + 4);
                        v[0] = fmaf(v[0], s0.x, m0.x); v[1] = fmaf(v[1], s0.y, m0.y); v[2] = fmaf(v[2], s0.z, m0.z); v[3] = fmaf(v[3], s0.w, m0.w);
                        v[4] = fmaf(v[4], s1.x, m1.x); v[5] = fmaf(v[5], s1.y, m1.y); v[6] = fmaf(v[6], s1.z, m1.z); v[7] = fmaf(v[7], s1.w, m1.w);
                        if (ACC) {
                            #pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= row_s;
                        }
                        store_a8<PREC>(tmem_a_lane + c * 4, lo_delta, v);
                    }
                } else {
                    for (int c = 0; c < full; ++c) {
                        float v[8];
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(xr[c * 8 + j], v_scale[c * 8 + j], v_min[c * 8 + j]) * row_s;
                        store_a8<PREC>(tmem_a_lane + c * 4, lo_delta, v);
                    }
                }
                // the chunk holding column T_in (the ones column that carries the first bias), then zero padding
                for (int c = full; c < Kp / 8; ++c) {
                    float v[8];
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = c * 8 + j;
                        v[j] = k < T_in ? fmaf(xr[k], v_scale[k], v_min[k]) * row_s : (k == T_in ? row_s : 0.0f);
                    }
                    store_a8<PREC>(tmem_a_lane + c * 4, lo_delta, v);
                }
                tmem_wait_st();
            }
            // ---- the Dense stack
            for (int l = 0; l < L; ++l) {
                const int Kp = a.lay.Kp[l], Np = a.lay.Np[l];
                tc_fence_before();              // A stores (tcgen05.st, already waited) ordered before the MMAs
                named_bar_sync(1 + wg, WG_THREADS);
                if (wtid == 0) {
                    tc_fence_after();
                    const uint32_t idesc = ACC ? make_idesc_f16(TILE, Np) : make_idesc(TILE, Np);
                    const uint32_t b_addr = smem_u32(w_img + a.lay.w_off[l]);
                    const uint32_t b_lo_addr = smem_u32(w_img + a.lay.w_lo[l]);
                    const uint32_t b_lbo = (uint32_t)(Np / 8) * 128;
                    for (int ks = 0; ks < Kp / 16; ++ks) {
                        const uint64_t db = make_desc(b_addr + ks * 2 * b_lbo, b_lbo, 128);
                        umma_bf16_ts(tmem_acc, tmem_a + ks * 8, db, idesc, ks > 0 ? 1u : 0u);      // K=16 halves = 8 columns of A
                        if (ACC) {                          // + A_hi.B_lo + A_lo.B_hi: the fp32-grade split product
                            const uint64_t dl = make_desc(b_lo_addr + ks * 2 * b_lbo, b_lbo, 128);
                            umma_bf16_ts(tmem_acc, tmem_a + ks * 8, dl, idesc, 1u);
                            umma_bf16_ts(tmem_acc, tmem_a + lo_delta + ks * 8, db, idesc, 1u);
                        }
                    }
                    umma_commit(&mma_bar[wg]);
                }
                mbar_wait(&mma_bar[wg], mma_phase); mma_phase ^= 1;
                tc_fence_after();
                if (l == L - 1) break;
                if (ACC && l == 0) {
                    GB_DISPATCH_ACT(a.arch.acts[l], (hidden_epilogue<ACT, PREC, true>(tmem_lane, Np / 16, a.arch.widths[l + 1], a.lay.Kp[l + 1], tmem_a_lane, lo_delta, row_inv_s)));
                } else {
                    GB_DISPATCH_ACT(a.arch.acts[l], (hidden_epilogue<ACT, PREC, false>(tmem_lane, Np / 16, a.arch.widths[l + 1], a.lay.Kp[l + 1], tmem_a_lane, lo_delta, 1.0f)));
                }
            }
            // ---- final epilogue
            const int code = a.arch.acts[L - 1];
            float* yrow = ybuf + wtid * T_out;
            const bool even = (T_out & 1) == 0;
            float su, ss;
            if (even) { GB_DISPATCH_ACT(code, (final_pass1<ACT, true, ACC>(tmem_lane, v_es, yrow, T_out, su, ss))); }
            else      { GB_DISPATCH_ACT(code, (final_pass1<ACT, false, ACC>(tmem_lane, v_es, yrow, T_out, su, ss))); }
            if (wtid < nrows) {
                const int64_t row = row0 + wtid;
                const float ts = ss * inv_T;
                if (a.total_scaled) a.total_scaled[row] = ts;
                if (a.total_unscaled) a.total_unscaled[row] = su * inv_T;
                if (a.total_conf && a.agg_thr) a.total_conf[row] = ts * inv_agg;
            }
            __syncwarp();
            {
                const int wrows = max(0, min(32, nrows - warp_in_wg * 32));
                const int cnt = wrows * T_out;
                const int64_t goff = (row0 + warp_in_wg * 32) * T_out;
                const float* src = ybuf + warp_in_wg * 32 * T_out;
                const bool v16 = (((goff * 4) | (int64_t)(cnt * 4)) & 15) == 0 && ((warp_in_wg * 32 * T_out * 4) & 15) == 0
                                 && ((reinterpret_cast<uintptr_t>(a.model_out) | reinterpret_cast<uintptr_t>(a.tag_scaled) |
                                      reinterpret_cast<uintptr_t>(a.tag_unscaled) | reinterpret_cast<uintptr_t>(a.conf)) & 15) == 0;
                if (cnt > 0) {
                    if (a.tag_unscaled) warp_copy_out<0>(a.tag_unscaled + goff, src, cnt, T_out, nullptr, lane, v16);
                    if (a.tag_scaled) warp_copy_out<1>(a.tag_scaled + goff, src, cnt, T_out, pat_es, lane, v16);
                    if (a.conf && a.feat_thr) warp_copy_out<1>(a.conf + goff, src, cnt, T_out, pat_ift, lane, v16);
                }
                __syncwarp();
                if (a.model_out) {
                    if (even) { GB_DISPATCH_ACT(code, (final_pass2<ACT, true, ACC>(tmem_lane, yrow, T_out))); }
                    else      { GB_DISPATCH_ACT(code, (final_pass2<ACT, false, ACC>(tmem_lane, yrow, T_out))); }
                    __syncwarp();
                    if (cnt > 0) warp_copy_out<0>(a.model_out + goff, src, cnt, T_out, nullptr, lane, v16);
                }
            }
            // the tile buffers are about to be overwritten through the async proxy
            fence_proxy_async();
            tc_fence_before();
            named_bar_sync(1 + wg, WG_THREADS);
        }
        __syncthreads();            // all warpgroups done with this Machine's operand image
        seg_begin = seg_end;
    }

    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(s_tmem_base, a.tmem_cols_total);
}

}  // namespace

int64_t gb_ff_packed_bytes(const gb200_ff_arch* arch, int prec) {
    PackLayout lay = make_layout(arch, prec);
    // eligible when the image + one warpgroup's buffers fit the 227 KB shared-memory window and
    // a layer fits one UMMA (N <= 256) and one accumulator slice (<= 512 TMEM columns)
    if (lay.max_Np > 256 || lay.max_Kp > 256) return 0;
    const int T_in = arch->widths[0], T_out = arch->widths[arch->n_layers];
    const int64_t per_wg = gb_round_up(TILE * T_in * 4, 128) + gb_round_up(TILE * T_out * 4, 128);
    if (lay.total_bytes + 4096 + per_wg > 227 * 1024 - 2048) return 0;
    if (prec == GB200_PREC_F16X3_TC) {
        // fp16 operands: hidden activations must be bounded (see the header note), and there must be one
        if (arch->n_layers < 2 || lay.max_Np + lay.max_Kp > 512) return 0;
        for (int l = 0; l + 1 < arch->n_layers; ++l)
            if (arch->acts[l] != GB200_ACT_TANH && arch->acts[l] != GB200_ACT_SIGMOID) return 0;
    }
    return lay.total_bytes;
}

int gb_launch_ff_pack(const gb200_ff_arch* arch, int prec, int n_machines, const float* params, void* packed,
                      cudaStream_t stream) {
    if (n_machines <= 0) return GB_OK;
    PackLayout lay = make_layout(arch, prec);
    if (prec == GB200_PREC_F16X3_TC)
        pack_kernel<GB200_PREC_F16X3_TC><<<n_machines, 256, 0, stream>>>(*arch, lay, gb200_ff_param_count(arch), params, (uint8_t*)packed);
    else
        pack_kernel<GB200_PREC_BF16_TC><<<n_machines, 256, 0, stream>>>(*arch, lay, gb200_ff_param_count(arch), params, (uint8_t*)packed);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_ff_score_tc(const gb200_fleet* f, const gb200_ff_arch* arch, int prec, const void* packed,
                          const float* in_scale, const float* in_min, const float* err_scale,
                          const float* feat_thr, const float* agg_thr, const float* x, const float* y,
                          float* model_out, float* tag_scaled, float* tag_unscaled,
                          float* total_scaled, float* total_unscaled, float* conf, float* total_conf,
                          cudaStream_t stream) {
    if (f->tiles_total == 0) return GB_OK;
    const bool x3 = prec == GB200_PREC_F16X3_TC;
    TcArgs a{};
    a.arch = *arch; a.lay = make_layout(arch, prec);
    a.row_lo = f->d_row_lo; a.row_hi = f->d_row_hi; a.tile_off = f->d_tile_off;
    a.n_machines = f->n_machines; a.tiles_total = f->tiles_total;
    a.packed = (const uint8_t*)packed;
    a.in_scale = in_scale; a.in_min = in_min; a.err_scale = err_scale; a.feat_thr = feat_thr; a.agg_thr = agg_thr;
    a.x = x; a.y = (y == x) ? nullptr : y;
    a.model_out = model_out; a.tag_scaled = tag_scaled; a.tag_unscaled = tag_unscaled;
    a.total_scaled = total_scaled; a.total_unscaled = total_unscaled; a.conf = conf; a.total_conf = total_conf;
    const int T_in = arch->widths[0], T_out = arch->widths[arch->n_layers];
    a.xtile_bytes = gb_round_up(TILE * T_in * 4, 128);
    a.ytile_bytes = a.y ? gb_round_up(TILE * T_out * 4, 128) : 0;
    a.vp = gb_round_up((T_in > T_out ? T_in : T_out) + 4, 16);
    // TMEM slice of a warpgroup: fp32 accumulator (max_Np columns) + 16-bit A operand (max_Kp/2 columns, twice for hi + lo)
    a.tmem_a_off = a.lay.max_Np;
    a.tmem_a_lo = x3 ? a.lay.max_Kp / 2 : 0;
    const int cols = a.lay.max_Np + (x3 ? a.lay.max_Kp : a.lay.max_Kp / 2);
    a.tmem_cols_wg = cols;
    const size_t cap = 227 * 1024 - 1024;
    const size_t fixed = (size_t)a.lay.total_bytes + (size_t)(5 * a.vp + 32) * 4;
    const size_t per_wg = (size_t)a.xtile_bytes + a.ytile_bytes;
    int nwg = MAX_WG;
    { const char* e = getenv("GB200_FF_NWG"); if (e && atoi(e) >= 1 && atoi(e) <= MAX_WG) nwg = atoi(e); }   // tuning knob
    while (nwg > 1 && (fixed + nwg * per_wg > cap || nwg * cols > 512)) --nwg;
    GB_REQUIRE(fixed + nwg * per_wg <= cap && nwg * cols <= 512, "ff_score_tc: topology does not fit in shared / tensor memory");
    a.nwg = nwg;
    int tot = 32; while (tot < nwg * cols) tot <<= 1;
    a.tmem_cols_total = tot;
    const size_t smem = fixed + nwg * per_wg;
    auto* kern = x3 ? ff_score_tc_kernel<GB200_PREC_F16X3_TC> : ff_score_tc_kernel<GB200_PREC_BF16_TC>;
    GB_CUDA_CHECK(gb_allow_max_smem(kern));
    int grid = f->sm_count;
    const int min_tiles_per_cta = nwg;            // keep every warpgroup of a CTA busy
    const int max_grid = (f->tiles_total + min_tiles_per_cta - 1) / min_tiles_per_cta;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    kern<<<grid, nwg * WG_THREADS, smem, stream>>>(a);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}
