// lstm_tc: the stacked-LSTM forward (KerasLSTMBaseEstimator.predict, models.py:618-660) on tcgen05.
//
// Everything a tile needs is kept IN HBM IN THE UMMA CANONICAL LAYOUT, so
// every operand is a plain 1-D bulk async copy (no tensor maps, no repacking on the way in):
//   * Xc   : bf16(x*scale+min) of the chunk's rows, K-chunk-major [Kx/8][rows][8]: the 128
//            consecutive rows of a window tile at time t are one contiguous 2 KB run per K chunk
//            (windows are never materialised: window w at time t is row w+t);
//   * h_l  : bf16 hidden state, per 128-window tile [Kh/8][128][8] (ping-pong over t);
//   * W|U  : per 16-unit block the B operand [K/8][8][8][8] with columns ordered unit*4+gate, so
//            a 16-column tcgen05.ld delivers the i,f,c,o pre-activations of 4 units;
//   * c_l  : fp32 cell state, [tile][unit][128 windows] (coalesced, constant strides inside a tile).
// One PERSISTENT CTA per 128-window tile walks all L x n_layers (timestep, layer) steps itself
// (window tiles are independent sequences: no grid sync, no per-step launch) with two loader
// warps, an MMA warp and 4 gate-epilogue warpgroups (see lstm_persist_tc_kernel).
// (A cluster-multicast variant of the weight loads was measured slower -- 971k / 824k / 731k
// windows/s at cluster size 1 / 2 / 4 on the c4 shape -- and dropped.)
// bf16 operands / fp32 accumulate / fp32 cell state; gates via tanh.approx (sigmoid = .5*tanh(.5z)+.5).
#include "common.cuh"
#include "ptx.cuh"

using namespace gbptx;

namespace {

constexpr int TILE = 128;
constexpr int UB = 16;             // units per block -> 64 GEMM columns
constexpr int NB_COLS = UB * 4;
constexpr int WG = 128;
constexpr int B_STAGES_HOST = 4;    // == B_STAGES of the step kernel

struct TcLayer {
    int in, u, Kx, Kh, n_blocks, act;
    int64_t w_off, u_off, b_off;   // offsets in the Keras parameter vector
    size_t wp_off;                 // byte offset of this layer's packed blocks
    size_t bias_off;               // float offset of this layer's interleaved biases
    size_t h_off[2];               // byte offsets of the two h buffers
    size_t c_off;                  // byte offset of the cell state
};

struct TcPlan {
    int n_layers, T_in, T_out, L, lookahead, out_act;
    TcLayer ly[GB200_MAX_LAYERS];
    int64_t dense_w_off, dense_b_off, n_params;
    size_t wp_bytes, bias_floats;
    size_t smem_bytes;             // max over layers
    bool eligible;
};

TcPlan make_tc_plan(const gb200_lstm_arch* a) {
    TcPlan p{};
    p.n_layers = a->n_layers; p.T_in = a->n_features; p.T_out = a->n_features_out;
    p.L = a->lookback_window; p.lookahead = a->lookahead; p.out_act = a->out_act;
    int64_t off = 0; int in = a->n_features;
    size_t wp = 0, bf = 0, smem = 0, max_a = 0, max_s = 0;
    for (int l = 0; l < a->n_layers; ++l) {
        TcLayer& y = p.ly[l];
        y.in = in; y.u = a->units[l]; y.act = a->acts[l];
        y.Kx = l == 0 ? gb_round_up(in, 16) : p.ly[l - 1].Kh;
        y.Kh = gb_round_up(y.u, 16);
        y.n_blocks = (y.u + UB - 1) / UB;
        y.w_off = off; off += (int64_t)in * 4 * y.u;
        y.u_off = off; off += (int64_t)y.u * 4 * y.u;
        y.b_off = off; off += 4 * y.u;
        y.wp_off = wp; wp += (size_t)y.n_blocks * (y.Kx + y.Kh) * NB_COLS * 2;
        y.bias_off = bf; bf += (size_t)y.n_blocks * NB_COLS;
        const size_t ab = (size_t)TILE * (y.Kx + y.Kh) * 2, sb = (size_t)(y.Kx > y.Kh ? y.Kx : y.Kh) * NB_COLS * 2;
        if (ab > max_a) max_a = ab;
        if (sb > max_s) max_s = sb;
        smem = max_a + (size_t)B_STAGES_HOST * max_s;
        in = y.u;
    }
    p.dense_w_off = off; off += (int64_t)in * a->n_features_out;
    p.dense_b_off = off; off += a->n_features_out;
    p.n_params = off; p.wp_bytes = wp; p.bias_floats = bf;
    p.smem_bytes = smem + 1024;
    p.eligible = p.smem_bytes <= 227 * 1024 - 1024;
    return p;
}

size_t tc_state_bytes(TcPlan& p, int64_t S, bool assign) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 255) & ~(size_t)255; return r; };
    const size_t wp = take(p.wp_bytes), bias = take(p.bias_floats * 4);
    const size_t xc = take((size_t)(p.ly[0].Kx / 8) * (S + p.L) * 16);
    (void)wp; (void)bias; (void)xc;
    for (int l = 0; l < p.n_layers; ++l) {
        const size_t h0 = take((size_t)p.ly[l].Kh * S * 2), h1 = take((size_t)p.ly[l].Kh * S * 2);
        const size_t c = take((size_t)p.ly[l].n_blocks * UB * S * 4);
        if (assign) { p.ly[l].h_off[0] = h0; p.ly[l].h_off[1] = h1; p.ly[l].c_off = c; }
    }
    return o + 1024;
}

// ---------------------------------------------------------------- packing
// B blocks of one layer: block b, element (n = unit_local*4 + gate, k) at
//   (k/8)*(8*64) + (n/8)*64 + (n%8)*8 + (k%8)   [bf16 elements], k = [x rows | h rows]
__global__ void lstm_pack_w_kernel(TcLayer y, const float* __restrict__ P, uint8_t* __restrict__ wp,
                                   float* __restrict__ bias) {
    const int b = blockIdx.x;
    const int K = y.Kx + y.Kh;
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(wp + y.wp_off + (size_t)b * K * NB_COLS * 2);
    for (int i = threadIdx.x; i < K * NB_COLS; i += blockDim.x) {
        const int k = i / NB_COLS, n = i - k * NB_COLS;
        const int unit = b * UB + (n >> 2), gate = n & 3;
        float w = 0.0f;
        if (unit < y.u) {
            if (k < y.Kx) { if (k < y.in) w = P[y.w_off + (int64_t)k * 4 * y.u + gate * y.u + unit]; }
            else { const int kh = k - y.Kx; if (kh < y.u) w = P[y.u_off + (int64_t)kh * 4 * y.u + gate * y.u + unit]; }
        }
        if (gate != 2) w *= 0.5f;              // sigmoid(z) = 0.5*tanh(0.5 z) + 0.5: the inner 0.5 lives in the weights
        B[(k >> 3) * (8 * 64) + (n >> 3) * 64 + (n & 7) * 8 + (k & 7)] = __float2bfloat16_rn(w);
    }
    for (int n = threadIdx.x; n < NB_COLS; n += blockDim.x) {
        const int unit = b * UB + (n >> 2), gate = n & 3;
        bias[y.bias_off + (size_t)b * NB_COLS + n] = unit < y.u ? P[y.b_off + gate * y.u + unit] * (gate != 2 ? 0.5f : 1.0f) : 0.0f;
    }
}

// Xc[(k/8)][row][k%8] = bf16(x[row0+row][k]*scale[k]+min[k]); rows past the Machine's end are zero
__global__ void lstm_pack_x_kernel(const float* __restrict__ x, int64_t row0, int64_t rows_avail, int64_t rows_chunk,
                                   int T, int Kx, const float* __restrict__ sc, const float* __restrict__ mn,
                                   __nv_bfloat16* __restrict__ xc) {
    const int64_t total = rows_chunk * (Kx / 8);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % rows_chunk; const int c = (int)(i / rows_chunk);
        uint4 pk;
        uint32_t w[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[2];
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = c * 8 + j * 2 + h;
                float t = 0.0f;
                if (r < rows_avail && k < T) { t = x[(row0 + r) * T + k]; if (sc) t = fmaf(t, sc[k], mn[k]); }
                v[h] = t;
            }
            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[0], v[1]);
            w[j] = *reinterpret_cast<uint32_t*>(&t2);
        }
        pk.x = w[0]; pk.y = w[1]; pk.z = w[2]; pk.w = w[3];
        *reinterpret_cast<uint4*>(xc + ((size_t)c * rows_chunk + r) * 8) = pk;
    }
}

template <int ACT>
__device__ __forceinline__ float act_fast(float z) {
    if (ACT == GB200_ACT_TANH) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(z)); return r; }
    if (ACT == GB200_ACT_RELU) return fmaxf(z, 0.0f);
    if (ACT == GB200_ACT_SIGMOID) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(0.5f * z)); return fmaf(0.5f, r, 0.5f); }
    if (ACT == GB200_ACT_ELU) return z > 0.0f ? z : __expf(z) - 1.0f;
    if (ACT == GB200_ACT_SOFTPLUS) return z > 15.0f ? z : __logf(1.0f + __expf(z));
    return z;
}
// z_half = 0.5 * z (the 0.5 is folded into the packed weights / bias of the i, f, o gates)
__device__ __forceinline__ float sigmoid_from_half(float z_half) {
    float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(z_half));
    return fmaf(0.5f, r, 0.5f);
}

struct PLayer {
    int Kx, Kh, n_blocks, act;
    uint32_t wp_off, bias_off;             // byte / float offsets of this layer's packed blocks / biases
    uint64_t h_off[2], c_off;              // byte offsets (from `base`) of the two h buffers and the cell state
};
struct PersistTc {
    int n_layers, L, tiles;
    int64_t xc_rows;                       // rows of Xc
    const uint8_t* xc;                     // bf16 scaled samples, K-chunk-major
    uint8_t* base;                         // state scratch
    const uint8_t* wp; const float* bias;
    uint32_t stage_bytes;                  // weight-ring entry size (max over layers)
    uint32_t a_bytes;                      // A operand region (max over layers)
    PLayer ly[GB200_MAX_LAYERS];
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

constexpr int EPI_WG = 4;          // epilogue warpgroups = TMEM accumulator stages
constexpr int B_STAGES = 4;        // weight ring; one entry = the x rows OR the h rows of a 16-unit block
constexpr int PERSIST_THREADS = EPI_WG * WG + 96;

// gates + cell update + h for one 16-unit block of one step (thread = window = TMEM lane)
template <int ACT>
__device__ __forceinline__ void gate_block(uint32_t tmem_lane, const float* __restrict__ bias, float* __restrict__ cblk,
                                           uint8_t* __restrict__ hdst, bool has_h) {
    #pragma unroll
    for (int half = 0; half < 2; ++half) {
        float cprev[8];
        #pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) cprev[u8] = has_h ? __ldcg(cblk + (half * 8 + u8) * TILE) : 0.0f;
        float hreg[8];
        #pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            float v[16];
            tmem_ld16(tmem_lane + (half * 2 + c2) * 16, v);
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int u8 = c2 * 4 + q, ul = half * 8 + u8;
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + ul * 4));
                const float ig = sigmoid_from_half(v[q * 4 + 0] + bb.x), fg = sigmoid_from_half(v[q * 4 + 1] + bb.y);
                const float gg = act_fast<ACT>(v[q * 4 + 2] + bb.z), og = sigmoid_from_half(v[q * 4 + 3] + bb.w);
                const float cn = fmaf(fg, cprev[u8], ig * gg);
                __stcg(cblk + ul * TILE, cn);
                hreg[u8] = og * act_fast<ACT>(cn);
            }
        }
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(hreg[0], hreg[1]), t1 = __floats2bfloat162_rn(hreg[2], hreg[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(hreg[4], hreg[5]), t3 = __floats2bfloat162_rn(hreg[6], hreg[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(hdst + (size_t)half * 2048) = pk;
    }
}

// PERSISTENT kernel: one CTA owns one 128-window tile for ALL L x n_layers (timestep, layer) steps.
// Window tiles are independent sequences, so there is no grid-wide synchronisation and no
// per-step launch: the CTA walks the steps itself, its h / c state round-trips through L2 only.
//   steps are walked along ANTI-DIAGONALS (l + t = d, l descending): the steps of a diagonal are mutually
//   independent, so while one step's gates finish the next step's operands and MMAs proceed;
//   loader A : per step the A operand [x_t | h_{t-1}] (bulk copies; waits for the two producing steps)
//   loader B : the weight ring, free-running across steps (prefetches the next step's blocks)
//   MMA warp : per block K/16 tcgen05.mma into accumulator stage g%4 (g = global block counter)
//   4 epilogue warpgroups : gates / cell update / h -> HBM in next-step A-operand layout; at the
//              end of a step each publishes its stores to the async proxy and arrives on step_done.
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
lstm_persist_tc_kernel(const __grid_constant__ PersistTc a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty;
    // completion of step (l, t), published twice: done_t[l] is consumed by step (l, t+1), done_l[l] by step (l+1, t)
    __shared__ __align__(8) uint64_t done_t[GB200_MAX_LAYERS], done_l[GB200_MAX_LAYERS];
    __shared__ __align__(8) uint64_t b_full[B_STAGES], b_empty[B_STAGES];
    __shared__ __align__(8) uint64_t t_full[EPI_WG], t_empty[EPI_WG];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, wg = tid / WG, wtid = tid - wg * WG, warp = wtid >> 5;
    const int tile = blockIdx.x;
    uint8_t* A = smem;                                           // [K/8][128][16 B]
    uint8_t* Bbase = smem + a.a_bytes;

    if (tid == 0) {
        mbar_init(&a_full[0], 1); mbar_init(&a_full[1], 1); mbar_init(&a_empty, 1);
        for (int i = 0; i < a.n_layers; ++i) { mbar_init(&done_t[i], EPI_WG); mbar_init(&done_l[i], EPI_WG); }
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < EPI_WG; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(&s_tmem, EPI_WG * NB_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (wg == EPI_WG) {
        if (wtid == 0) {
            // ===================== loader A: the per-step A operand =====================
            int s = 0;
            for (int d = 0; d < a.L + a.n_layers - 1; ++d) {
                const int l_hi = min(a.n_layers - 1, d), l_lo = max(0, d - (a.L - 1));
                for (int l = l_hi; l >= l_lo; --l, ++s) {
                    const int t = d - l;
                    const PLayer& y = a.ly[l];
                    const uint32_t xb = (uint32_t)y.Kx * 256, hb = (uint32_t)y.Kh * 256;
                    if (s > 0) mbar_wait(&a_empty, (s - 1) & 1);           // previous step's MMAs have read A
                    if (t > 0) {                                           // h_{l,t-1}
                        mbar_wait(&done_t[l], (t - 1) & 1);
                        mbar_expect_tx(&a_full[1], hb);
                        bulk_g2s(A + xb, a.base + y.h_off[(t + 1) & 1] + (size_t)tile * hb, hb, &a_full[1]);
                    }
                    mbar_expect_tx(&a_full[0], xb);
                    if (l == 0) {
                        for (int c = 0; c < y.Kx / 8; ++c)
                            bulk_g2s(A + c * 2048, a.xc + ((size_t)c * a.xc_rows + (size_t)tile * TILE + t) * 16, 2048, &a_full[0]);
                    } else {
                        mbar_wait(&done_l[l - 1], t & 1);                  // h_{l-1,t}
                        bulk_g2s(A, a.base + a.ly[l - 1].h_off[t & 1] + (size_t)tile * xb, xb, &a_full[0]);
                    }
                }
            }
        } else if (wtid == 32) {
            // ===================== loader B: free-running weight ring =====================
            int it = 0;
            for (int d = 0; d < a.L + a.n_layers - 1; ++d) {
                const int l_hi = min(a.n_layers - 1, d), l_lo = max(0, d - (a.L - 1));
                for (int l = l_hi; l >= l_lo; --l) {
                    const int parts = (d - l) > 0 ? 2 : 1;
                    const PLayer& y = a.ly[l];
                    const uint32_t blk_bytes = (uint32_t)(y.Kx + y.Kh) * NB_COLS * 2;
                    const int rot = tile % y.n_blocks;
                    for (int b = 0; b < y.n_blocks; ++b) {
                        const int blk = (b + rot) % y.n_blocks;
                        for (int part = 0; part < parts; ++part, ++it) {
                            const int st = it % B_STAGES;
                            if (it >= B_STAGES) mbar_wait(&b_empty[st], ((it / B_STAGES) - 1) & 1);
                            const uint32_t bytes = (uint32_t)(part == 0 ? y.Kx : y.Kh) * NB_COLS * 2;
                            mbar_expect_tx(&b_full[st], bytes);
                            bulk_g2s(Bbase + (size_t)st * a.stage_bytes,
                                     a.wp + y.wp_off + (size_t)blk * blk_bytes + (part ? (size_t)y.Kx * NB_COLS * 2 : 0), bytes, &b_full[st]);
                        }
                    }
                }
            }
        } else if (wtid == 64) {
            // ===================== MMA issuer =====================
            const uint32_t idesc = make_idesc(TILE, NB_COLS);
            const uint32_t a_addr = smem_u32(A);
            int it = 0, g = 0, s = 0, hs = 0;
            for (int d = 0; d < a.L + a.n_layers - 1; ++d) {
                const int l_hi = min(a.n_layers - 1, d), l_lo = max(0, d - (a.L - 1));
                for (int l = l_hi; l >= l_lo; --l, ++s) {
                    const int t = d - l;
                    const int parts = t > 0 ? 2 : 1;
                    const PLayer& y = a.ly[l];
                    mbar_wait(&a_full[0], s & 1);
                    if (t > 0) { mbar_wait(&a_full[1], hs & 1); ++hs; }
                    for (int b = 0; b < y.n_blocks; ++b, ++g) {
                        const int q = g % EPI_WG;
                        if (g >= EPI_WG) mbar_wait(&t_empty[q], ((g / EPI_WG) - 1) & 1);
                        const uint32_t d_tmem = s_tmem + (uint32_t)(q * NB_COLS);
                        for (int part = 0; part < parts; ++part, ++it) {
                            const int st = it % B_STAGES;
                            mbar_wait(&b_full[st], (it / B_STAGES) & 1);
                            tc_fence_after();
                            uint64_t da = make_desc(a_addr + (part ? (uint32_t)y.Kx * 256 : 0), 2048, 128);
                            uint64_t db = make_desc(smem_u32(Bbase + (size_t)st * a.stage_bytes), 1024, 128);
                            const int ksteps = (part ? y.Kh : y.Kx) / 16;
                            umma_bf16(d_tmem, da, db, idesc, part ? 1u : 0u);
                            #pragma unroll 4
                            for (int ks = 1; ks < ksteps; ++ks) {
                                da += (2 * 2048) >> 4; db += (2 * 1024) >> 4;
                                umma_bf16(d_tmem, da, db, idesc, 1u);
                            }
                            umma_commit(&b_empty[st]);     // ring entry reusable once these MMAs retire
                        }
                        umma_commit(&t_full[q]);           // accumulator ready for warpgroup q
                    }
                    umma_commit(&a_empty);                 // the A operand may be overwritten
                }
            }
        }
    } else {
        // ===================== gate epilogue, warpgroup q = wg =====================
        const uint32_t tmem_lane = s_tmem + (uint32_t)(wg * NB_COLS) + ((uint32_t)(warp * 32) << 16);
        int g0 = 0;
        for (int d = 0; d < a.L + a.n_layers - 1; ++d) {
            const int l_hi = min(a.n_layers - 1, d), l_lo = max(0, d - (a.L - 1));
            for (int l = l_hi; l >= l_lo; --l) {
                const int t = d - l;
                const PLayer& y = a.ly[l];
                const int rot = tile % y.n_blocks;
                float* ctile = reinterpret_cast<float*>(a.base + y.c_off) + (size_t)tile * y.n_blocks * UB * TILE + wtid;
                uint8_t* htile = a.base + y.h_off[t & 1] + (size_t)tile * y.Kh * 256 + wtid * 16;
                const float* lbias = a.bias + y.bias_off;
                for (int j = 0; j < y.n_blocks; ++j) {
                    const int g = g0 + j;
                    if (g % EPI_WG != wg) continue;
                    const int b = (j + rot) % y.n_blocks;
                    // only warp 0 polls the accumulator barrier; the other three park on a hardware barrier
                    if (warp == 0) mbar_wait(&t_full[wg], (g / EPI_WG) & 1);
                    named_bar_sync(1 + wg, WG);
                    tc_fence_after();
                    {
                        const float* bias = lbias + (size_t)b * NB_COLS;
                        float* cblk = ctile + b * UB * TILE;
                        uint8_t* hdst = htile + (size_t)(b * 2) * 2048;
                        switch (y.act) {
                            case GB200_ACT_TANH:     gate_block<GB200_ACT_TANH>(tmem_lane, bias, cblk, hdst, t > 0); break;
                            case GB200_ACT_RELU:     gate_block<GB200_ACT_RELU>(tmem_lane, bias, cblk, hdst, t > 0); break;
                            case GB200_ACT_SIGMOID:  gate_block<GB200_ACT_SIGMOID>(tmem_lane, bias, cblk, hdst, t > 0); break;
                            case GB200_ACT_ELU:      gate_block<GB200_ACT_ELU>(tmem_lane, bias, cblk, hdst, t > 0); break;
                            case GB200_ACT_SOFTPLUS: gate_block<GB200_ACT_SOFTPLUS>(tmem_lane, bias, cblk, hdst, t > 0); break;
                            default:                 gate_block<GB200_ACT_LINEAR>(tmem_lane, bias, cblk, hdst, t > 0); break;
                        }
                    }
                    tc_fence_before();
                    named_bar_sync(1 + wg, WG);          // every lane has drained its TMEM reads
                    if (wtid == 0) mbar_arrive(&t_empty[wg]);
                }
                g0 += y.n_blocks;
                // publish this step's h stores to the async proxy (the next step's bulk copies read them)
                fence_proxy_async_all();
                // a warpgroup without a block in some step must not arrive for (l, t) before (l, t-1) has
                // closed, or its early arrival would complete the older phase without a busy warpgroup
                if (t > 0 && warp == 0) mbar_wait(&done_t[l], (t - 1) & 1);
                named_bar_sync(1 + wg, WG);
                // done_l first: once done_t's phase t is complete every warpgroup has also arrived on done_l
                if (wtid == 0) { mbar_arrive(&done_l[l]); mbar_arrive(&done_t[l]); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(s_tmem, EPI_WG * NB_COLS);
}

// yhat = out_act(h_last . Wd + bd) with h_last read from the canonical bf16 tiles
__global__ void lstm_dense_tc_kernel(const uint8_t* __restrict__ h, int Kh, int u, int T_out, int act,
                                     const float* __restrict__ Wd, const float* __restrict__ bd,
                                     int nb, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb * T_out; i += gridDim.x * blockDim.x) {
        const int s = i / T_out, n = i - s * T_out;
        const __nv_bfloat16* tile = reinterpret_cast<const __nv_bfloat16*>(h + (size_t)(s / TILE) * Kh * 256);
        const int r = s % TILE;
        float acc = bd[n];
        for (int k = 0; k < u; ++k)
            acc = fmaf(__bfloat162float(tile[(k >> 3) * 1024 + r * 8 + (k & 7)]), Wd[(size_t)k * T_out + n], acc);
        out[(size_t)s * T_out + n] = gb_act(act, acc);
    }
}

}  // namespace

int64_t gb_lstm_tc_scratch_bytes(const gb200_lstm_arch* arch, int64_t max_windows) {
    TcPlan p = make_tc_plan(arch);
    if (!p.eligible) return 0;
    const int64_t S = (max_windows + TILE - 1) / TILE * TILE;
    return (int64_t)tc_state_bytes(p, S, false);
}

int gb_lstm_predict_tc(const gb200_fleet* f, const gb200_lstm_arch* arch, const float* params,
                       const float* in_scale, const float* in_min, const float* x,
                       const int64_t* out_row_off_host, float* model_out,
                       void* scratch, int64_t scratch_bytes, cudaStream_t stream) {
    TcPlan p = make_tc_plan(arch);
    GB_REQUIRE(p.eligible, "LSTM topology too wide for the tensor-core step kernel (use GB200_PREC_F32)");
    // largest chunk (multiple of 128 windows) that fits the scratch
    int64_t S = TILE;
    while (true) {
        TcPlan q = p;
        if ((int64_t)tc_state_bytes(q, S * 2, false) > scratch_bytes || S * 2 > (1 << 20)) break;
        S *= 2;
    }
    {
        TcPlan q = p;
        GB_REQUIRE((int64_t)tc_state_bytes(q, S, false) <= scratch_bytes, "scratch too small (see gb200_lstm_scratch_bytes)");
        // grow linearly past the last power of two
        while ((int64_t)tc_state_bytes(q, S + TILE, false) <= scratch_bytes && S + TILE <= (1 << 20)) S += TILE;
    }
    tc_state_bytes(p, S, true);
    uint8_t* base = (uint8_t*)scratch;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 255) & ~(size_t)255; return r; };
    uint8_t* wp = base + take(p.wp_bytes);
    float* bias = (float*)(base + take(p.bias_floats * 4));
    __nv_bfloat16* xc = (__nv_bfloat16*)(base + take((size_t)(p.ly[0].Kx / 8) * (S + p.L) * 16));
    const size_t state0 = o;
    // re-derive the state offsets relative to `base` (tc_state_bytes assigned them with the same walk)
    (void)state0;

    for (int m = 0; m < f->n_machines; ++m) {
        const int64_t rows = f->h_row_hi[m] - f->h_row_lo[m];
        if (rows <= 0) continue;
        GB_REQUIRE(p.L < rows, "For KerasLSTMForecast lookback_window must be < size of X (machine %d)", m);
        const int64_t n_win = rows - p.L + 1 - p.lookahead;
        if (n_win <= 0) continue;
        const float* P = params + (size_t)m * p.n_params;
        for (int l = 0; l < p.n_layers; ++l)
            lstm_pack_w_kernel<<<p.ly[l].n_blocks, 256, 0, stream>>>(p.ly[l], P, wp, bias);
        const float* sc = in_scale ? in_scale + (size_t)m * p.T_in : nullptr;
        const float* mn = in_min ? in_min + (size_t)m * p.T_in : nullptr;
        // balanced chunks (each a multiple of 128 windows, <= S): no tiny tail launch
        const int64_t n_chunks = (n_win + S - 1) / S;
        const int64_t Sc = ((n_win + n_chunks - 1) / n_chunks + TILE - 1) / TILE * TILE;
        for (int64_t k0 = 0; k0 < n_win; k0 += Sc) {
            const int nb = (int)((n_win - k0) < Sc ? (n_win - k0) : Sc);
            const int tiles = (nb + TILE - 1) / TILE;
            const int64_t rows_chunk = S + p.L;
            const int64_t row0 = f->h_row_lo[m] + k0;
            const int64_t rows_avail = f->h_row_hi[m] - row0;
            lstm_pack_x_kernel<<<148 * 4, 256, 0, stream>>>(x, row0, rows_avail, rows_chunk, p.T_in, p.ly[0].Kx, sc, mn, xc);
            {
                PersistTc a{};
                a.n_layers = p.n_layers; a.L = p.L; a.tiles = tiles; a.xc_rows = rows_chunk;
                a.xc = (const uint8_t*)xc; a.base = base; a.wp = wp; a.bias = bias;
                uint32_t stage = 0, abytes = 0;
                for (int l = 0; l < p.n_layers; ++l) {
                    const TcLayer& y = p.ly[l];
                    a.ly[l].Kx = y.Kx; a.ly[l].Kh = y.Kh; a.ly[l].n_blocks = y.n_blocks; a.ly[l].act = y.act;
                    a.ly[l].wp_off = (uint32_t)y.wp_off; a.ly[l].bias_off = (uint32_t)y.bias_off;
                    a.ly[l].h_off[0] = y.h_off[0]; a.ly[l].h_off[1] = y.h_off[1]; a.ly[l].c_off = y.c_off;
                    const uint32_t sb = (uint32_t)(y.Kx > y.Kh ? y.Kx : y.Kh) * NB_COLS * 2;
                    if (sb > stage) stage = sb;
                    const uint32_t ab = (uint32_t)TILE * (y.Kx + y.Kh) * 2;
                    if (ab > abytes) abytes = ab;
                }
                a.stage_bytes = stage; a.a_bytes = abytes;
                const size_t smem = (size_t)abytes + (size_t)B_STAGES_HOST * stage;
                GB_CUDA_CHECK(gb_allow_max_smem(lstm_persist_tc_kernel));
                lstm_persist_tc_kernel<<<tiles, PERSIST_THREADS, smem, stream>>>(a);
            }
            const TcLayer& yl = p.ly[p.n_layers - 1];
            int blocks = (nb * p.T_out + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
            lstm_dense_tc_kernel<<<blocks, 256, 0, stream>>>(base + yl.h_off[(p.L - 1) & 1], yl.Kh, yl.u, p.T_out, p.out_act,
                                                              P + p.dense_w_off, P + p.dense_b_off, nb,
                                                              model_out + (out_row_off_host[m] + k0) * p.T_out);
            GB_CUDA_CHECK(cudaGetLastError());
        }
    }
    return GB_OK;
}
