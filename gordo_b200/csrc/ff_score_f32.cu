// ff_score_f32: fp32 fused  MinMax-scale -> Dense stack -> DiffBasedAnomalyDetector columns.
//
// The exact-arithmetic path (GB200_PREC_F32): plain fp32 FMAs in the reference's operation order
// (x*scale+min, z = h.W + b, act, |yhat - y| ...), any topology up to GB200_MAX_LAYERS layers.
// Work unit = (Machine, 128-row tile); one thread owns one row; the Machine's weights are staged
// once per CTA in shared memory (padded to 8-float rows for 128-bit broadcast loads) when they
// fit, otherwise streamed through L1/L2.  Replaces models.py:289-300 + diff.py:336-444.
#include "common.cuh"

namespace {

constexpr int TILE = 128;

struct ScoreArgs {
    gb200_ff_arch arch;
    const int64_t* row_lo; const int64_t* row_hi;
    const int32_t* tile_off;
    int n_machines;
    int tiles_total;
    int64_t n_params;
    const float* params;
    const float* in_scale; const float* in_min; const float* err_scale;
    const float* feat_thr; const float* agg_thr;
    const float* x; const float* y;
    float* model_out; float* tag_scaled; float* tag_unscaled;
    float* total_scaled; float* total_unscaled; float* conf; float* total_conf;
    float* activity;      // optional [rows]: sum_l l1[l] * sum_j |h_l[j]| (the Keras activity-regulariser loss of the row)
    int max_w;            // max layer width
    int smem_w_floats;    // padded weight image size (0 = weights stay in global)
};

// Activations of the scorer: tanh through ex2.approx + rcp.approx (1 - 2 / (1 + e^2z)) and the logistic through one
// ex2 + one rcp: 5-6 instructions instead of libm's 20-30, ABSOLUTE error below 3e-7 (the parity tolerance on yhat is
// 2e-5); relu / linear are exact, elu / softplus keep libm.  Training (ff_fit, lstm) keeps gb_act: there the value also
// feeds act'(h).  (Measured: no change in kernel time -- profiles/README.md r2g -- kept for the instruction count.)
__device__ __forceinline__ float score_act(int code, float z) {
    switch (code) {
        case GB200_ACT_TANH:    return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * z));
        case GB200_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-z));
        default:                return gb_act(code, z);
    }
}

__device__ __forceinline__ int find_machine(const int32_t* tile_off, int n, int tile) {
    int lo = 0, hi = n;                       // largest m with tile_off[m] <= tile
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// NT = threads per CTA = rows per pass; a 128-row tile of the schedule is walked in 128 / NT passes.
// Wide topologies take NT = 64 or 32 so that the weights still fit next to the activation buffers.
template <bool SMEM_W, int NT>
__global__ void __launch_bounds__(NT)
ff_score_f32_kernel(const __grid_constant__ ScoreArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* act0 = smem;                               // [max_w][NT]
    float* act1 = act0 + (size_t)a.max_w * NT;        // [max_w][NT]
    float* wsm  = act1 + (size_t)a.max_w * NT;        // padded weights (SMEM_W)
    // the Machine of the current tile and its row / tile range, refreshed by thread 0 only when the CTA's contiguous
    // tile range crosses into another Machine (a binary search + three dependent global loads PER TILE was ~5 us of
    // pure latency in front of every 128 rows: the bound of the 5-tag fleet)
    __shared__ int s_m, s_tile0, s_tile1;
    __shared__ int64_t s_lo, s_hi;

    const int tid = threadIdx.x;
    const int L = a.arch.n_layers;
    const int T_in = a.arch.widths[0], T_out = a.arch.widths[L];
    int cur_m = -1;

    // contiguous tile range per CTA: a CTA re-stages weights only when it crosses a Machine boundary
    const int per_cta = (a.tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta;
    const int t_end = min(t_begin + per_cta, a.tiles_total);
    if (tid == 0) { s_m = -1; s_tile1 = -1; }
    for (int tile = t_begin; tile < t_end; ++tile) {
        if (tid == 0 && (s_m < 0 || tile >= s_tile1)) {
            int mm = s_m < 0 ? find_machine(a.tile_off, a.n_machines, tile) : s_m + 1;
            while (a.tile_off[mm + 1] <= tile) ++mm;             // Machines without rows own no tile
            s_m = mm; s_tile0 = a.tile_off[mm]; s_tile1 = a.tile_off[mm + 1];
            s_lo = a.row_lo[mm]; s_hi = a.row_hi[mm];
        }
        __syncthreads();
        const int m = s_m;
        const int tile0 = s_tile0;
        const int64_t m_lo = s_lo, m_hi = s_hi;
        const float* P = a.params + (size_t)m * a.n_params;
        if (SMEM_W && m != cur_m) {
            // stage this Machine's weights: layer l -> W padded [in][ldw] then bias [ldw]
            int so = 0; int64_t go = 0;
            for (int l = 0; l < L; ++l) {
                const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
                const int ldw = (wout + 7) & ~7;
                for (int i = tid; i < win * ldw; i += NT) {
                    int k = i / ldw, n = i - k * ldw;
                    wsm[so + i] = n < wout ? P[go + (int64_t)k * wout + n] : 0.0f;
                }
                so += win * ldw; go += (int64_t)win * wout;
                for (int i = tid; i < ldw; i += NT) wsm[so + i] = i < wout ? P[go + i] : 0.0f;
                so += ldw; go += wout;
            }
            cur_m = m;
        }
        for (int sub = 0; sub < TILE / NT; ++sub) {
        const int64_t row = m_lo + (int64_t)(tile - tile0) * TILE + sub * NT + tid;
        const bool valid = row < m_hi;
        if (sub > 0) __syncthreads();         // the previous pass is done with the activation buffers

        // ---- input: MinMaxScaler.transform in fp32
        {
            const float* sc = a.in_scale ? a.in_scale + (size_t)m * T_in : nullptr;
            const float* mn = a.in_min ? a.in_min + (size_t)m * T_in : nullptr;
            const float* xr = a.x + row * T_in;
            for (int k = 0; k < T_in; ++k) {
                float v = valid ? xr[k] : 0.0f;
                if (sc) v = fmaf(v, sc[k], mn[k]);
                act0[k * NT + tid] = v;
            }
        }
        __syncthreads();            // weights staged (and s_m consumed)

        float* hin = act0; float* hout = act1;
        int so = 0; int64_t go = 0;
        float act_l1 = 0.0f;
        for (int l = 0; l < L; ++l) {
            const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
            const int code = a.arch.acts[l];
            const float c1 = a.activity ? a.arch.l1[l] : 0.0f;
            const int ldw = (wout + 7) & ~7;
            const float* W; const float* B;
            if (SMEM_W) { W = wsm + so; B = W + win * ldw; so += win * ldw + ldw; }
            else        { W = P + go;   B = W + (int64_t)win * wout; go += (int64_t)win * wout + wout; }
            for (int n0 = 0; n0 < wout; n0 += 8) {
                float acc[8];
                if (SMEM_W) {
                    const float4 b0 = *reinterpret_cast<const float4*>(B + n0);
                    const float4 b1 = *reinterpret_cast<const float4*>(B + n0 + 4);
                    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
                    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
                    #pragma unroll 4
                    for (int k = 0; k < win; ++k) {
                        const float h = hin[k * NT + tid];
                        const float4 w0 = *reinterpret_cast<const float4*>(W + k * ldw + n0);
                        const float4 w1 = *reinterpret_cast<const float4*>(W + k * ldw + n0 + 4);
                        acc[0] = fmaf(h, w0.x, acc[0]); acc[1] = fmaf(h, w0.y, acc[1]);
                        acc[2] = fmaf(h, w0.z, acc[2]); acc[3] = fmaf(h, w0.w, acc[3]);
                        acc[4] = fmaf(h, w1.x, acc[4]); acc[5] = fmaf(h, w1.y, acc[5]);
                        acc[6] = fmaf(h, w1.z, acc[6]); acc[7] = fmaf(h, w1.w, acc[7]);
                    }
                } else {
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = (n0 + j < wout) ? __ldg(B + n0 + j) : 0.0f;
                    for (int k = 0; k < win; ++k) {
                        const float h = hin[k * NT + tid];
                        const float* wr = W + (int64_t)k * wout + n0;
                        #pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (n0 + j < wout) acc[j] = fmaf(h, __ldg(wr + j), acc[j]);
                    }
                }
                #pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (n0 + j < wout) {
                        const float h = score_act(code, acc[j]);
                        hout[(n0 + j) * NT + tid] = h;
                        if (c1 != 0.0f) act_l1 = fmaf(c1, fabsf(h), act_l1);
                    }
            }
            float* t = hin; hin = hout; hout = t;     // each thread reads only its own column
        }

        // ---- epilogue: DiffBasedAnomalyDetector columns (diff.py:350-444).  A thread owns a row, but rows
        // are only 4*T bytes apart: writing them straight from the thread would touch every 32-byte sector
        // T times with 4 bytes.  Each column is staged row-major in the free activation buffer and leaves
        // as one contiguous, fully coalesced run (the tile's rows are consecutive in every output array).
        {
            const int64_t row0 = m_lo + (int64_t)(tile - tile0) * TILE + sub * NT;
            const int nvalid = (int)max((int64_t)0, min((int64_t)NT, m_hi - row0));
            const float* yr = (a.y ? a.y : a.x) + row * T_out;
            const float* es = a.err_scale ? a.err_scale + (size_t)m * T_out : nullptr;
            const float* ft = a.feat_thr ? a.feat_thr + (size_t)m * T_out : nullptr;
            float* stage = hout;                                    // [NT][T_out] row-major
            __syncthreads();                                        // hout was the last layer's input: every thread is done reading it
            auto flush = [&](float* dst) {
                __syncthreads();
                float* o = dst + row0 * T_out;
                for (int i = tid; i < nvalid * T_out; i += NT) o[i] = stage[i];
                __syncthreads();
            };
            if (a.model_out) {
                for (int j = 0; j < T_out; ++j) stage[tid * T_out + j] = hin[j * NT + tid];
                flush(a.model_out);
            }
            float ss = 0.0f, su = 0.0f;
            for (int j = 0; j < T_out; ++j) {
                const float d = valid ? fabsf(hin[j * NT + tid] - yr[j]) : 0.0f;
                const float sc = es ? d * fabsf(es[j]) : d;
                su = fmaf(d, d, su); ss = fmaf(sc, sc, ss);
                stage[tid * T_out + j] = d;
            }
            if (a.tag_unscaled) flush(a.tag_unscaled); else __syncthreads();
            if (a.tag_scaled) {
                for (int j = 0; j < T_out; ++j) { const float d = stage[tid * T_out + j]; stage[tid * T_out + j] = es ? d * fabsf(es[j]) : d; }
                flush(a.tag_scaled);
                if (a.conf && ft)       // back to the unscaled error for the confidence column
                    for (int j = 0; j < T_out; ++j) stage[tid * T_out + j] = valid ? fabsf(hin[j * NT + tid] - yr[j]) : 0.0f;
            }
            if (a.conf && ft) {
                for (int j = 0; j < T_out; ++j) stage[tid * T_out + j] = stage[tid * T_out + j] / ft[j];
                flush(a.conf);
            }
            if (valid) {
                const float ts = ss / (float)T_out, tu = su / (float)T_out;
                if (a.total_scaled) a.total_scaled[row] = ts;
                if (a.total_unscaled) a.total_unscaled[row] = tu;
                if (a.total_conf && a.agg_thr) a.total_conf[row] = ts / a.agg_thr[m];
                if (a.activity) a.activity[row] = act_l1;
            }
        }
        }
        __syncthreads();            // act buffers / s_m reused by the next tile
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// ff_score_small_kernel: the same fp32 scorer for TINY topologies (every width <= 8: the c5 fleet, 10 000 Machines of 5
// tags).  The general kernel above spends ~2 200 thread-instructions per row at 5 tags (ncu: issue slots 80 % busy, the
// kernel is bound by its own loop / dispatch bookkeeping: 105 useful FMAs per row).  Here a thread keeps its row's
// activations in REGISTERS (h[8]), a layer is an unrolled k-loop over 8-wide zero-padded weight rows in shared memory
// (two 128-bit broadcast loads + 8 FMAs per k, left at k == fan-in), the activation is dispatched once per layer, the
// sample tile comes in with one coalesced pass and all four output matrices leave through one staging buffer with two
// block barriers per tile.  Same arithmetic order per output as the general kernel (bias first, k ascending).
template <int ACT>
__device__ __forceinline__ float small_act(float z) {
    if (ACT == GB200_ACT_TANH) return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * z));
    if (ACT == GB200_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-z));
    return gb_act(ACT, z);
}
template <int ACT>
__device__ __forceinline__ void small_layer_act(float (&h)[8], const float (&o)[8], int wout) {
    #pragma unroll
    for (int j = 0; j < 8; ++j) { if (j >= wout) break; h[j] = small_act<ACT>(o[j]); }
}

constexpr int SM_NT = 128;
constexpr int SM_LW = 72;                     // floats per layer in shared memory: W [8][8] + bias [8]

__global__ void __launch_bounds__(SM_NT)
ff_score_small_kernel(const __grid_constant__ ScoreArgs a) {
    __shared__ __align__(16) float wsm[GB200_MAX_LAYERS * SM_LW];
    __shared__ __align__(16) float vec[4][8];                 // in_scale, in_min, |err_scale|, feat_thr
    __shared__ __align__(16) float xt[SM_NT * 8];             // sample tile, [row][T_in] compact
    __shared__ __align__(16) float yt[SM_NT * 8];             // target tile when y is a separate matrix
    __shared__ __align__(16) float stage[4][SM_NT * 8];       // model-output, unscaled, scaled, confidence: [row][T_out]
    __shared__ int s_m, s_tile0, s_tile1;
    __shared__ int64_t s_lo, s_hi;
    const int tid = threadIdx.x;
    const int L = a.arch.n_layers;
    const int T_in = a.arch.widths[0], T_out = a.arch.widths[L];
    int cur_m = -1;
    const int per_cta = (a.tiles_total + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per_cta;
    const int t_end = min(t_begin + per_cta, a.tiles_total);
    if (tid == 0) { s_m = -1; s_tile1 = -1; }
    for (int tile = t_begin; tile < t_end; ++tile) {
        if (tid == 0 && (s_m < 0 || tile >= s_tile1)) {          // Machine lookup only at Machine boundaries
            int mm = s_m < 0 ? find_machine(a.tile_off, a.n_machines, tile) : s_m + 1;
            while (a.tile_off[mm + 1] <= tile) ++mm;
            s_m = mm; s_tile0 = a.tile_off[mm]; s_tile1 = a.tile_off[mm + 1];
            s_lo = a.row_lo[mm]; s_hi = a.row_hi[mm];
        }
        __syncthreads();
        const int m = s_m;
        const int64_t row0 = s_lo + (int64_t)(tile - s_tile0) * TILE;
        const int nv = (int)max((int64_t)0, min((int64_t)SM_NT, s_hi - row0));
        if (m != cur_m) {
            const float* P = a.params + (size_t)m * a.n_params;
            int64_t go = 0;
            for (int l = 0; l < L; ++l) {
                const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
                for (int i = tid; i < SM_LW; i += SM_NT) {
                    const int k = i >> 3, n = i & 7;
                    float v = 0.0f;
                    if (n < wout) { if (k < win) v = P[go + (int64_t)k * wout + n]; else if (k == 8) v = P[go + (int64_t)win * wout + n]; }
                    wsm[l * SM_LW + i] = v;
                }
                go += (int64_t)win * wout + wout;
            }
            if (tid < 8) {
                vec[0][tid] = (tid < T_in && a.in_scale) ? a.in_scale[(size_t)m * T_in + tid] : 1.0f;
                vec[1][tid] = (tid < T_in && a.in_min) ? a.in_min[(size_t)m * T_in + tid] : 0.0f;
                vec[2][tid] = (tid < T_out && a.err_scale) ? fabsf(a.err_scale[(size_t)m * T_out + tid]) : 1.0f;
                vec[3][tid] = (tid < T_out && a.feat_thr) ? a.feat_thr[(size_t)m * T_out + tid] : 1.0f;
            }
            cur_m = m;
        }
        {   // the tile's samples (and targets): one coalesced pass
            const float* src = a.x + row0 * T_in;
            for (int i = tid; i < nv * T_in; i += SM_NT) xt[i] = src[i];
            if (a.y) { const float* ys = a.y + row0 * T_out; for (int i = tid; i < nv * T_out; i += SM_NT) yt[i] = ys[i]; }
        }
        __syncthreads();
        const bool valid = tid < nv;
        float h[8], yv[8];
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            h[k] = (valid && k < T_in) ? (a.in_scale ? fmaf(xt[tid * T_in + k], vec[0][k], vec[1][k]) : xt[tid * T_in + k]) : 0.0f;
            yv[k] = (valid && k < T_out) ? (a.y ? yt[tid * T_out + k] : xt[tid * T_out + k]) : 0.0f;
        }
        float act_l1 = 0.0f;
        for (int l = 0; l < L; ++l) {
            const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
            const float* W = wsm + l * SM_LW;
            float o[8];
            {
                const float4 b0 = *reinterpret_cast<const float4*>(W + 64), b1 = *reinterpret_cast<const float4*>(W + 68);
                o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y; o[6] = b1.z; o[7] = b1.w;
            }
            #pragma unroll
            for (int k = 0; k < 8; ++k) {           // (fan-in / fan-out as template parameters measured no faster)
                if (k >= win) break;
                const float4 w0 = *reinterpret_cast<const float4*>(W + k * 8), w1 = *reinterpret_cast<const float4*>(W + k * 8 + 4);
                const float hk = h[k];
                o[0] = fmaf(hk, w0.x, o[0]); o[1] = fmaf(hk, w0.y, o[1]); o[2] = fmaf(hk, w0.z, o[2]); o[3] = fmaf(hk, w0.w, o[3]);
                o[4] = fmaf(hk, w1.x, o[4]); o[5] = fmaf(hk, w1.y, o[5]); o[6] = fmaf(hk, w1.z, o[6]); o[7] = fmaf(hk, w1.w, o[7]);
            }
            switch (a.arch.acts[l]) {
                case GB200_ACT_TANH:     small_layer_act<GB200_ACT_TANH>(h, o, wout); break;
                case GB200_ACT_RELU:     small_layer_act<GB200_ACT_RELU>(h, o, wout); break;
                case GB200_ACT_SIGMOID:  small_layer_act<GB200_ACT_SIGMOID>(h, o, wout); break;
                case GB200_ACT_ELU:      small_layer_act<GB200_ACT_ELU>(h, o, wout); break;
                case GB200_ACT_SOFTPLUS: small_layer_act<GB200_ACT_SOFTPLUS>(h, o, wout); break;
                default:                 small_layer_act<GB200_ACT_LINEAR>(h, o, wout); break;
            }
            if (a.activity && a.arch.l1[l] != 0.0f) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) { if (j >= wout) break; act_l1 = fmaf(a.arch.l1[l], fabsf(h[j]), act_l1); }
            }
        }
        // ---- DiffBasedAnomalyDetector columns (diff.py:350-444), staged row-major and written as contiguous runs
        float ss = 0.0f, su = 0.0f;
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= T_out) break;
            const float d = valid ? fabsf(h[j] - yv[j]) : 0.0f;
            const float sc = d * vec[2][j];
            su = fmaf(d, d, su); ss = fmaf(sc, sc, ss);
            const int o = tid * T_out + j;
            stage[0][o] = h[j]; stage[1][o] = d; stage[2][o] = a.err_scale ? sc : d; stage[3][o] = d / vec[3][j];
        }
        if (valid) {
            const int64_t row = row0 + tid;
            const float ts = ss / (float)T_out, tu = su / (float)T_out;
            if (a.total_scaled) a.total_scaled[row] = ts;
            if (a.total_unscaled) a.total_unscaled[row] = tu;
            if (a.total_conf && a.agg_thr) a.total_conf[row] = ts / a.agg_thr[m];
            if (a.activity) a.activity[row] = act_l1;
        }
        __syncthreads();
        {
            const int cnt = nv * T_out;
            const int64_t base = row0 * T_out;
            if (a.model_out)    for (int i = tid; i < cnt; i += SM_NT) a.model_out[base + i] = stage[0][i];
            if (a.tag_unscaled) for (int i = tid; i < cnt; i += SM_NT) a.tag_unscaled[base + i] = stage[1][i];
            if (a.tag_scaled)   for (int i = tid; i < cnt; i += SM_NT) a.tag_scaled[base + i] = stage[2][i];
            if (a.conf && a.feat_thr) for (int i = tid; i < cnt; i += SM_NT) a.conf[base + i] = stage[3][i];
        }
        __syncthreads();            // the tile buffers are reused by the next tile
    }
}

}  // namespace

int gb_launch_ff_score_f32(const gb200_fleet* f, const gb200_ff_arch* arch, const float* params,
                           const float* in_scale, const float* in_min, const float* err_scale,
                           const float* feat_thr, const float* agg_thr, const float* x, const float* y,
                           float* model_out, float* tag_scaled, float* tag_unscaled,
                           float* total_scaled, float* total_unscaled, float* conf, float* total_conf,
                           float* activity, cudaStream_t stream) {
    ScoreArgs a{};
    a.activity = activity;
    a.arch = *arch;
    a.row_lo = f->d_row_lo; a.row_hi = f->d_row_hi; a.tile_off = f->d_tile_off;
    a.n_machines = f->n_machines; a.tiles_total = f->tiles_total;
    a.n_params = gb200_ff_param_count(arch);
    a.params = params; a.in_scale = in_scale; a.in_min = in_min; a.err_scale = err_scale;
    a.feat_thr = feat_thr; a.agg_thr = agg_thr; a.x = x; a.y = y;
    a.model_out = model_out; a.tag_scaled = tag_scaled; a.tag_unscaled = tag_unscaled;
    a.total_scaled = total_scaled; a.total_unscaled = total_unscaled;
    a.conf = conf; a.total_conf = total_conf;
    int max_w = 0, wfloats = 0;
    for (int l = 0; l <= arch->n_layers; ++l) max_w = arch->widths[l] > max_w ? arch->widths[l] : max_w;
    for (int l = 0; l < arch->n_layers; ++l) {
        int ldw = gb_round_up(arch->widths[l + 1], 8);
        wfloats += arch->widths[l] * ldw + ldw;
    }
    a.max_w = max_w;
    if (f->tiles_total == 0) return GB_OK;
    if (max_w <= 8) {               // tiny topologies: the register-resident kernel
        int per_sm = 1;
        GB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ff_score_small_kernel, SM_NT, 0));
        if (per_sm < 1) per_sm = 1;
        int grid = f->sm_count * per_sm;
        if (grid > f->tiles_total) grid = f->tiles_total;
        ff_score_small_kernel<<<grid, SM_NT, 0, stream>>>(a);
        GB_CUDA_CHECK(cudaGetLastError());
        return GB_OK;
    }
    const size_t smem_cap = 227 * 1024 - 64;
    // rows per pass: 128 when the padded weights fit beside two [max_w][128] activation buffers, else 64 / 32
    int nt = 128;
    while (nt > 32 && (size_t)2 * max_w * nt * sizeof(float) + (size_t)wfloats * sizeof(float) > smem_cap) nt >>= 1;
    const size_t act_bytes = (size_t)2 * max_w * nt * sizeof(float);
    GB_REQUIRE(act_bytes <= smem_cap, "ff_score_f32: layer width %d too large for the fp32 kernel", max_w);
    const bool smem_w = act_bytes + (size_t)wfloats * sizeof(float) <= smem_cap;
    if (!smem_w) nt = 128;                    // weights stay in global memory anyway: widest pass
    const size_t act_b = (size_t)2 * max_w * nt * sizeof(float);
    a.smem_w_floats = smem_w ? wfloats : 0;
    const size_t smem = act_b + (smem_w ? (size_t)wfloats * sizeof(float) : 0);
    void (*kern)(ScoreArgs) = !smem_w ? ff_score_f32_kernel<false, 128>
                            : nt == 128 ? ff_score_f32_kernel<true, 128>
                            : nt == 64 ? ff_score_f32_kernel<true, 64> : ff_score_f32_kernel<true, 32>;
    GB_CUDA_CHECK(gb_allow_max_smem(kern));
    int per_sm = 1;
    GB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = f->sm_count * per_sm;
    if (grid > f->tiles_total) grid = f->tiles_total;
    kern<<<grid, nt, smem, stream>>>(a);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}
