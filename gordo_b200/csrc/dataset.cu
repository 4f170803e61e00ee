// Upstream of X: what `dataset.get_data()` does to the raw tag series before the builder sees a
// matrix (gordo/builder/build_model.py:208-213 -> [3P] gordo-core 0.3.6, gordo_core/time_series.py
// `TimeSeriesDataset.join_timeseries` / `get_data`, gordo_core/filters/rows.py `pandas_filter_rows` / `apply_buffer`;
// the package is not vendored in the reference tree, its algorithm is restated in oracle/dataset.py):
//
//   gb200_resample     every raw series (sorted int64-ns timestamps, float64 values) -> one value per resolution bin:
//                      pandas `series.resample(resolution, label="left").agg(method)`  (mean / min / max / sum / count /
//                      first / last, NaN samples skipped, an empty bin is NaN -- 0 for sum and count)
//   gb200_interpolate  pandas `.interpolate(limit=n)` (linear over bin positions, forward only: leading NaNs stay, at
//                      most `limit` consecutive NaNs behind a value are filled, trailing NaNs take the last value) or
//                      `.fillna(method="ffill", limit=n)`, in place, per series
//   gb200_filter_rows  a row predicate compiled on the host from the dataset's `row_filter` / `known_filter_periods`
//                      strings (pandas `DataFrame.eval` subset) or one of the built-in stages (all columns finite =
//                      the `dropna()` after the inner join; all columns inside (low, high) = the global thresholds),
//                      widened by `buffer_size` rows around every rejected row (`apply_buffer`)
//   gb200_compact_rows the rows a predicate keeps, packed per Machine (what `df[mask]` leaves), with their timestamps
//
// Layout: a Machine's grid is a row-major [bins, tags] float64 matrix; a series is one of its columns (element offset
// + stride), so a warp of the per-series kernels sits on adjacent columns.  Resampling: a CTA owns 256 consecutive
// bins of one series (fewer when the series are few); two CTA-wide searches bound its points, one coalesced pass over
// their timestamps marks where
// every bin starts (a bin edge is crossed by exactly one adjacent pair of points: no atomics), then groups of G lanes
// (G chosen on the host from points per bin) reduce one bin each in a fixed order -- results do not depend on
// scheduling.  Bound: HBM, 16 B per raw point (timestamp + value) + 8 B per bin.
#include "common.cuh"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_BINS = 256;               // bins per CTA

struct ResampleArgs {
    const int64_t* point_off; const int64_t* ts; const double* val;
    const int64_t* bin0; const int64_t* n_bins; const int64_t* out_off; const int64_t* out_stride;
    int64_t step; uint64_t inv_step; int agg; int G; int bins_per_cta; double* out;
};

// first i in [lo, hi) with ts[i] >= key, found by the whole CTA: every round probes blockDim.x evenly spaced samples
// (one coalesced-ish wave of loads instead of one dependent load per halving) and keeps the stretch between the last
// probe below the key and the first one not below it
__device__ __forceinline__ int64_t cta_lower_bound(const int64_t* __restrict__ ts, int64_t lo, int64_t hi, int64_t key,
                                                   int64_t* s_lo, int64_t* s_hi) {
    const int tid = threadIdx.x, nt = blockDim.x;
    while (hi - lo > nt) {
        const int64_t span = hi - lo, chunk = (span + nt - 1) / nt;
        const int64_t p = lo + (int64_t)tid * chunk;                       // probe positions lo, lo+chunk, ...
        const bool below = p < hi && ts[p] < key;
        const int64_t pn = p + chunk;
        const bool next_below = pn < hi && ts[pn] < key;
        __syncthreads();
        if (tid == 0 && !below) { *s_lo = lo; *s_hi = lo; }               // ts[lo] >= key: the answer is lo
        if (below && !next_below) { *s_lo = p + 1; *s_hi = min(pn, hi); }   // exactly one thread sees the crossing
        __syncthreads();
        lo = *s_lo; hi = *s_hi;
    }
    // final stretch (<= nt candidates): each thread tests one, the smallest index that is not below wins
    __syncthreads();
    if (tid == 0) *s_lo = hi;
    __syncthreads();
    const int64_t p = lo + tid;
    if (p < hi && ts[p] >= key && (p == lo || ts[p - 1] < key)) *s_lo = p;
    __syncthreads();
    return *s_lo;
}

template <int AGG>
__global__ void __launch_bounds__(RS_THREADS) resample_kernel(const __grid_constant__ ResampleArgs a) {
    __shared__ int64_t s_start[RS_BINS + 1];
    __shared__ int64_t s_range[4];
    const int s = blockIdx.y, tid = threadIdx.x;
    const int64_t nb = a.n_bins[s];
    const int bpc = a.bins_per_cta;
    const int64_t b0 = (int64_t)blockIdx.x * bpc;
    if (b0 >= nb) return;
    const int nbl = (int)min((int64_t)bpc, nb - b0);
    const int64_t p0 = a.point_off[s], p1 = a.point_off[s + 1];
    const int64_t edge0 = a.bin0[s] + b0 * a.step;
    const int64_t P0 = cta_lower_bound(a.ts, p0, p1, edge0, &s_range[0], &s_range[1]);
    const int64_t P1 = cta_lower_bound(a.ts, P0, p1, edge0 + (int64_t)nbl * a.step, &s_range[2], &s_range[3]);
    for (int q = tid; q <= nbl; q += RS_THREADS) s_start[q] = P1;
    __syncthreads();
    // where every bin starts: point i opens all bins in (bin(i-1), bin(i)].  The bin of a point through a multiply-high
    // by a precomputed reciprocal and one correction step: a 64-bit division (or the float64 conversions of a
    // floating reciprocal: the first version issued 70 % of its slots at 33 % of the HBM peak, profiles/r2l) per point
    // makes this pass instruction-bound.  Four timestamps per thread are loaded before any of them is used (one 2 KB
    // request per warp at a time leaves HBM idle); the previous point's bin comes from the lane below.
    const uint64_t inv = a.inv_step;                   // floor(2^64 / step): the high word of dt * inv is the bin or one less
    auto bin_of = [&](int64_t t) -> int {
        const uint64_t dt = (uint64_t)(t - edge0);
        uint64_t q = __umul64hi(dt, inv);
        if (dt - q * (uint64_t)a.step >= (uint64_t)a.step) ++q;
        return (int)q;
    };
    const int64_t* __restrict__ tsp = a.ts;
    const int lane = tid & 31;
    for (int64_t base = P0; base < P1; base += 4 * RS_THREADS) {
        int64_t t[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = base + k * RS_THREADS + tid;
            t[k] = i < P1 ? tsp[i] : 0;
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = base + k * RS_THREADS + tid;
            const int bl = i < P1 ? bin_of(t[k]) : 0;
            int prev = __shfl_up_sync(0xffffffffu, bl, 1);
            if (i < P1) {
                if (i == P0) prev = -1;
                else if (lane == 0) prev = bin_of(tsp[i - 1]);
                for (int q = prev + 1; q <= bl; ++q) s_start[q] = i;
            }
        }
    }
    __syncthreads();
    const int G = a.G, lane_g = tid % G, grp = tid / G, ngrp = RS_THREADS / G;
    double* out = a.out + a.out_off[s];
    const int64_t stride = a.out_stride[s];
    for (int q0 = 0; q0 < nbl; q0 += ngrp) {          // uniform trip count: the shuffles below need whole warps
        const int q = q0 + grp;
        const bool live = q < nbl;
        const int64_t lo = live ? s_start[q] : 0, hi = live ? s_start[q + 1] : 0;
        // only what the aggregation needs is accumulated (AGG is a template parameter)
        constexpr bool W_SUM = AGG == GB200_AGG_MEAN || AGG == GB200_AGG_SUM, W_CNT = AGG == GB200_AGG_MEAN || AGG == GB200_AGG_COUNT ||
                       AGG == GB200_AGG_MIN || AGG == GB200_AGG_MAX;
        constexpr bool W_MIN = AGG == GB200_AGG_MIN, W_MAX = AGG == GB200_AGG_MAX, W_FIRST = AGG == GB200_AGG_FIRST, W_LAST = AGG == GB200_AGG_LAST;
        double sum = 0.0, mn = INFINITY, mx = -INFINITY, fv = NAN, lv = NAN;
        long long cnt = 0, fi = LLONG_MAX, li = -1;
        auto take = [&](double v, int64_t ik) {
            if (v == v) {
                if (W_SUM) sum += v;
                if (W_CNT) ++cnt;
                if (W_MIN) mn = fmin(mn, v);
                if (W_MAX) mx = fmax(mx, v);
                if (W_FIRST && ik < fi) { fi = ik; fv = v; }
                if (W_LAST && ik > li) { li = ik; lv = v; }
            }
        };
        const double* __restrict__ vp = a.val;
        int64_t i = lo + lane_g;
        for (; i + 3 * (int64_t)G < hi; i += 4 * (int64_t)G) {      // same order as one at a time, four loads in flight
            const double v4[4] = {vp[i], vp[i + G], vp[i + 2 * (int64_t)G], vp[i + 3 * (int64_t)G]};
            #pragma unroll
            for (int k = 0; k < 4; ++k) take(v4[k], i + k * (int64_t)G);
        }
        for (; i < hi; i += G) take(vp[i], i);
        // butterfly inside the group (G is a power of two, groups are aligned inside a warp): both partners of an
        // exchange compute the same sum, so every lane ends with the same bits and the order depends on G only
        for (int o = 1; o < G; o <<= 1) {
            if (W_SUM) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (W_CNT) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (W_MIN) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (W_MAX) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (W_FIRST) {
                const double fv2 = __shfl_xor_sync(0xffffffffu, fv, o); const long long fi2 = __shfl_xor_sync(0xffffffffu, fi, o);
                if (fi2 < fi) { fi = fi2; fv = fv2; }
            }
            if (W_LAST) {
                const double lv2 = __shfl_xor_sync(0xffffffffu, lv, o); const long long li2 = __shfl_xor_sync(0xffffffffu, li, o);
                if (li2 > li) { li = li2; lv = lv2; }
            }
        }
        if (live && lane_g == 0) {
            double r;
            switch (AGG) {
                case GB200_AGG_MEAN:  r = cnt ? sum / (double)cnt : NAN; break;
                case GB200_AGG_MIN:   r = cnt ? mn : NAN; break;
                case GB200_AGG_MAX:   r = cnt ? mx : NAN; break;
                case GB200_AGG_SUM:   r = sum; break;
                case GB200_AGG_COUNT: r = (double)cnt; break;
                case GB200_AGG_FIRST: r = fv; break;
                default:              r = lv; break;
            }
            out[(b0 + q) * stride] = r;
        }
    }
}

// ---------------------------------------------------------------- interpolation / forward fill
struct InterpArgs {
    const int64_t* n_bins; const int64_t* off; const int64_t* stride;
    int n_series; int method; int64_t limit; double* data;
};

__global__ void interpolate_kernel(const __grid_constant__ InterpArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_series) return;
    const int64_t n = a.n_bins[s], st = a.stride[s];
    double* d = a.data + a.off[s];
    const int64_t lim = a.limit < 0 ? INT64_MAX : a.limit;
    int64_t last = -1; double last_v = 0.0;
    for (int64_t b = 0; b < n; ++b) {
        const double v = d[b * st];
        if (v != v) continue;
        if (last >= 0 && b - last > 1) {
            const int64_t gap_end = min(b - 1, last + min(lim, b - last - 1));
            if (a.method == GB200_INTERP_LINEAR) {
                // numpy.interp: slope * (x - x0) + y0, multiply and add rounded separately (no FMA contraction)
                const double slope = __ddiv_rn(__dsub_rn(v, last_v), (double)(b - last));
                for (int64_t p = last + 1; p <= gap_end; ++p)
                    d[p * st] = __dadd_rn(__dmul_rn(slope, (double)(p - last)), last_v);
            } else {
                for (int64_t p = last + 1; p <= gap_end; ++p) d[p * st] = last_v;
            }
        }
        last = b; last_v = v;
    }
    if (last >= 0 && last < n - 1) {            // trailing NaNs: both methods carry the last value forward
        const int64_t gap_end = min(n - 1, last + min(lim, n - 1 - last));
        for (int64_t p = last + 1; p <= gap_end; ++p) d[p * st] = last_v;
    }
}

// ---------------------------------------------------------------- row predicates
constexpr int FP_MAX_OPS = 96, FP_MAX_CONSTS = 48, FP_STACK = 16;
struct FilterArgs {
    const int64_t* lo; const int64_t* hi; const double* data; const int64_t* ts;
    int64_t ts_base; int n_cols, n_ops;
    uint8_t* keep;
    int16_t op[FP_MAX_OPS]; int16_t arg[FP_MAX_OPS];
    double consts[FP_MAX_CONSTS];
};

__global__ void __launch_bounds__(256) filter_rows_kernel(const __grid_constant__ FilterArgs a) {
    const int job = blockIdx.y;
    const int64_t r = a.lo[job] + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.hi[job]) return;
    const double* row = a.data + r * a.n_cols;
    double st[FP_STACK];
    int sp = 0;
    for (int i = 0; i < a.n_ops; ++i) {
        const int op = a.op[i], ar = a.arg[i];
        switch (op) {
            case GB200_OP_CONST: st[sp++] = a.consts[ar]; break;
            case GB200_OP_COL:   st[sp++] = row[ar]; break;
            case GB200_OP_INDEX: st[sp++] = (double)(a.ts[r] - a.ts_base); break;
            case GB200_OP_NEG:   st[sp - 1] = -st[sp - 1]; break;
            case GB200_OP_ABS:   st[sp - 1] = fabs(st[sp - 1]); break;
            case GB200_OP_NOT:   st[sp - 1] = st[sp - 1] != 0.0 ? 0.0 : 1.0; break;
            case GB200_OP_ALL_FINITE: {
                bool ok = true;
                for (int c = 0; c < a.n_cols; ++c) ok = ok && fabs(row[c]) < INFINITY;
                st[sp++] = ok ? 1.0 : 0.0; break;
            }
            case GB200_OP_ALL_NOTNAN: {
                bool ok = true;
                for (int c = 0; c < a.n_cols; ++c) ok = ok && row[c] == row[c];
                st[sp++] = ok ? 1.0 : 0.0; break;
            }
            case GB200_OP_ALL_BETWEEN: {
                const double lo = a.consts[ar], hi = a.consts[ar + 1];
                bool ok = true;
                for (int c = 0; c < a.n_cols; ++c) ok = ok && row[c] > lo && row[c] < hi;
                st[sp++] = ok ? 1.0 : 0.0; break;
            }
            default: {
                const double y = st[--sp], x = st[sp - 1];
                double z;
                switch (op) {
                    case GB200_OP_ADD: z = x + y; break;
                    case GB200_OP_SUB: z = x - y; break;
                    case GB200_OP_MUL: z = x * y; break;
                    case GB200_OP_DIV: z = x / y; break;
                    case GB200_OP_POW: z = pow(x, y); break;
                    case GB200_OP_GT:  z = x > y; break;
                    case GB200_OP_GE:  z = x >= y; break;
                    case GB200_OP_LT:  z = x < y; break;
                    case GB200_OP_LE:  z = x <= y; break;
                    case GB200_OP_EQ:  z = x == y; break;
                    case GB200_OP_NE:  z = x != y; break;
                    case GB200_OP_AND: z = (x != 0.0) && (y != 0.0); break;
                    default:           z = (x != 0.0) || (y != 0.0); break;       // GB200_OP_OR
                }
                st[sp - 1] = z;
            }
        }
    }
    a.keep[r] = sp > 0 && st[sp - 1] != 0.0;
}

// apply_buffer: every rejected row also rejects the `buf` rows on either side of it (inside its job)
__global__ void __launch_bounds__(256) dilate_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                                                     const uint8_t* __restrict__ in, int buf, uint8_t* __restrict__ out) {
    const int job = blockIdx.y;
    const int64_t j0 = lo[job], j1 = hi[job];
    const int64_t r = j0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= j1) return;
    const int64_t w0 = max(j0, r - buf), w1 = min(j1 - 1, r + buf);
    bool ok = true;
    for (int64_t i = w0; i <= w1 && ok; ++i) ok = in[i] != 0;
    out[r] = ok;
}

// ---------------------------------------------------------------- compaction
constexpr int CP_ROWS = 1024;              // rows per chunk = threads per CTA
struct CompactArgs {
    const int32_t* chunk_job; const int64_t* chunk_row0; const int64_t* hi;
    const double* data; const int64_t* ts; const uint8_t* keep;
    int n_cols, n_chunks, n_jobs;
    int64_t* chunk_cnt;                     // [n_chunks + 1]: counts, then exclusive prefix
    double* out; float* out_f32; int64_t* out_ts; int64_t* new_lo; int64_t* new_hi;
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? warp_sum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        warp_sum[lane] = w;
    }
    __syncthreads();
    const int base = warp ? warp_sum[warp - 1] : 0;
    *total = warp_sum[nw - 1];
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(CP_ROWS) compact_count_kernel(const __grid_constant__ CompactArgs a) {
    const int c = blockIdx.x;
    const int64_t r = a.chunk_row0[c] + threadIdx.x;
    const int k = r < a.hi[a.chunk_job[c]] ? (a.keep[r] != 0) : 0;
    int total;
    block_exclusive_scan(k, &total);
    if (threadIdx.x == 0) a.chunk_cnt[c] = total;
}

// one CTA: exclusive prefix over the chunk counts (chunks are job-major), and every job's new row range
__global__ void __launch_bounds__(1024) compact_scan_kernel(const __grid_constant__ CompactArgs a) {
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < a.n_chunks; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const int v = c < a.n_chunks ? (int)a.chunk_cnt[c] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        const long long base = carry;
        if (c < a.n_chunks) {
            const long long pre = base + ex;
            a.chunk_cnt[c] = pre;
            const int job = a.chunk_job[c];
            if (c == 0 || a.chunk_job[c - 1] != job) a.new_lo[job] = pre;
            if (c == a.n_chunks - 1 || a.chunk_job[c + 1] != job) a.new_hi[job] = pre + v;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.chunk_cnt[a.n_chunks] = carry;
}

__global__ void __launch_bounds__(CP_ROWS) compact_scatter_kernel(const __grid_constant__ CompactArgs a) {
    __shared__ int s_rank[CP_ROWS];
    const int c = blockIdx.x;
    const int64_t row0 = a.chunk_row0[c];
    const int64_t r = row0 + threadIdx.x;
    const int nrow = (int)min((int64_t)CP_ROWS, a.hi[a.chunk_job[c]] - row0);
    const int k = threadIdx.x < nrow ? (a.keep[r] != 0) : 0;
    int total;
    const int ex = block_exclusive_scan(k, &total);
    s_rank[threadIdx.x] = k ? ex : -1;
    const int64_t base = a.chunk_cnt[c];
    if (k && a.out_ts) a.out_ts[base + ex] = a.ts[r];
    __syncthreads();
    const int T = a.n_cols;
    for (int e = threadIdx.x; e < nrow * T; e += CP_ROWS) {
        const int rl = e / T, col = e - rl * T;
        const int rk = s_rank[rl];
        if (rk >= 0) {
            const double v = a.data[(row0 + rl) * T + col];
            a.out[(base + rk) * T + col] = v;
            if (a.out_f32) a.out_f32[(base + rk) * T + col] = (float)v;
        }
    }
}

int copy_ranges_to_host(int n_jobs, const int64_t* lo, const int64_t* hi, cudaStream_t stream, std::vector<int64_t>& h) {
    h.resize((size_t)2 * n_jobs);
    GB_CUDA_CHECK(cudaMemcpyAsync(h.data(), lo, sizeof(int64_t) * n_jobs, cudaMemcpyDeviceToHost, stream));
    GB_CUDA_CHECK(cudaMemcpyAsync(h.data() + n_jobs, hi, sizeof(int64_t) * n_jobs, cudaMemcpyDeviceToHost, stream));
    GB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return GB_OK;
}

}  // namespace

int gb_launch_resample(int n_series, const int64_t* point_off, const int64_t* ts, const double* val,
                       const int64_t* bin0, const int64_t* n_bins, const int64_t* out_off, const int64_t* out_stride,
                       int64_t step, int agg, int64_t max_bins, int64_t n_points, int64_t total_bins, double* out,
                       cudaStream_t stream) {
    if (n_series <= 0 || max_bins <= 0) return GB_OK;
    GB_REQUIRE(step > 0, "resample: the resolution must be positive");
    GB_REQUIRE(agg >= GB200_AGG_MEAN && agg <= GB200_AGG_LAST, "resample: unknown aggregation %d", agg);
    GB_REQUIRE(n_series <= 65535, "resample: at most 65535 series per call");
    ResampleArgs a{};
    a.point_off = point_off; a.ts = ts; a.val = val; a.bin0 = bin0; a.n_bins = n_bins;
    a.out_off = out_off; a.out_stride = out_stride; a.step = step; a.agg = agg; a.out = out;
    // lanes per bin: a quarter of the average points per bin, as a power of two in [1, 32]
    const double per_bin = total_bins > 0 ? (double)n_points / (double)total_bins : 1.0;
    int G = 1;
    while (G < 32 && (double)G * 4.0 < per_bin) G <<= 1;
    a.G = G;
    // bins per CTA: 256 when that already makes several waves of CTAs, fewer (down to 16) when the series are few
    int bpc = RS_BINS;
    while (bpc > 16 && (int64_t)n_series * ((max_bins + bpc - 1) / bpc) < 148LL * 8 * 4) bpc >>= 1;
    a.bins_per_cta = bpc;
    dim3 grid((unsigned)((max_bins + bpc - 1) / bpc), (unsigned)n_series);
    a.inv_step = step == 1 ? ~0ull : (uint64_t)((((unsigned __int128)1) << 64) / (unsigned __int128)step);   // floor(2^64 / step)
    switch (agg) {
        case GB200_AGG_MEAN:  resample_kernel<GB200_AGG_MEAN><<<grid, RS_THREADS, 0, stream>>>(a); break;
        case GB200_AGG_MIN:   resample_kernel<GB200_AGG_MIN><<<grid, RS_THREADS, 0, stream>>>(a); break;
        case GB200_AGG_MAX:   resample_kernel<GB200_AGG_MAX><<<grid, RS_THREADS, 0, stream>>>(a); break;
        case GB200_AGG_SUM:   resample_kernel<GB200_AGG_SUM><<<grid, RS_THREADS, 0, stream>>>(a); break;
        case GB200_AGG_COUNT: resample_kernel<GB200_AGG_COUNT><<<grid, RS_THREADS, 0, stream>>>(a); break;
        case GB200_AGG_FIRST: resample_kernel<GB200_AGG_FIRST><<<grid, RS_THREADS, 0, stream>>>(a); break;
        default:              resample_kernel<GB200_AGG_LAST><<<grid, RS_THREADS, 0, stream>>>(a); break;
    }
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_interpolate(int n_series, const int64_t* n_bins, const int64_t* off, const int64_t* stride, int method,
                          int64_t limit, double* data, cudaStream_t stream) {
    if (n_series <= 0 || method == GB200_INTERP_NONE) return GB_OK;
    GB_REQUIRE(method == GB200_INTERP_LINEAR || method == GB200_INTERP_FFILL, "interpolate: unknown method %d", method);
    InterpArgs a{n_bins, off, stride, n_series, method, limit, data};
    interpolate_kernel<<<(n_series + 127) / 128, 128, 0, stream>>>(a);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_filter_rows(int n_jobs, const int64_t* lo, const int64_t* hi, const double* data, int n_cols,
                          const int64_t* ts, int64_t ts_base, const int32_t* ops, const int32_t* args, int n_ops,
                          const double* consts, int n_consts, int buffer_size, uint8_t* keep, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    GB_REQUIRE(n_ops >= 1 && n_ops <= FP_MAX_OPS, "filter_rows: program of %d operations (1..%d supported)", n_ops, FP_MAX_OPS);
    GB_REQUIRE(n_consts >= 0 && n_consts <= FP_MAX_CONSTS, "filter_rows: %d constants (max %d)", n_consts, FP_MAX_CONSTS);
    GB_REQUIRE(buffer_size >= 0, "filter_rows: buffer_size must be >= 0");
    FilterArgs a{};
    a.lo = lo; a.hi = hi; a.data = data; a.ts = ts; a.ts_base = ts_base; a.n_cols = n_cols; a.n_ops = n_ops;
    int depth = 0, max_depth = 0;
    for (int i = 0; i < n_ops; ++i) {
        const int op = ops[i], ar = args[i];
        GB_REQUIRE(op >= 0 && op <= GB200_OP_ALL_BETWEEN, "filter_rows: unknown operation %d", op);
        const bool push = op == GB200_OP_CONST || op == GB200_OP_COL || op == GB200_OP_INDEX || op == GB200_OP_ALL_FINITE ||
                          op == GB200_OP_ALL_NOTNAN || op == GB200_OP_ALL_BETWEEN;
        const bool unary = op == GB200_OP_NEG || op == GB200_OP_ABS || op == GB200_OP_NOT;
        if (op == GB200_OP_CONST) GB_REQUIRE(ar >= 0 && ar < n_consts, "filter_rows: constant %d out of range", ar);
        if (op == GB200_OP_ALL_BETWEEN) GB_REQUIRE(ar >= 0 && ar + 1 < n_consts, "filter_rows: constant %d out of range", ar);
        if (op == GB200_OP_COL) GB_REQUIRE(ar >= 0 && ar < n_cols, "filter_rows: column %d out of range", ar);
        if (op == GB200_OP_INDEX) GB_REQUIRE(ts != nullptr, "filter_rows: the program reads `index` but no timestamps were given");
        if (push) ++depth;
        else if (unary) GB_REQUIRE(depth >= 1, "filter_rows: malformed program (stack underflow at %d)", i);
        else { GB_REQUIRE(depth >= 2, "filter_rows: malformed program (stack underflow at %d)", i); --depth; }
        if (depth > max_depth) max_depth = depth;
        a.op[i] = (int16_t)op; a.arg[i] = (int16_t)ar;
    }
    GB_REQUIRE(depth == 1 && max_depth <= FP_STACK, "filter_rows: malformed program (final depth %d, max depth %d)", depth, max_depth);
    for (int i = 0; i < n_consts; ++i) a.consts[i] = consts[i];
    std::vector<int64_t> h;
    int rc = copy_ranges_to_host(n_jobs, lo, hi, stream, h);
    if (rc) return rc;
    int64_t max_rows = 0, r_end = 0;
    for (int j = 0; j < n_jobs; ++j) { max_rows = std::max(max_rows, h[n_jobs + j] - h[j]); r_end = std::max(r_end, h[n_jobs + j]); }
    if (max_rows <= 0) return GB_OK;
    GB_REQUIRE(n_jobs <= 65535, "filter_rows: at most 65535 jobs per call");
    dim3 grid((unsigned)((max_rows + 255) / 256), (unsigned)n_jobs);
    uint8_t* tmp = nullptr;
    if (buffer_size > 0) {
        GB_CUDA_CHECK(cudaMallocAsync(&tmp, (size_t)r_end, stream));
        a.keep = tmp;
    } else {
        a.keep = keep;
    }
    filter_rows_kernel<<<grid, 256, 0, stream>>>(a);
    if (buffer_size > 0) {
        dilate_kernel<<<grid, 256, 0, stream>>>(lo, hi, tmp, buffer_size, keep);
        cudaFreeAsync(tmp, stream);
    }
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_compact_rows(int n_jobs, const int64_t* lo, const int64_t* hi, const double* data, int n_cols,
                           const int64_t* ts, const uint8_t* keep, double* out, float* out_f32, int64_t* out_ts,
                           int64_t* new_lo, int64_t* new_hi, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    std::vector<int64_t> h;
    int rc = copy_ranges_to_host(n_jobs, lo, hi, stream, h);
    if (rc) return rc;
    std::vector<int32_t> cj; std::vector<int64_t> cr;
    for (int j = 0; j < n_jobs; ++j) {
        GB_REQUIRE(h[n_jobs + j] >= h[j], "compact_rows: job %d has a negative row range", j);
        if (h[n_jobs + j] == h[j]) { cj.push_back(j); cr.push_back(h[j]); continue; }     // keeps new_lo / new_hi defined
        for (int64_t r = h[j]; r < h[n_jobs + j]; r += CP_ROWS) { cj.push_back(j); cr.push_back(r); }
    }
    const int n_chunks = (int)cj.size();
    int32_t* d_cj = nullptr; int64_t* d_cr = nullptr; int64_t* d_cnt = nullptr;
    GB_CUDA_CHECK(cudaMallocAsync(&d_cj, sizeof(int32_t) * n_chunks, stream));
    GB_CUDA_CHECK(cudaMallocAsync(&d_cr, sizeof(int64_t) * n_chunks, stream));
    GB_CUDA_CHECK(cudaMallocAsync(&d_cnt, sizeof(int64_t) * (n_chunks + 1), stream));
    GB_CUDA_CHECK(cudaMemcpyAsync(d_cj, cj.data(), sizeof(int32_t) * n_chunks, cudaMemcpyHostToDevice, stream));
    GB_CUDA_CHECK(cudaMemcpyAsync(d_cr, cr.data(), sizeof(int64_t) * n_chunks, cudaMemcpyHostToDevice, stream));
    CompactArgs a{};
    a.chunk_job = d_cj; a.chunk_row0 = d_cr; a.hi = hi; a.data = data; a.ts = ts; a.keep = keep;
    a.n_cols = n_cols; a.n_chunks = n_chunks; a.n_jobs = n_jobs; a.chunk_cnt = d_cnt;
    a.out = out; a.out_f32 = out_f32; a.out_ts = out_ts; a.new_lo = new_lo; a.new_hi = new_hi;
    compact_count_kernel<<<n_chunks, CP_ROWS, 0, stream>>>(a);
    compact_scan_kernel<<<1, 1024, 0, stream>>>(a);
    compact_scatter_kernel<<<n_chunks, CP_ROWS, 0, stream>>>(a);
    GB_CUDA_CHECK(cudaGetLastError());
    // the host vectors were staged by the pageable copies above; the device arrays are released in stream order
    cudaFreeAsync(d_cj, stream); cudaFreeAsync(d_cr, stream); cudaFreeAsync(d_cnt, stream);
    return GB_OK;
}
