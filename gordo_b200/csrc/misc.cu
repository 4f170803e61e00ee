// Small HBM-bound kernels around the dense stacks: MinMaxScaler.fit, rolling-min-max thresholds,
// scoring of precomputed model output.  All are column reductions over row ranges of row-major
// [rows, T] matrices: threads are laid out [row-lane][tag] so that a warp reads consecutive
// addresses, partials are combined in shared memory and published with one atomic per tag.
#include "common.cuh"

namespace {

// float min / max through integer atomics: the branch is on the SIGN BIT (not on v >= 0, which is
// true for -0.0f and would send its bit pattern 0x80000000 = INT_MIN down the signed path)
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
    v += 0.0f;                                   // canonicalise -0.0 to +0.0
    if (__float_as_int(v) >= 0) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else                        atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
    v += 0.0f;
    if (__float_as_int(v) >= 0) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else                        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

constexpr int RED_THREADS = 256;
constexpr int ROWS_PER_CHUNK = 4096;

__global__ void fill2_kernel(float* a, float va, float* b, float vb, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = va; if (b) b[i] = vb; }
}

// scale <- running min, min_ <- running max (temporarily), finalised below
__global__ void __launch_bounds__(RED_THREADS)
minmax_partial_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                      const float* __restrict__ x, int T, float* run_min, float* run_max) {
    const int job = blockIdx.y;
    const int64_t r0 = lo[job] + (int64_t)blockIdx.x * ROWS_PER_CHUNK;
    const int64_t r1 = min(r0 + (int64_t)ROWS_PER_CHUNK, hi[job]);
    if (r0 >= r1) return;
    extern __shared__ float sm[];            // [2][R][T]
    for (int c0 = 0; c0 < T; c0 += RED_THREADS) {       // T > 256: column blocks
        const int Tc = min(T - c0, RED_THREADS);
        const int R = RED_THREADS / Tc;
        const int lane = threadIdx.x / Tc, tag = threadIdx.x - lane * Tc;
        float mn = INFINITY, mx = -INFINITY;
        if (lane < R) {
            for (int64_t r = r0 + lane; r < r1; r += R) {
                const float v = x[r * T + c0 + tag];
                mn = fminf(mn, v); mx = fmaxf(mx, v);   // NaN-ignoring, as np.nanmin / nanmax
            }
            sm[lane * Tc + tag] = mn; sm[R * Tc + lane * Tc + tag] = mx;
        }
        __syncthreads();
        if (threadIdx.x < Tc) {
            for (int l = 1; l < R; ++l) {
                mn = fminf(mn, sm[l * Tc + threadIdx.x]);
                mx = fmaxf(mx, sm[R * Tc + l * Tc + threadIdx.x]);
            }
            if (mn != INFINITY) atomic_min_f(run_min + (size_t)job * T + c0 + threadIdx.x, mn);
            if (mx != -INFINITY) atomic_max_f(run_max + (size_t)job * T + c0 + threadIdx.x, mx);
        }
        __syncthreads();
    }
}

__global__ void minmax_finalize_kernel(float* scale, float* min_, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // sklearn: data_range = max - min, _handle_zeros_in_scale -> 1 where < 10*eps; computed in
    // double so that the fp32 result is the correctly rounded float64 value
    const double mn = scale[i], mx = min_[i];
    double rng = mx - mn;
    if (!(rng >= 10.0 * 2.220446049250313e-16)) rng = 1.0;
    const double s = 1.0 / rng;
    scale[i] = (float)s;
    min_[i] = (float)(0.0 - mn * s);
}

__global__ void __launch_bounds__(RED_THREADS)
rolling_partial_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi,
                       const float* __restrict__ v, int C, int window, float* out) {
    const int job = blockIdx.y;
    const int64_t j0 = lo[job];
    const int64_t r0 = j0 + (int64_t)blockIdx.x * ROWS_PER_CHUNK;
    const int64_t r1 = min(r0 + (int64_t)ROWS_PER_CHUNK, hi[job]);
    if (r0 >= r1) return;
    extern __shared__ float sm[];
    for (int c0 = 0; c0 < C; c0 += RED_THREADS) {
        const int Cc = min(C - c0, RED_THREADS);
        const int R = RED_THREADS / Cc;
        const int lane = threadIdx.x / Cc, col = threadIdx.x - lane * Cc;
        float best = -INFINITY;
        if (lane < R) {
            for (int64_t t = r0 + lane; t < r1; t += R) {
                if (t - j0 < window - 1) continue;          // pandas: first window-1 rows are NaN
                float m = INFINITY; bool bad = false;
                for (int w = 0; w < window; ++w) {
                    const float e = v[(t - w) * C + c0 + col];
                    bad |= isnan(e); m = fminf(m, e);
                }
                if (!bad) best = fmaxf(best, m);            // .max() skips NaN windows
            }
            sm[lane * Cc + col] = best;
        }
        __syncthreads();
        if (threadIdx.x < Cc) {
            for (int l = 1; l < R; ++l) best = fmaxf(best, sm[l * Cc + threadIdx.x]);
            if (best != -INFINITY) atomic_max_f(out + (size_t)job * C + c0 + threadIdx.x, best);
        }
        __syncthreads();
    }
}

__global__ void neg_inf_to_nan_kernel(float* out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && out[i] == -INFINITY) out[i] = NAN;
}

__device__ __forceinline__ int find_seg(const int64_t* off, int n, int64_t r) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= r) lo = mid; else hi = mid; }
    return lo;
}

// one warp per output row
__global__ void __launch_bounds__(256)
score_outputs_kernel(int M, const int64_t* __restrict__ out_off, const int64_t* __restrict__ y_off, int T,
                     const float* __restrict__ mo, const float* __restrict__ y,
                     const float* __restrict__ es, const float* __restrict__ ft, const float* __restrict__ at,
                     float* tag_scaled, float* tag_unscaled, float* total_scaled, float* total_unscaled,
                     float* conf, float* total_conf, int64_t rows_total) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < rows_total; r += warps) {
        const int m = find_seg(out_off, M, r);
        const int64_t yr = y_off[m] + (r - out_off[m]);
        float ss = 0.0f, su = 0.0f;
        for (int j = lane; j < T; j += 32) {
            const float d = fabsf(mo[r * T + j] - y[yr * T + j]);
            const float s = es ? d * fabsf(es[(size_t)m * T + j]) : d;
            su = fmaf(d, d, su); ss = fmaf(s, s, ss);
            if (tag_unscaled) tag_unscaled[r * T + j] = d;
            if (tag_scaled) tag_scaled[r * T + j] = s;
            if (conf && ft) conf[r * T + j] = d / ft[(size_t)m * T + j];
        }
        #pragma unroll
        for (int o = 16; o; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o); su += __shfl_xor_sync(0xffffffffu, su, o);
        }
        if (lane == 0) {
            const float ts = ss / (float)T;
            if (total_scaled) total_scaled[r] = ts;
            if (total_unscaled) total_unscaled[r] = su / (float)T;
            if (total_conf && at) total_conf[r] = ts / at[m];
        }
    }
}

// per (job, tag): sum y, sum y^2, sum e, sum e^2, sum |e| with e = y - yhat, in double (the variance
// terms of r2 / explained variance cancel catastrophically in float32)
__global__ void __launch_bounds__(RED_THREADS)
cv_sums_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi, const float* __restrict__ y,
               const float* __restrict__ yhat, int T, double* __restrict__ out) {
    const int job = blockIdx.y;
    const int64_t r0 = lo[job] + (int64_t)blockIdx.x * ROWS_PER_CHUNK;
    const int64_t r1 = min(r0 + (int64_t)ROWS_PER_CHUNK, hi[job]);
    if (r0 >= r1) return;
    extern __shared__ double smd[];          // [5][R][Tc]
    for (int c0 = 0; c0 < T; c0 += RED_THREADS) {
        const int Tc = min(T - c0, RED_THREADS);
        const int R = RED_THREADS / Tc;
        const int lane = threadIdx.x / Tc, tag = threadIdx.x - lane * Tc;
        double acc[5] = {0, 0, 0, 0, 0};
        if (lane < R) {
            for (int64_t r = r0 + lane; r < r1; r += R) {
                const double yv = y[r * T + c0 + tag], e = yv - (double)yhat[r * T + c0 + tag];
                acc[0] += yv; acc[1] += yv * yv; acc[2] += e; acc[3] += e * e; acc[4] += fabs(e);
            }
            #pragma unroll
            for (int q = 0; q < 5; ++q) smd[(q * R + lane) * Tc + tag] = acc[q];
        }
        __syncthreads();
        if (threadIdx.x < Tc) {
            #pragma unroll
            for (int q = 0; q < 5; ++q) {
                double t = 0.0;
                for (int l = 0; l < R; ++l) t += smd[(q * R + l) * Tc + threadIdx.x];
                atomicAdd(out + ((size_t)job * 5 + q) * T + c0 + threadIdx.x, t);
            }
        }
        __syncthreads();
    }
}

int max_chunks_host(int n_jobs, const int64_t* lo, const int64_t* hi, cudaStream_t stream, int64_t* out) {
    // row ranges live on the device; the grid must cover the longest job
    int64_t* hl = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)n_jobs);
    if (!hl) { gb_set_error("host malloc failed"); return GB_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(hl, lo, sizeof(int64_t) * n_jobs, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hl + n_jobs, hi, sizeof(int64_t) * n_jobs, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { free(hl); gb_set_error("row range copy: %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
    int64_t mx = 0;
    for (int i = 0; i < n_jobs; ++i) { int64_t n = hl[n_jobs + i] - hl[i]; if (n > mx) mx = n; }
    free(hl);
    *out = mx;
    return GB_OK;
}

}  // namespace

int gb_launch_minmax_fit(int n_jobs, const int64_t* lo, const int64_t* hi, const float* x, int n_tags,
                         float* scale, float* min_, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    int64_t max_rows = 0;
    int rc = max_chunks_host(n_jobs, lo, hi, stream, &max_rows);
    if (rc) return rc;
    const int64_t n = (int64_t)n_jobs * n_tags;
    fill2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scale, INFINITY, min_, -INFINITY, n);
    const int chunks = (int)((max_rows + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK);
    if (chunks > 0) {
        dim3 grid(chunks, n_jobs);
        minmax_partial_kernel<<<grid, RED_THREADS, 2 * RED_THREADS * sizeof(float), stream>>>(lo, hi, x, n_tags, scale, min_);
    }
    minmax_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scale, min_, n);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_rolling_min_max(int n_jobs, const int64_t* lo, const int64_t* hi, const float* v,
                              int n_cols, int window, float* out, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    GB_REQUIRE(window >= 1, "rolling_min_max: window must be >= 1");
    int64_t max_rows = 0;
    int rc = max_chunks_host(n_jobs, lo, hi, stream, &max_rows);
    if (rc) return rc;
    const int64_t n = (int64_t)n_jobs * n_cols;
    fill2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, -INFINITY, nullptr, 0.0f, n);
    const int chunks = (int)((max_rows + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK);
    if (chunks > 0) {
        dim3 grid(chunks, n_jobs);
        rolling_partial_kernel<<<grid, RED_THREADS, RED_THREADS * sizeof(float), stream>>>(lo, hi, v, n_cols, window, out);
    }
    neg_inf_to_nan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, n);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_score_outputs(int n_machines, const int64_t* out_row_off, const int64_t* y_row_off,
                            int n_tags, const float* model_out, const float* y, const float* err_scale,
                            const float* feat_thr, const float* agg_thr, float* tag_scaled,
                            float* tag_unscaled, float* total_scaled, float* total_unscaled,
                            float* conf, float* total_conf, cudaStream_t stream) {
    if (n_machines <= 0) return GB_OK;
    int64_t rows_total = 0;
    GB_CUDA_CHECK(cudaMemcpyAsync(&rows_total, out_row_off + n_machines, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    GB_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (rows_total <= 0) return GB_OK;
    int64_t blocks = (rows_total + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    score_outputs_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n_machines, out_row_off, y_row_off, n_tags,
        model_out, y, err_scale, feat_thr, agg_thr, tag_scaled, tag_unscaled, total_scaled, total_unscaled,
        conf, total_conf, rows_total);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}

int gb_launch_cv_sums(int n_jobs, const int64_t* lo, const int64_t* hi, const float* y, const float* yhat,
                      int n_tags, double* out, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    int64_t max_rows = 0;
    int rc = max_chunks_host(n_jobs, lo, hi, stream, &max_rows);
    if (rc) return rc;
    GB_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)n_jobs * 5 * n_tags, stream));
    const int chunks = (int)((max_rows + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK);
    if (chunks > 0) {
        dim3 grid(chunks, n_jobs);
        cv_sums_kernel<<<grid, RED_THREADS, 5 * RED_THREADS * sizeof(double), stream>>>(lo, hi, y, yhat, n_tags, out);
    }
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}
