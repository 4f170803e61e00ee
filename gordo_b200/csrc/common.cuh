// Internal helpers shared by the sm_100a kernels of libgordo_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/gordo_b200.h"

// Opt-in shared-memory ceiling of sm_100 (227 KB).  Every kernel with a run-time smem size raises its
// limit to THIS constant, never to the size of the launch at hand: cudaFuncSetAttribute is
// process-global, and two host threads launching the same kernel for different topologies on
// different streams would otherwise race on it.
#define GB_SMEM_OPTIN_MAX 232448

// raise a kernel's dynamic shared-memory limit to everything the SM offers it: the opt-in maximum minus the
// kernel's own static shared memory (the attribute counts dynamic bytes only)
template <typename K>
static inline cudaError_t gb_allow_max_smem(K kern) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(kern));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(reinterpret_cast<const void*>(kern), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GB_SMEM_OPTIN_MAX - (int)fa.sharedSizeBytes);
}

#define GB_OK 0
#define GB_ERR_ARG -1
#define GB_ERR_CUDA -2
#define GB_ERR_UNSUPPORTED -3

void gb_set_error(const char* fmt, ...);

#define GB_CUDA_CHECK(expr)                                                         \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            gb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                         cudaGetErrorString(_e));                                   \
            return GB_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)

#define GB_REQUIRE(cond, ...)                                                       \
    do {                                                                            \
        if (!(cond)) { gb_set_error(__VA_ARGS__); return GB_ERR_ARG; }              \
    } while (0)

struct gb200_fleet {
    int32_t n_machines;
    int64_t rows_total;
    int32_t tiles_total;        // 128-row tiles over all machines
    int64_t* d_row_lo;          // [M] first row of each Machine
    int64_t* d_row_hi;          // [M] one past the last row
    int32_t* d_tile_off;        // [M+1] prefix of ceil(rows/128)
    int64_t* h_row_lo;          // host copies
    int64_t* h_row_hi;
    int32_t* h_tile_off;
    int sm_count;
};

// ---------------------------------------------------------------- activations (precise fp32)
__device__ __forceinline__ float gb_act(int code, float z) {
    switch (code) {
        case GB200_ACT_TANH:     return tanhf(z);
        case GB200_ACT_RELU:     return fmaxf(z, 0.0f);
        case GB200_ACT_SIGMOID:  return 1.0f / (1.0f + expf(-z));
        case GB200_ACT_ELU:      return z > 0.0f ? z : expm1f(z);
        case GB200_ACT_SOFTPLUS: return z > 0.0f ? z + log1pf(expf(-z)) : log1pf(expf(z));
        default:                 return z;
    }
}
// d act / dz given pre-activation z and output h
__device__ __forceinline__ float gb_act_grad(int code, float z, float h) {
    switch (code) {
        case GB200_ACT_TANH:     return 1.0f - h * h;
        case GB200_ACT_RELU:     return z > 0.0f ? 1.0f : 0.0f;
        case GB200_ACT_SIGMOID:  return h * (1.0f - h);
        case GB200_ACT_ELU:      return z > 0.0f ? 1.0f : h + 1.0f;
        case GB200_ACT_SOFTPLUS: return 1.0f / (1.0f + expf(-z));
        default:                 return 1.0f;
    }
}

static inline int gb_round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ int gb_round_up_dev(int x, int m) { return (x + m - 1) / m * m; }

// launchers implemented in the individual .cu files
int gb_launch_ff_score_f32(const gb200_fleet* f, const gb200_ff_arch* arch, const float* params,
                           const float* in_scale, const float* in_min, const float* err_scale,
                           const float* feat_thr, const float* agg_thr, const float* x, const float* y,
                           float* model_out, float* tag_scaled, float* tag_unscaled,
                           float* total_scaled, float* total_unscaled, float* conf, float* total_conf,
                           float* activity, cudaStream_t stream);
int gb_launch_ff_score_tc(const gb200_fleet* f, const gb200_ff_arch* arch, int prec, const void* packed,
                          const float* in_scale, const float* in_min, const float* err_scale,
                          const float* feat_thr, const float* agg_thr, const float* x, const float* y,
                          float* model_out, float* tag_scaled, float* tag_unscaled,
                          float* total_scaled, float* total_unscaled, float* conf, float* total_conf,
                          cudaStream_t stream);
int64_t gb_ff_packed_bytes(const gb200_ff_arch* arch, int prec);
int gb_launch_ff_pack(const gb200_ff_arch* arch, int prec, int n_machines, const float* params, void* packed,
                      cudaStream_t stream);
int gb_launch_minmax_fit(int n_jobs, const int64_t* lo, const int64_t* hi, const float* x, int n_tags,
                         float* scale, float* min_, cudaStream_t stream);
int gb_launch_rolling_min_max(int n_jobs, const int64_t* lo, const int64_t* hi, const float* v,
                              int n_cols, int window, float* out, cudaStream_t stream);
int gb_launch_ff_fit(const gb200_ff_arch* arch, const gb200_adam* adam, int n_jobs,
                     const int64_t* lo, const int64_t* hi, const int32_t* scale_slot,
                     const float* in_scale, const float* in_min, const float* x, const float* y,
                     const int64_t* perm_off, const int32_t* perm_pool, int epochs, int batch_size,
                     int l1_mean, float* params, float* adam_mv, int64_t* adam_t, float* hist_loss,
                     float* hist_acc, cudaStream_t stream);
int gb_launch_smooth(int n_jobs, const int64_t* lo, const int64_t* hi, const float* v, int n_cols,
                     int method, int window, float* out, cudaStream_t stream);
int gb_launch_quantile(int n_jobs, const int64_t* lo, const int64_t* hi, const float* v, int n_cols,
                       double q, double* out, cudaStream_t stream);
int gb_launch_resample(int n_series, const int64_t* point_off, const int64_t* ts, const double* val,
                       const int64_t* bin0, const int64_t* n_bins, const int64_t* out_off, const int64_t* out_stride,
                       int64_t step, int agg, int64_t max_bins, int64_t n_points, int64_t total_bins, double* out,
                       cudaStream_t stream);
int gb_launch_interpolate(int n_series, const int64_t* n_bins, const int64_t* off, const int64_t* stride, int method,
                          int64_t limit, double* data, cudaStream_t stream);
int gb_launch_filter_rows(int n_jobs, const int64_t* lo, const int64_t* hi, const double* data, int n_cols,
                          const int64_t* ts, int64_t ts_base, const int32_t* ops, const int32_t* args, int n_ops,
                          const double* consts, int n_consts, int buffer_size, uint8_t* keep, cudaStream_t stream);
int gb_launch_compact_rows(int n_jobs, const int64_t* lo, const int64_t* hi, const double* data, int n_cols,
                           const int64_t* ts, const uint8_t* keep, double* out, float* out_f32, int64_t* out_ts,
                           int64_t* new_lo, int64_t* new_hi, cudaStream_t stream);
int gb_launch_cv_sums(int n_jobs, const int64_t* lo, const int64_t* hi, const float* y, const float* yhat,
                      int n_tags, double* out, cudaStream_t stream);
int64_t gb_lstm_tc_scratch_bytes(const gb200_lstm_arch* arch, int64_t max_windows);
int gb_lstm_predict_tc(const gb200_fleet* f, const gb200_lstm_arch* arch, const float* params,
                       const float* in_scale, const float* in_min, const float* x,
                       const int64_t* out_row_off_host, float* model_out,
                       void* scratch, int64_t scratch_bytes, cudaStream_t stream);
int gb_launch_score_outputs(int n_machines, const int64_t* out_row_off, const int64_t* y_row_off,
                            int n_tags, const float* model_out, const float* y, const float* err_scale,
                            const float* feat_thr, const float* agg_thr, float* tag_scaled,
                            float* tag_unscaled, float* total_scaled, float* total_unscaled,
                            float* conf, float* total_conf, cudaStream_t stream);
