// extern "C" surface of libgordo_b200.so (see include/gordo_b200.h).  Argument validation,
// the fleet handle, dispatch to the kernels' launchers.  No torch types, no hidden state other
// than the thread-local error string.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void gb_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int check_ff_arch(const gb200_ff_arch* a) {
    GB_REQUIRE(a != nullptr, "arch is NULL");
    GB_REQUIRE(a->n_layers >= 1 && a->n_layers <= GB200_MAX_LAYERS, "n_layers=%d out of range [1,%d]",
               a->n_layers, GB200_MAX_LAYERS);
    for (int l = 0; l <= a->n_layers; ++l)
        GB_REQUIRE(a->widths[l] >= 1, "widths[%d]=%d must be >= 1", l, a->widths[l]);
    for (int l = 0; l < a->n_layers; ++l)
        GB_REQUIRE(a->acts[l] >= GB200_ACT_LINEAR && a->acts[l] <= GB200_ACT_SOFTPLUS,
                   "acts[%d]=%d is not a GB200_ACT_* code", l, a->acts[l]);
    return GB_OK;
}

extern "C" {

int gb200_abi_version(void) { return GB200_ABI_VERSION; }

const char* gb200_last_error(void) { return g_err; }

int gb200_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    int dev = 0;
    GB_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GB_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (name && name_len > 0) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    return GB_OK;
}

static int fleet_create_impl(gb200_fleet** out, int32_t n_machines, const int64_t* lo, const int64_t* hi) {
    gb200_fleet* f = (gb200_fleet*)calloc(1, sizeof(gb200_fleet));
    GB_REQUIRE(f != nullptr, "host allocation failed");
    f->n_machines = n_machines;
    const size_t nb = sizeof(int64_t) * (size_t)n_machines;
    f->h_row_lo = (int64_t*)malloc(nb);
    f->h_row_hi = (int64_t*)malloc(nb);
    f->h_tile_off = (int32_t*)malloc(sizeof(int32_t) * (n_machines + 1));
    if (!f->h_row_lo || !f->h_row_hi || !f->h_tile_off) { gb200_fleet_destroy(f); gb_set_error("host allocation failed"); return GB_ERR_ARG; }
    memcpy(f->h_row_lo, lo, nb); memcpy(f->h_row_hi, hi, nb);
    int64_t tiles = 0, rows = 0;
    for (int m = 0; m < n_machines; ++m) {
        f->h_tile_off[m] = (int32_t)tiles;
        tiles += (hi[m] - lo[m] + 127) / 128;
        rows += hi[m] - lo[m];
        if (tiles > 0x7fffffff) { gb200_fleet_destroy(f); gb_set_error("too many row tiles for one fleet"); return GB_ERR_ARG; }
    }
    f->h_tile_off[n_machines] = (int32_t)tiles;
    f->tiles_total = (int32_t)tiles;
    f->rows_total = rows;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&f->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_row_lo, nb);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_row_hi, nb);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_tile_off, sizeof(int32_t) * (n_machines + 1));
    if (e == cudaSuccess) e = cudaMemcpy(f->d_row_lo, f->h_row_lo, nb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(f->d_row_hi, f->h_row_hi, nb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(f->d_tile_off, f->h_tile_off, sizeof(int32_t) * (n_machines + 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        gb_set_error("gb200_fleet_create: %s", cudaGetErrorString(e));
        gb200_fleet_destroy(f);
        return GB_ERR_CUDA;
    }
    *out = f;
    return GB_OK;
}

int gb200_fleet_create(gb200_fleet** out, int32_t n_machines, const int64_t* row_off_host) {
    GB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    GB_REQUIRE(n_machines >= 1, "n_machines=%d must be >= 1", n_machines);
    GB_REQUIRE(row_off_host != nullptr, "row_off_host is NULL");
    for (int m = 0; m < n_machines; ++m)
        GB_REQUIRE(row_off_host[m + 1] >= row_off_host[m], "row_off_host must be non-decreasing (machine %d)", m);
    return fleet_create_impl(out, n_machines, row_off_host, row_off_host + 1);
}

int gb200_fleet_create_ranges(gb200_fleet** out, int32_t n_machines, const int64_t* rows_lo_host,
                              const int64_t* rows_hi_host) {
    GB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    GB_REQUIRE(n_machines >= 1, "n_machines=%d must be >= 1", n_machines);
    GB_REQUIRE(rows_lo_host && rows_hi_host, "row range arrays are NULL");
    for (int m = 0; m < n_machines; ++m)
        GB_REQUIRE(rows_hi_host[m] >= rows_lo_host[m] && rows_lo_host[m] >= 0, "bad row range for machine %d", m);
    return fleet_create_impl(out, n_machines, rows_lo_host, rows_hi_host);
}

void gb200_fleet_destroy(gb200_fleet* f) {
    if (!f) return;
    if (f->d_row_lo) cudaFree(f->d_row_lo);
    if (f->d_row_hi) cudaFree(f->d_row_hi);
    if (f->d_tile_off) cudaFree(f->d_tile_off);
    free(f->h_row_lo); free(f->h_row_hi); free(f->h_tile_off);
    free(f);
}

int64_t gb200_ff_param_count(const gb200_ff_arch* a) {
    if (!a || a->n_layers < 1 || a->n_layers > GB200_MAX_LAYERS) return 0;
    int64_t n = 0;
    for (int l = 0; l < a->n_layers; ++l) n += (int64_t)a->widths[l] * a->widths[l + 1] + a->widths[l + 1];
    return n;
}

static int is_tc_prec(int precision) { return precision == GB200_PREC_BF16_TC || precision == GB200_PREC_F16X3_TC; }

int64_t gb200_ff_packed_bytes(const gb200_ff_arch* arch) {
    if (check_ff_arch(arch)) return 0;
    return gb_ff_packed_bytes(arch, GB200_PREC_BF16_TC);
}

int64_t gb200_ff_packed_bytes_prec(const gb200_ff_arch* arch, int precision) {
    if (check_ff_arch(arch) || !is_tc_prec(precision)) return 0;
    return gb_ff_packed_bytes(arch, precision);
}

int gb200_ff_pack_bf16(const gb200_ff_arch* arch, int32_t n_machines, const float* params,
                       void* packed_bf16, void* stream) {
    return gb200_ff_pack(arch, GB200_PREC_BF16_TC, n_machines, params, packed_bf16, stream);
}

int gb200_ff_pack(const gb200_ff_arch* arch, int precision, int32_t n_machines, const float* params,
                  void* packed, void* stream) {
    int rc = check_ff_arch(arch); if (rc) return rc;
    GB_REQUIRE(is_tc_prec(precision), "gb200_ff_pack: precision %d has no operand image", precision);
    GB_REQUIRE(params && packed, "params / packed is NULL");
    GB_REQUIRE(gb_ff_packed_bytes(arch, precision) > 0, "topology is not eligible for this tensor-core path");
    return gb_launch_ff_pack(arch, precision, n_machines, params, packed, (cudaStream_t)stream);
}

int gb200_ff_score(gb200_fleet* f, const gb200_ff_arch* arch, int precision,
                   const float* params, const void* packed_bf16,
                   const float* in_scale, const float* in_min, const float* err_scale,
                   const float* feat_thr, const float* agg_thr,
                   const float* x, const float* y,
                   float* model_out, float* tag_scaled, float* tag_unscaled,
                   float* total_scaled, float* total_unscaled,
                   float* conf, float* total_conf, float* activity_l1, void* stream) {
    GB_REQUIRE(f != nullptr, "fleet is NULL");
    int rc = check_ff_arch(arch); if (rc) return rc;
    GB_REQUIRE(x != nullptr, "x is NULL");
    GB_REQUIRE((in_scale == nullptr) == (in_min == nullptr), "in_scale and in_min must be given together");
    GB_REQUIRE(y != nullptr || arch->widths[0] == arch->widths[arch->n_layers],
               "y may alias x only when n_features == n_features_out");
    if (precision == GB200_PREC_F32) {
        GB_REQUIRE(params != nullptr, "params is NULL");
        return gb_launch_ff_score_f32(f, arch, params, in_scale, in_min, err_scale, feat_thr, agg_thr, x, y,
                                      model_out, tag_scaled, tag_unscaled, total_scaled, total_unscaled,
                                      conf, total_conf, activity_l1, (cudaStream_t)stream);
    }
    if (is_tc_prec(precision)) {
        GB_REQUIRE(activity_l1 == nullptr, "activity_l1 is only produced by GB200_PREC_F32");
        GB_REQUIRE(packed_bf16 != nullptr, "the operand image is NULL (call gb200_ff_pack first)");
        GB_REQUIRE(gb_ff_packed_bytes(arch, precision) > 0, "topology is not eligible for this tensor-core path");
        return gb_launch_ff_score_tc(f, arch, precision, packed_bf16, in_scale, in_min, err_scale, feat_thr, agg_thr, x, y,
                                     model_out, tag_scaled, tag_unscaled, total_scaled, total_unscaled,
                                     conf, total_conf, (cudaStream_t)stream);
    }
    gb_set_error("unknown precision %d", precision);
    return GB_ERR_ARG;
}

int gb200_minmax_fit(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi,
                     const float* x, int32_t n_tags, float* scale, float* min_, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_tags >= 1, "bad n_jobs / n_tags");
    GB_REQUIRE(rows_lo && rows_hi && x && scale && min_, "NULL argument");
    return gb_launch_minmax_fit(n_jobs, rows_lo, rows_hi, x, n_tags, scale, min_, (cudaStream_t)stream);
}

int gb200_rolling_min_max(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi,
                          const float* v, int32_t n_cols, int32_t window, float* out, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_cols >= 1, "bad n_jobs / n_cols");
    GB_REQUIRE(rows_lo && rows_hi && v && out, "NULL argument");
    return gb_launch_rolling_min_max(n_jobs, rows_lo, rows_hi, v, n_cols, window, out, (cudaStream_t)stream);
}

int gb200_ff_fit(const gb200_ff_arch* arch, const gb200_adam* adam, int32_t n_jobs,
                 const int64_t* job_rows_lo, const int64_t* job_rows_hi,
                 const int32_t* job_scale_slot, const float* in_scale, const float* in_min,
                 const float* x, const float* y,
                 const int64_t* perm_off, const int32_t* perm_pool,
                 int32_t epochs, int32_t batch_size, int32_t l1_mean,
                 float* params, float* adam_mv, int64_t* adam_t,
                 float* hist_loss, float* hist_acc, void* stream) {
    int rc = check_ff_arch(arch); if (rc) return rc;
    GB_REQUIRE(adam != nullptr, "adam is NULL");
    GB_REQUIRE(n_jobs >= 0 && epochs >= 1 && batch_size >= 1, "bad n_jobs / epochs / batch_size");
    GB_REQUIRE(job_rows_lo && job_rows_hi && x && params && adam_mv, "NULL argument");
    GB_REQUIRE((in_scale == nullptr) == (in_min == nullptr), "in_scale and in_min must be given together");
    GB_REQUIRE((perm_off == nullptr) == (perm_pool == nullptr), "perm_off and perm_pool must be given together");
    GB_REQUIRE(y != nullptr || arch->widths[0] == arch->widths[arch->n_layers],
               "y may alias x only when n_features == n_features_out");
    return gb_launch_ff_fit(arch, adam, n_jobs, job_rows_lo, job_rows_hi, job_scale_slot, in_scale, in_min,
                            x, y, perm_off, perm_pool, epochs, batch_size, l1_mean, params, adam_mv, adam_t,
                            hist_loss, hist_acc, (cudaStream_t)stream);
}

int gb200_smooth(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* v,
                 int32_t n_cols, int32_t method, int32_t window, float* out, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_cols >= 1, "bad n_jobs / n_cols");
    GB_REQUIRE(rows_lo && rows_hi && v && out, "NULL argument");
    return gb_launch_smooth(n_jobs, rows_lo, rows_hi, v, n_cols, method, window, out, (cudaStream_t)stream);
}

int gb200_quantile(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* v,
                   int32_t n_cols, double q, double* out, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_cols >= 1, "bad n_jobs / n_cols");
    GB_REQUIRE(rows_lo && rows_hi && v && out, "NULL argument");
    return gb_launch_quantile(n_jobs, rows_lo, rows_hi, v, n_cols, q, out, (cudaStream_t)stream);
}

int gb200_cv_sums(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* y,
                  const float* yhat, int32_t n_tags, double* sums, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_tags >= 1, "bad n_jobs / n_tags");
    GB_REQUIRE(rows_lo && rows_hi && y && yhat && sums, "NULL argument");
    return gb_launch_cv_sums(n_jobs, rows_lo, rows_hi, y, yhat, n_tags, sums, (cudaStream_t)stream);
}

int gb200_score_outputs(int32_t n_machines, const int64_t* out_row_off, const int64_t* y_row_off,
                        int32_t n_tags, const float* model_out, const float* y, const float* err_scale,
                        const float* feat_thr, const float* agg_thr,
                        float* tag_scaled, float* tag_unscaled, float* total_scaled,
                        float* total_unscaled, float* conf, float* total_conf, void* stream) {
    GB_REQUIRE(n_machines >= 0 && n_tags >= 1, "bad n_machines / n_tags");
    GB_REQUIRE(out_row_off && y_row_off && model_out && y, "NULL argument");
    return gb_launch_score_outputs(n_machines, out_row_off, y_row_off, n_tags, model_out, y, err_scale,
                                   feat_thr, agg_thr, tag_scaled, tag_unscaled, total_scaled, total_unscaled,
                                   conf, total_conf, (cudaStream_t)stream);
}

int gb200_resample(int32_t n_series, const int64_t* point_off, const int64_t* ts_ns, const double* values,
                   const int64_t* bin0_ns, const int64_t* n_bins, const int64_t* out_off, const int64_t* out_stride,
                   int64_t step_ns, int32_t agg, int64_t max_bins, int64_t n_points_total, int64_t total_bins,
                   double* out, void* stream) {
    GB_REQUIRE(n_series >= 0, "bad n_series");
    GB_REQUIRE(n_series == 0 || (point_off && ts_ns && values && bin0_ns && n_bins && out_off && out_stride && out), "NULL argument");
    return gb_launch_resample(n_series, point_off, ts_ns, values, bin0_ns, n_bins, out_off, out_stride, step_ns, agg,
                              max_bins, n_points_total, total_bins, out, (cudaStream_t)stream);
}

int gb200_interpolate(int32_t n_series, const int64_t* n_bins, const int64_t* off, const int64_t* stride,
                      int32_t method, int64_t limit, double* data, void* stream) {
    GB_REQUIRE(n_series >= 0, "bad n_series");
    GB_REQUIRE(n_series == 0 || (n_bins && off && stride && data), "NULL argument");
    return gb_launch_interpolate(n_series, n_bins, off, stride, method, limit, data, (cudaStream_t)stream);
}

int gb200_filter_rows(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const double* data,
                      int32_t n_cols, const int64_t* row_ts_ns, int64_t ts_base_ns,
                      const int32_t* ops_host, const int32_t* args_host, int32_t n_ops,
                      const double* consts_host, int32_t n_consts, int32_t buffer_size, uint8_t* keep, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_cols >= 1, "bad n_jobs / n_cols");
    GB_REQUIRE(rows_lo && rows_hi && data && ops_host && args_host && keep, "NULL argument");
    GB_REQUIRE(n_consts == 0 || consts_host, "NULL constants");
    return gb_launch_filter_rows(n_jobs, rows_lo, rows_hi, data, n_cols, row_ts_ns, ts_base_ns, ops_host, args_host, n_ops,
                                 consts_host, n_consts, buffer_size, keep, (cudaStream_t)stream);
}

int gb200_compact_rows(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const double* data,
                       int32_t n_cols, const int64_t* row_ts_ns, const uint8_t* keep, double* out, float* out_f32,
                       int64_t* out_ts, int64_t* new_rows_lo, int64_t* new_rows_hi, void* stream) {
    GB_REQUIRE(n_jobs >= 0 && n_cols >= 1, "bad n_jobs / n_cols");
    GB_REQUIRE(rows_lo && rows_hi && data && keep && out && new_rows_lo && new_rows_hi, "NULL argument");
    GB_REQUIRE(!out_ts || row_ts_ns, "out_ts needs row_ts_ns");
    return gb_launch_compact_rows(n_jobs, rows_lo, rows_hi, data, n_cols, row_ts_ns, keep, out, out_f32, out_ts,
                                  new_rows_lo, new_rows_hi, (cudaStream_t)stream);
}

}  // extern "C"
