/*
 * gordo_b200.h -- C-ABI of libgordo_b200.so: the B200 (sm_100a) implementation of gordo's
 * per-machine autoencoder anomaly path for a FLEET of independent Machines.
 *
 * The reference (equinor/gordo @ 99a4819d) has no FFI: the path is Python calling Keras / sklearn
 * / pandas one Machine at a time.  Each entry point below cites the reference interface whose
 * arithmetic it replaces; gordo_b200/ (Python) mirrors the reference's estimator surface on top.
 *
 * Conventions
 *   - plain C types only; every data pointer is a CALLER-OWNED DEVICE pointer (e.g.
 *     torch.Tensor.data_ptr()) unless the name ends in _host; nothing is allocated or freed
 *     behind the caller's back after gb200_fleet_create();
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on that stream and safe to issue concurrently on distinct streams;
 *   - return value: 0 = ok, negative = error; gb200_last_error() gives the thread-local text;
 *   - a "fleet" is a set of M Machines sharing ONE topology (same widths / activations);
 *     heterogeneous projects are bucketed by topology on the host, one fleet per bucket;
 *   - matrices are row-major; Machine m owns rows [row_off[m], row_off[m+1]) of the
 *     concatenated [rows_total, T] sample matrices;
 *   - parameters are one flat float32 vector per Machine: for each Dense layer W[in][out]
 *     then b[out]; for each LSTM layer W[in][4u], U[u][4u], b[4u] (gate order i,f,c,o),
 *     then the output Dense.
 */
#ifndef GORDO_B200_H
#define GORDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_MAX_LAYERS 16
#define GB200_ABI_VERSION 1

/* activation codes (Keras names: factories/feedforward_autoencoder.py:21-24) */
enum { GB200_ACT_LINEAR = 0, GB200_ACT_TANH = 1, GB200_ACT_RELU = 2, GB200_ACT_SIGMOID = 3,
       GB200_ACT_ELU = 4, GB200_ACT_SOFTPLUS = 5 };

/* precision of the dense stacks */
enum { GB200_PREC_F32 = 0,      /* fp32 FMA, bit-faithful ordering of the reference ops   */
       GB200_PREC_BF16_TC = 1,  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate */
       GB200_PREC_F16X3_TC = 2  /* fp32-grade on tcgen05: operands split into hi + lo fp16, 3 MMAs per K step,
                                   fp32 accumulate; feed-forward scorer only, tanh / sigmoid hidden layers  */ };

/* Feed-forward topology: what gordo/machine/model/factories/feedforward_autoencoder.py:15-104
 * builds as a Keras Sequential of Dense layers. */
typedef struct {
    int32_t n_layers;                        /* Dense layers (2*encoding_layers + 1 for hourglass) */
    int32_t widths[GB200_MAX_LAYERS + 1];    /* widths[0] = n_features, widths[n_layers] = n_features_out */
    int32_t acts[GB200_MAX_LAYERS];          /* GB200_ACT_* per layer */
    float   l1[GB200_MAX_LAYERS];            /* activity_regularizer l1 coefficient per layer (0 = none) */
} gb200_ff_arch;

/* LSTM topology: gordo/machine/model/factories/lstm_autoencoder.py:15-103. */
typedef struct {
    int32_t n_layers;                        /* LSTM layers (encoder + decoder) */
    int32_t n_features, n_features_out;
    int32_t units[GB200_MAX_LAYERS];
    int32_t acts[GB200_MAX_LAYERS];          /* `activation` of each LSTM layer */
    int32_t out_act;                         /* activation of the final Dense */
    int32_t lookback_window;                 /* models.py:467-470 */
    int32_t lookahead;                       /* 0 = KerasLSTMAutoEncoder, 1 = KerasLSTMForecast (models.py:701-710) */
} gb200_lstm_arch;

/* Adam as [3P] keras.optimizers.Adam applies it. */
typedef struct { float lr, beta_1, beta_2, epsilon; } gb200_adam;

typedef struct gb200_fleet gb200_fleet;      /* opaque */

int         gb200_abi_version(void);
const char* gb200_last_error(void);
/* number of SMs / device name of the current device (diagnostics for bench + tests) */
int         gb200_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* Create the schedule for a fleet of M machines whose sample rows are laid out by
 * row_off_host[M+1] (HOST pointer, int64).  Allocates the (small) device-side schedule once.
 * Replaces the `for machine in machines` loop of gordo/builder/local_build.py:69-70 and the
 * one-pod-per-Machine fan-out of argo-workflow.yml.template:1543-1557. */
int  gb200_fleet_create(gb200_fleet** out, int32_t n_machines, const int64_t* row_off_host);
/* Same, with an explicit row range [rows_lo_host[m], rows_hi_host[m]) per Machine (HOST int64):
 * used to treat sub-ranges -- e.g. the TimeSeriesSplit test folds of diff.py:209-215 -- as
 * virtual Machines of one launch.  Outputs are always indexed by absolute row. */
int  gb200_fleet_create_ranges(gb200_fleet** out, int32_t n_machines, const int64_t* rows_lo_host,
                               const int64_t* rows_hi_host);
void gb200_fleet_destroy(gb200_fleet* f);

/* ---------------------------------------------------------------------------------------------
 * Feed-forward autoencoder: forward + DiffBasedAnomalyDetector scoring, fused.
 * Replaces, per Machine: Pipeline.predict = MinMaxScaler.transform + KerasAutoEncoder.predict
 * (gordo/machine/model/models.py:289-300) and the column arithmetic of
 * DiffBasedAnomalyDetector.anomaly (gordo/machine/model/anomaly/diff.py:336-444):
 *   yhat                   = DenseStack(x * in_scale + in_min)
 *   tag_anomaly_unscaled   = |yhat - y|                        (diff.py:371-382)
 *   tag_anomaly_scaled     = |S(yhat) - S(y)| = |err_scale| * |yhat - y|   (diff.py:350-363)
 *   total_anomaly_*        = mean_tags(tag_anomaly_*^2)        (diff.py:366-368, 383-385)
 *   anomaly_confidence     = tag_anomaly_unscaled / feat_thr   (diff.py:420-434)
 *   total_anomaly_conf     = total_anomaly_scaled / agg_thr    (diff.py:438-444)
 * Any output pointer may be NULL (skipped).  y == NULL means y aliases x (autoencoder).
 * params       : [M, P] fp32 (GB200_PREC_F32) -- always required
 * packed_bf16  : the operand image of the tensor-core precisions: [M, gb200_ff_packed_bytes_prec()] from
 *                gb200_ff_pack (GB200_PREC_BF16_TC / GB200_PREC_F16X3_TC), else NULL
 * in_scale/in_min : [M, T_in] fp32 (NULL = identity); err_scale : [M, T_out] fp32 (NULL = 1)
 * feat_thr : [M, T_out] or NULL; agg_thr : [M] or NULL
 * activity_l1 : optional [rows] (GB200_PREC_F32 only): sum_l l1[l]*sum_j|h_l[j]| per row -- the activity-
 *   regulariser part of the Keras loss (feedforward_autoencoder.py:78-81), needed for val_loss
 */
int gb200_ff_score(gb200_fleet* f, const gb200_ff_arch* arch, int precision,
                   const float* params, const void* packed_bf16,
                   const float* in_scale, const float* in_min, const float* err_scale,
                   const float* feat_thr, const float* agg_thr,
                   const float* x, const float* y,
                   float* model_out, float* tag_scaled, float* tag_unscaled,
                   float* total_scaled, float* total_unscaled,
                   float* conf, float* total_conf, float* activity_l1, void* stream);

/* bytes per Machine of the tensor-core operand image (bf16 weights in the tcgen05 canonical
 * K-major shared-memory layout + fp32 biases); 0 if the topology is not eligible (a width > 256) */
int64_t gb200_ff_packed_bytes(const gb200_ff_arch* arch);
int gb200_ff_pack_bf16(const gb200_ff_arch* arch, int32_t n_machines, const float* params,
                       void* packed_bf16, void* stream);
/* the same for either tensor-core precision (the F16X3 image holds a hi and a lo fp16 copy of every layer) */
int64_t gb200_ff_packed_bytes_prec(const gb200_ff_arch* arch, int precision);
int gb200_ff_pack(const gb200_ff_arch* arch, int precision, int32_t n_machines, const float* params,
                  void* packed, void* stream);
int64_t gb200_ff_param_count(const gb200_ff_arch* arch);

/* ---------------------------------------------------------------------------------------------
 * MinMaxScaler.fit for every Machine of the fleet (sklearn MinMaxScaler: examples/config.yaml:75-82,
 * diff.py:173): per tag min / max over the Machine's rows [lo, hi) (job-relative, see below)
 * -> scale = 1/(max-min) (1 where max == min), min_ = -min*scale.
 * rows_lo / rows_hi: [n_jobs] int64 DEVICE arrays of absolute row ranges; outputs [n_jobs, T].
 */
int gb200_minmax_fit(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi,
                     const float* x, int32_t n_tags, float* scale, float* min_, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Thresholds: pandas `.rolling(window).min().max()` per column over row ranges
 * (diff.py:229-233, 241-248).  v: [rows_total, n_cols]; out: [n_jobs, n_cols] (NaN if the range
 * has fewer than `window` rows).
 */
int gb200_rolling_min_max(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi,
                          const float* v, int32_t n_cols, int32_t window, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Smoothing of the anomaly scores (diff.py:302-308 `_smoothing`; the smooth-* columns of
 * .anomaly(), diff.py:387-415; the validation metric of DiffBasedKFCVAnomalyDetector, diff.py:631-635):
 *   GB200_SMOOTH_SMM  pandas rolling(window).median()     (NaN for the first window-1 rows of a job)
 *   GB200_SMOOTH_SMA  pandas rolling(window).mean()       (same)
 *   GB200_SMOOTH_EWMA pandas ewm(span=window).mean()      (adjust=True, ignore_na=False)
 * applied to every column of v [rows_total, n_cols] over each job's rows [lo, hi), restarted at lo.
 * out: [rows_total, n_cols]; rows outside every job are left untouched.  NaN inputs follow pandas
 * (a window holding a NaN yields NaN; EWMA carries the last mean over a NaN).
 */
enum { GB200_SMOOTH_SMM = 0, GB200_SMOOTH_SMA = 1, GB200_SMOOTH_EWMA = 2 };
int gb200_smooth(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* v,
                 int32_t n_cols, int32_t method, int32_t window, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Percentile thresholds: pandas DataFrame.quantile(q) (linear interpolation, NaN skipped) of every
 * column of v [rows_total, n_cols] over each job's rows (diff.py:631-635 `_calculate_threshold`).
 * out: [n_jobs, n_cols] float64 (NaN when a column has no finite row in the range).
 */
int gb200_quantile(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* v,
                   int32_t n_cols, double q, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cross-validation scoring sums (gordo/builder/build_model.py:377-446 builds 4 x (T+1) sklearn
 * scorers -- explained variance, r2, MSE, MAE per tag and averaged -- each a host pass over the
 * fold): one pass per fold on the device.  sums: [n_jobs, 5, n_tags] float64 =
 * sum y, sum y^2, sum e, sum e^2, sum |e| with e = y - yhat over rows [lo, hi) (absolute rows of
 * both y and yhat); the four metrics of MinMax-scaled data follow on the host (r2 / explained
 * variance are scale invariant, MSE scales with scale^2, MAE with |scale|).
 */
int gb200_cv_sums(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const float* y,
                  const float* yhat, int32_t n_tags, double* sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Feed-forward training: Keras Model.fit (models.py:243-287 -> scikeras -> [3P] keras) for
 * n_jobs independent fits in one launch (CV folds and the final fit are separate jobs):
 * float32, mini-batches of `batch_size` in the order given by `perm` (NULL = natural order),
 * loss = MSE + sum_l l1[l]*sum|h_l| (l1_mean != 0 divides the activity term by the batch size),
 * Keras-form Adam.  One CTA owns one job; weights + Adam moments stay in shared memory.
 *   job_rows_lo/hi : [n_jobs] int64 absolute row range of the TRAINING rows in x / y
 *   job_scale_slot : [n_jobs] int32 index into in_scale/in_min rows (NULL = job index)
 *   perm           : [n_jobs] int64 offsets into perm_pool (int32 row indices relative to the
 *                    job's rows_lo, epochs * n_rows entries per job) or NULL
 *   params         : [n_jobs, P] fp32, in: initial weights, out: trained weights
 *   adam_mv        : [n_jobs, 2P] fp32 Adam moments (m then v), in/out -- zero them for a fresh
 *                    optimizer; a second fit on the same model continues from them, as Keras does
 *                    (models.py:282: the model is only built when self.model is None)
 *   adam_t         : [n_jobs] int64 optimizer step counts, in/out (NULL = start at 0, not stored)
 *   hist_loss/acc  : [n_jobs, epochs] fp32 outputs (Keras History 'loss' / 'accuracy'), may be NULL
 */
int gb200_ff_fit(const gb200_ff_arch* arch, const gb200_adam* adam, int32_t n_jobs,
                 const int64_t* job_rows_lo, const int64_t* job_rows_hi,
                 const int32_t* job_scale_slot, const float* in_scale, const float* in_min,
                 const float* x, const float* y,
                 const int64_t* perm_off, const int32_t* perm_pool,
                 int32_t epochs, int32_t batch_size, int32_t l1_mean,
                 float* params, float* adam_mv, int64_t* adam_t,
                 float* hist_loss, float* hist_acc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSTM autoencoder / forecast: windows are never materialised; window k of a Machine is rows
 * [k, k+L) of its (scaled) x and the target is row k + L - 1 + lookahead of y
 * (models.py:713-793).  Output rows: n_rows - L + 1 - lookahead per Machine, written at
 * out_row_off[m] (HOST-side layout is the caller's: gb200_lstm_out_rows()).
 */
int64_t gb200_lstm_param_count(const gb200_lstm_arch* arch);
int64_t gb200_lstm_out_rows(const gb200_lstm_arch* arch, int64_t n_rows);
/* scratch bytes the caller must provide for gb200_lstm_predict over `max_windows` windows at once
 * (depends on the precision: the tensor-core path keeps bf16 state tiles + packed weights there);
 * 0 if the topology is not eligible for that precision */
int64_t gb200_lstm_scratch_bytes(const gb200_lstm_arch* arch, int64_t max_windows, int precision);
/* KerasLSTMBaseEstimator.predict (models.py:618-660): yhat for every window of every Machine.
 * out_row_off: [M+1] int64 DEVICE offsets of each Machine's output rows in model_out: the prefix
 * sum of gb200_lstm_out_rows(arch, rows_m).  precision: GB200_PREC_F32 (exact fp32 FMAs) or
 * GB200_PREC_BF16_TC (tcgen05 step kernel, bf16 operands / fp32 accumulate and cell state). */
int gb200_lstm_predict(gb200_fleet* f, const gb200_lstm_arch* arch, int precision, const float* params,
                       const float* in_scale, const float* in_min, const float* x,
                       const int64_t* out_row_off, float* model_out,
                       void* scratch, int64_t scratch_bytes, void* stream);
/* KerasLSTMBaseEstimator.fit (models.py:557-616): primer step on the first window, then
 * time-ordered batches; one job per fit.  Returns Keras History 'loss' of the main fit in
 * hist_loss [n_jobs, epochs] and the primer's loss in primer_loss [n_jobs].
 * Arithmetic: float32 throughout; the batched GEMMs run on the tensor cores with a 3xTF32 split
 * (relative error of a product below 2^-21), the recurrence on CUDA cores.  Environment knobs for
 * tests: GB200_LSTM_GEMM = simt | tc | tcs, GB200_LSTM_REC = 0 (per-time-step launches). */
int gb200_lstm_fit(const gb200_lstm_arch* arch, const gb200_adam* adam, int32_t n_jobs,
                   const int64_t* job_rows_lo_host, const int64_t* job_rows_hi_host,
                   const float* in_scale, const float* in_min, const float* x, const float* y,
                   int32_t epochs, int32_t batch_size, float* params,
                   float* hist_loss, float* primer_loss,
                   void* scratch, int64_t scratch_bytes, void* stream);
int64_t gb200_lstm_fit_scratch_bytes(const gb200_lstm_arch* arch, int32_t n_jobs, int32_t batch_size);

/* Anomaly scoring of precomputed model output (used after gb200_lstm_predict, where the output
 * is offset against y): same columns as gb200_ff_score.  out_row_off [M+1] / y_row_off [M] are DEVICE
 * int64 arrays: Machine m's output rows [out_row_off[m], out_row_off[m+1]) align with rows of y
 * starting at y_row_off[m] (diff.py:359-363 `[-len(data):]`). */
int gb200_score_outputs(int32_t n_machines, const int64_t* out_row_off,
                        const int64_t* y_row_off, int32_t n_tags,
                        const float* model_out, const float* y, const float* err_scale,
                        const float* feat_thr, const float* agg_thr,
                        float* tag_scaled, float* tag_unscaled, float* total_scaled,
                        float* total_unscaled, float* conf, float* total_conf, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Host half of the fleet response path (pointers here are HOST pointers, no stream).
 *
 * gb200_host_expand_columns: the per-column rescalings of DiffBasedAnomalyDetector.anomaly
 * (gordo/machine/model/anomaly/diff.py:366-375 tag-anomaly-unscaled, :350-364 tag-anomaly-scaled,
 * :417-426 anomaly-confidence) written from the model output the device sent back, for the
 * n_machines Machines whose rows are [row_off_host[m], row_off_host[m+1]) RELATIVE to the data
 * pointers' first row (row_off_host[0] is that first row's absolute index and is subtracted).
 * Any of the three outputs may be NULL.  err_scale / feat_thr: [n_machines, T] float32.
 * Rows are split over n_threads host threads (non-temporal stores); the threads inherit the
 * caller's CPU affinity.  This is the host side of a PCIe-bound transfer plan, not a fallback:
 * model output and row totals always come from gb200_ff_score. */
int gb200_host_expand_columns(int32_t n_machines, const int64_t* row_off_host, int32_t T,
                              const float* yhat_host, const float* y_host,
                              const float* err_scale_host, const float* feat_thr_host,
                              float* tag_unscaled_host, float* tag_scaled_host, float* conf_host,
                              int32_t n_threads);
/* Host memory streaming probe (dst = src * c, non-temporal stores, best of reps): seconds per pass.
 * Lets the fleet server weigh PCIe bytes against host DRAM bytes on the box it runs on. */
double gb200_host_stream_seconds(float* dst_host, const float* src_host, int64_t n_floats,
                                 int32_t n_threads, int32_t reps);

/* ---------------------------------------------------------------------------------------------
 * Upstream of X: the arithmetic of `dataset.get_data()` between the data provider and the matrix
 * the builder trains on.  Reference call site: gordo/builder/build_model.py:208-213
 * (`GordoBaseDataset.from_dict(...).get_data()`); the implementation is [3P] gordo-core 0.3.6 (not vendored):
 * gordo_core/time_series.py `TimeSeriesDataset.join_timeseries` + `get_data`, gordo_core/filters/rows.py
 * `pandas_filter_rows` + `apply_buffer`.  All pointers are DEVICE pointers unless named *_host.
 *
 * gb200_resample: pandas `series.resample(resolution, label="left").agg(method)` of n_series raw series.
 *   Series s owns points [point_off[s], point_off[s+1]) of ts_ns (int64 ns since the epoch, ascending) / values
 *   (float64, NaN = missing sample).  Its bin b covers [bin0_ns[s] + b*step_ns, + step_ns) for b < n_bins[s] and is
 *   written to out[out_off[s] + b*out_stride[s]] (a column of its Machine's row-major [bins, tags] matrix).
 *   An empty bin is NaN (0 for SUM / COUNT).  max_bins = max n_bins, n_points_total / total_bins size the launch.
 */
enum { GB200_AGG_MEAN = 0, GB200_AGG_MIN = 1, GB200_AGG_MAX = 2, GB200_AGG_SUM = 3, GB200_AGG_COUNT = 4,
       GB200_AGG_FIRST = 5, GB200_AGG_LAST = 6 };
int gb200_resample(int32_t n_series, const int64_t* point_off, const int64_t* ts_ns, const double* values,
                   const int64_t* bin0_ns, const int64_t* n_bins, const int64_t* out_off, const int64_t* out_stride,
                   int64_t step_ns, int32_t agg, int64_t max_bins, int64_t n_points_total, int64_t total_bins,
                   double* out, void* stream);

/* gb200_interpolate: in place, per series (same addressing as gb200_resample's output):
 *   GB200_INTERP_LINEAR  pandas `.interpolate(limit=limit)`: linear over bin positions, forward only -- leading NaNs
 *                        stay, the first `limit` NaNs of a gap are filled, trailing NaNs take the last value
 *   GB200_INTERP_FFILL   pandas `.fillna(method="ffill", limit=limit)`
 * limit < 0 = no limit. */
enum { GB200_INTERP_NONE = 0, GB200_INTERP_LINEAR = 1, GB200_INTERP_FFILL = 2 };
int gb200_interpolate(int32_t n_series, const int64_t* n_bins, const int64_t* off, const int64_t* stride,
                      int32_t method, int64_t limit, double* data, void* stream);

/* gb200_filter_rows: keep[r] = predicate(row r) for the rows [rows_lo[j], rows_hi[j]) of every job j of a row-major
 * [rows_total, n_cols] float64 matrix, then every rejected row also rejects the buffer_size rows on either side of
 * it inside its job (`apply_buffer`).  The predicate is a postfix program (ops/args/consts are HOST arrays, at most
 * 96 operations / 48 constants / stack depth 16) compiled by the host side from `row_filter` /
 * `known_filter_periods` (pandas DataFrame.eval subset) or one of the built-in stages:
 *   [ALL_NOTNAN]              the dropna() after the inner join of the resampled series
 *   [ALL_BETWEEN c]           every column inside (consts[c], consts[c+1]): the dataset's low / high thresholds
 * GB200_OP_INDEX pushes (row_ts_ns[r] - ts_base_ns) as float64; comparisons push 1.0 / 0.0 (NaN compares false). */
enum { GB200_OP_CONST = 0, GB200_OP_COL = 1, GB200_OP_INDEX = 2, GB200_OP_NEG = 3, GB200_OP_ABS = 4, GB200_OP_NOT = 5,
       GB200_OP_ADD = 6, GB200_OP_SUB = 7, GB200_OP_MUL = 8, GB200_OP_DIV = 9, GB200_OP_POW = 10,
       GB200_OP_GT = 11, GB200_OP_GE = 12, GB200_OP_LT = 13, GB200_OP_LE = 14, GB200_OP_EQ = 15, GB200_OP_NE = 16,
       GB200_OP_AND = 17, GB200_OP_OR = 18, GB200_OP_ALL_FINITE = 19, GB200_OP_ALL_NOTNAN = 20, GB200_OP_ALL_BETWEEN = 21 };
int gb200_filter_rows(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const double* data,
                      int32_t n_cols, const int64_t* row_ts_ns, int64_t ts_base_ns,
                      const int32_t* ops_host, const int32_t* args_host, int32_t n_ops,
                      const double* consts_host, int32_t n_consts, int32_t buffer_size, uint8_t* keep, void* stream);

/* gb200_compact_rows: what `df[mask]` leaves -- the kept rows of every job packed back to back in job order.
 * out [<= rows_total, n_cols] float64, out_f32 the same as float32 (the dtype the model consumes; may be NULL),
 * out_ts the kept rows' timestamps (may be NULL with row_ts_ns), new_rows_lo / new_rows_hi [n_jobs] each job's rows
 * in the packed matrices. */
int gb200_compact_rows(int32_t n_jobs, const int64_t* rows_lo, const int64_t* rows_hi, const double* data,
                       int32_t n_cols, const int64_t* row_ts_ns, const uint8_t* keep, double* out, float* out_f32,
                       int64_t* out_ts, int64_t* new_rows_lo, int64_t* new_rows_hi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GORDO_B200_H */
