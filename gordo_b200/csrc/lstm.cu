// LSTM autoencoder / forecast kernels -- placeholder translation unit while the kernels are
// being brought up; every entry point fails loudly (no CPU fallback).
#include "common.cuh"

extern "C" {

int64_t gb200_lstm_param_count(const gb200_lstm_arch* a) {
    if (!a || a->n_layers < 1 || a->n_layers > GB200_MAX_LAYERS) return 0;
    int64_t n = 0, in = a->n_features;
    for (int l = 0; l < a->n_layers; ++l) { const int64_t u = a->units[l]; n += in * 4 * u + u * 4 * u + 4 * u; in = u; }
    return n + in * a->n_features_out + a->n_features_out;
}

int64_t gb200_lstm_out_rows(const gb200_lstm_arch* a, int64_t n_rows) {
    if (!a) return 0;
    const int64_t n = n_rows - a->lookback_window + 1 - a->lookahead;
    return n > 0 ? n : 0;
}

int64_t gb200_lstm_scratch_bytes(const gb200_lstm_arch*, int64_t) { return 0; }
int64_t gb200_lstm_fit_scratch_bytes(const gb200_lstm_arch*, int32_t, int32_t) { return 0; }

int gb200_lstm_predict(gb200_fleet*, const gb200_lstm_arch*, const float*, const float*, const float*,
                       const float*, const int64_t*, float*, void*, int64_t, void*) {
    gb_set_error("gb200_lstm_predict: not implemented in this build");
    return GB_ERR_UNSUPPORTED;
}

int gb200_lstm_fit(const gb200_lstm_arch*, const gb200_adam*, int32_t, const int64_t*, const int64_t*,
                   const float*, const float*, const float*, const float*, int32_t, int32_t, float*,
                   float*, float*, void*, int64_t, void*) {
    gb_set_error("gb200_lstm_fit: not implemented in this build");
    return GB_ERR_UNSUPPORTED;
}

}  // extern "C"
