// host_cols: the HOST half of the fleet response path (gb200_host_expand_columns).
//
// A fleet `.anomaly()` call is PCIe-bound: of the four [R, T] matrices of the frame
// (diff.py:350-444) three are per-column rescalings of the reconstruction error,
//     tag-anomaly-unscaled = |yhat - y|,  tag-anomaly-scaled = |.| * |scale_|,  anomaly-confidence = |.| / thr,
// so the device sends yhat (and the row totals it reduced on chip) once and the columns the transfer
// plan leaves to the host are written here, straight into the caller's response buffers, while the
// next chunk is on the wire.  Pure streaming: 2 reads + up to 3 non-temporal writes per element,
// rows split over `n_threads` host threads (they inherit the caller's CPU affinity = the GPU's NUMA
// node).  Not a fallback for anything: the Dense stack and the totals always come from the device.
#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>
#include "common.cuh"

namespace {

struct ExpandJob {
    int64_t n_rows; int T;
    const float* yhat; const float* y;
    float* tu; float* ts; float* cf;
};

// flat range [i0, i1) of one Machine's [n_rows, T] block; column of flat index i = i % T.
// pat_* are the per-column factors extended periodically by 4 so a 4-wide load never wraps.
void expand_range(const ExpandJob& j, int64_t i0, int64_t i1, const float* pat_es, const float* pat_ift) {
    const int T = j.T;
    const __m128 sign = _mm_castsi128_ps(_mm_set1_epi32(0x7fffffff));
    int64_t i = i0;
    auto scalar = [&](int64_t k) {
        const int c = (int)(k % T);
        const float d = fabsf(j.yhat[k] - j.y[k]);
        if (j.tu) j.tu[k] = d;
        if (j.ts) j.ts[k] = d * pat_es[c];
        if (j.cf) j.cf[k] = d * pat_ift[c];
    };
    // head: up to the first index where every output pointer is 16-byte aligned (they share the
    // alignment when the buffers are, which pinned allocations guarantee; otherwise go scalar)
    const float* first = j.tu ? j.tu : (j.ts ? j.ts : j.cf);
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(first) & 15;
    const bool same_align = (!j.tu || (reinterpret_cast<uintptr_t>(j.tu) & 15) == a0) &&
                            (!j.ts || (reinterpret_cast<uintptr_t>(j.ts) & 15) == a0) &&
                            (!j.cf || (reinterpret_cast<uintptr_t>(j.cf) & 15) == a0) && (a0 & 3) == 0;
    const uintptr_t mis = (reinterpret_cast<uintptr_t>(first) + (uintptr_t)i * 4) & 15;
    if (!same_align) { for (; i < i1; ++i) scalar(i); return; }
    if (mis) for (int64_t head = (16 - (int64_t)mis) / 4; head > 0 && i < i1; --head, ++i) scalar(i);
    int c = (int)(i % T);
    for (; i + 4 <= i1; i += 4) {
        const __m128 d = _mm_and_ps(_mm_sub_ps(_mm_loadu_ps(j.yhat + i), _mm_loadu_ps(j.y + i)), sign);
        if (j.tu) _mm_stream_ps(j.tu + i, d);
        if (j.ts) _mm_stream_ps(j.ts + i, _mm_mul_ps(d, _mm_loadu_ps(pat_es + c)));
        if (j.cf) _mm_stream_ps(j.cf + i, _mm_mul_ps(d, _mm_loadu_ps(pat_ift + c)));
        c += 4; if (c >= T) c -= T;
    }
    for (; i < i1; ++i) scalar(i);
}

}  // namespace

extern "C" int gb200_host_expand_columns(int32_t n_machines, const int64_t* row_off_host, int32_t T,
                                         const float* yhat_host, const float* y_host,
                                         const float* err_scale_host, const float* feat_thr_host,
                                         float* tag_unscaled_host, float* tag_scaled_host, float* conf_host,
                                         int32_t n_threads) {
    GB_REQUIRE(n_machines >= 0 && T >= 1, "host_expand_columns: bad sizes");
    GB_REQUIRE(row_off_host && yhat_host && y_host, "host_expand_columns: NULL input");
    GB_REQUIRE(!tag_scaled_host || err_scale_host, "host_expand_columns: tag_scaled needs err_scale");
    GB_REQUIRE(!conf_host || feat_thr_host, "host_expand_columns: conf needs feat_thr");
    if (n_machines == 0 || (!tag_unscaled_host && !tag_scaled_host && !conf_host)) return GB_OK;
    if (n_threads < 1) n_threads = 1;
    const int64_t base = row_off_host[0];
    const int64_t total = (row_off_host[n_machines] - base) * (int64_t)T;
    if (total <= 0) return GB_OK;
    if ((int64_t)n_threads * 4096 > total) n_threads = (int)((total + 4095) / 4096);
    // periodic factor tables per Machine: [M][2][T + 4]
    const int TP = T + 4;
    std::vector<float> pat((size_t)n_machines * 2 * TP);
    for (int m = 0; m < n_machines; ++m) {
        float* pe = pat.data() + (size_t)m * 2 * TP; float* pi = pe + TP;
        for (int i = 0; i < TP; ++i) {
            const int c = i % T;
            pe[i] = err_scale_host ? fabsf(err_scale_host[(size_t)m * T + c]) : 1.0f;
            pi[i] = feat_thr_host ? 1.0f / feat_thr_host[(size_t)m * T + c] : 0.0f;
        }
    }
    // all pointers are relative to row row_off[0] (the chunk's first row)
    auto work = [&](int t) {
        // this thread's share of the flat element range, cut at multiples of 16 elements (64-byte lines)
        int64_t f0 = total * t / n_threads, f1 = total * (t + 1) / n_threads;
        f0 &= ~(int64_t)15; if (t + 1 < n_threads) f1 &= ~(int64_t)15;
        for (int m = 0; m < n_machines && f0 < f1; ++m) {
            const int64_t m0 = (row_off_host[m] - base) * T, m1 = (row_off_host[m + 1] - base) * T;
            if (m1 <= f0) continue;
            if (m0 >= f1) break;
            const int64_t a = f0 > m0 ? f0 : m0, b = f1 < m1 ? f1 : m1;
            ExpandJob j{row_off_host[m + 1] - row_off_host[m], T, yhat_host + m0, y_host + m0,
                        tag_unscaled_host ? tag_unscaled_host + m0 : nullptr,
                        tag_scaled_host ? tag_scaled_host + m0 : nullptr, conf_host ? conf_host + m0 : nullptr};
            const float* pe = pat.data() + (size_t)m * 2 * TP;
            expand_range(j, a - m0, b - m0, pe, pe + TP);
            f0 = b;
        }
        _mm_sfence();
    };
    if (n_threads == 1) { work(0); return GB_OK; }
    std::vector<std::thread> th;
    th.reserve(n_threads - 1);
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return GB_OK;
}

// Host memory bandwidth probe for the transfer plan: a[i] = b[i] * s over `bytes` per array with
// n_threads threads and non-temporal stores; returns the seconds of the best of `reps` passes.
extern "C" double gb200_host_stream_seconds(float* dst_host, const float* src_host, int64_t n_floats,
                                            int32_t n_threads, int32_t reps) {
    if (!dst_host || !src_host || n_floats <= 0) return -1.0;
    if (n_threads < 1) n_threads = 1;
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        auto work = [&](int t) {
            int64_t f0 = (n_floats * t / n_threads) & ~(int64_t)15;
            int64_t f1 = t + 1 < n_threads ? (n_floats * (t + 1) / n_threads) & ~(int64_t)15 : n_floats;
            const __m128 s = _mm_set1_ps(1.0001f);
            int64_t i = f0;
            if (((reinterpret_cast<uintptr_t>(dst_host)) & 15) == 0)
                for (; i + 4 <= f1; i += 4) _mm_stream_ps(dst_host + i, _mm_mul_ps(_mm_loadu_ps(src_host + i), s));
            for (; i < f1; ++i) dst_host[i] = src_host[i] * 1.0001f;
            _mm_sfence();
        };
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    return best;
}
