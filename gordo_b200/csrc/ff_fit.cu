// ff_fit: Keras Model.fit for many independent small autoencoders, one CTA per fit job.
//
// Replaces models.py:243-287 -> scikeras -> [3P] keras Model.fit: float32
// mini-batch SGD with Keras-form Adam, loss = MSE + L1 activity terms, History loss/accuracy.
// The optimizer steps of one job are strictly sequential (ceil(N/batch) dependent steps per
// epoch), so parallelism comes from jobs: a CTA keeps its job's weights and both Adam moments in
// shared memory for the whole fit (T=50 hourglass: 3 x 9 497 x 4 B = 114 KB) and walks the
// mini-batches; CV folds and the final fit of every Machine are separate jobs of one launch.
//
// Per step (phases separated by __syncthreads):
//   gather  : h0[k][b] = x[row_b][k]*scale+min            (rows given by the permutation)
//   forward : h_{l+1}[n][b] = act(b_l[n] + sum_k h_l[k][b] W_l[k][n])
//   loss    : D_L[b][n] = dLoss/dz_L, MSE / accuracy / L1 sums
//   backward, l = L-1..0:  A: D_{l-1}[b][k] = (sum_n D_l[b][n] W_l[k][n] + l1 sign(h)) act'(h)
//                          B: g = sum_b h_l[k][b] D_l[b][n]  -> Adam update of W_l[k][n], b_l[n]
// Layouts: h feature-major [w][Bs] (Bs odd), D batch-major [B][ldd] (ldd odd): every phase reads
// shared memory either with consecutive lanes on consecutive words or as a broadcast.
#include "common.cuh"

namespace {

constexpr int FIT_THREADS = 256;

struct FitArgs {
    gb200_ff_arch arch;
    gb200_adam adam;
    const int64_t* lo; const int64_t* hi; const int32_t* scale_slot;
    const float* in_scale; const float* in_min;
    const float* x; const float* y;
    const int64_t* perm_off; const int32_t* perm_pool;
    int epochs, batch, l1_mean;
    float* params; float* adam_mv; int64_t* adam_t;
    float* hist_loss; float* hist_acc;
    int64_t n_params;
    int state_in_smem;      // 2: W,m,v in smem  1: W in smem  0: all global
    int h_floats;           // sum_l w_l * Bs
    int ldd;                // odd row stride of the D buffers
    int Bs;                 // odd column stride of h
};

// act' from the OUTPUT h alone (tanh 1-h^2, sigmoid h(1-h), relu/elu via sign of h,
// softplus sigma(z) = 1 - exp(-h))
__device__ __forceinline__ float act_grad_from_h(int code, float h) {
    switch (code) {
        case GB200_ACT_TANH:     return 1.0f - h * h;
        case GB200_ACT_RELU:     return h > 0.0f ? 1.0f : 0.0f;
        case GB200_ACT_SIGMOID:  return h * (1.0f - h);
        case GB200_ACT_ELU:      return h > 0.0f ? 1.0f : h + 1.0f;
        case GB200_ACT_SOFTPLUS: return 1.0f - expf(-h);
        default:                 return 1.0f;
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    #pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    #pragma unroll
    for (int i = 0; i < FIT_THREADS / 32; ++i) s += red[i];
    return s;
}

__global__ void __launch_bounds__(FIT_THREADS, 1)
ff_fit_kernel(const __grid_constant__ FitArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[FIT_THREADS / 32];
    __shared__ int s_argmax_hits;

    const int job = blockIdx.x, tid = threadIdx.x;
    const int L = a.arch.n_layers;
    const int T_in = a.arch.widths[0], T_out = a.arch.widths[L];
    const int B = a.batch, Bs = a.Bs, ldd = a.ldd;
    const int64_t P = a.n_params;

    float* hbuf = smem;                               // all layer activations of the batch
    float* D0 = hbuf + a.h_floats;                    // [B][ldd]
    float* D1 = D0 + B * ldd;
    float* st = D1 + B * ldd;                         // optional on-chip W / m / v
    float* gW = a.params + (size_t)job * P;
    float* gM = a.adam_mv + (size_t)job * 2 * P;
    float* gV = gM + P;
    float* W = a.state_in_smem >= 1 ? st : gW;
    float* Mo = a.state_in_smem >= 2 ? st + P : gM;
    float* Vo = a.state_in_smem >= 2 ? st + 2 * P : gV;
    if (a.state_in_smem >= 1) for (int64_t i = tid; i < P; i += FIT_THREADS) W[i] = gW[i];
    if (a.state_in_smem >= 2) for (int64_t i = tid; i < P; i += FIT_THREADS) { Mo[i] = gM[i]; Vo[i] = gV[i]; }

    const int64_t r_lo = a.lo[job];
    const int n = (int)(a.hi[job] - r_lo);
    const int slot = a.scale_slot ? a.scale_slot[job] : job;
    const float* sc = a.in_scale ? a.in_scale + (size_t)slot * T_in : nullptr;
    const float* mn = a.in_min ? a.in_min + (size_t)slot * T_in : nullptr;
    const float* ysrc = a.y ? a.y : a.x;
    const int32_t* perm = a.perm_pool ? a.perm_pool + a.perm_off[job] : nullptr;
    int64_t t = a.adam_t ? a.adam_t[job] : 0;
    const float b1 = a.adam.beta_1, b2 = a.adam.beta_2, eps = a.adam.epsilon, lr = a.adam.lr;
    __syncthreads();

    for (int e = 0; e < a.epochs; ++e) {
        double loss_acc = 0.0; int hits_acc = 0;          // meaningful in thread 0 only
        for (int s0 = 0; s0 < n; s0 += B) {
            const int nb = min(B, n - s0);
            // ---- gather + MinMax scale
            for (int i = tid; i < nb * T_in; i += FIT_THREADS) {
                const int b = i / T_in, k = i - b * T_in;
                const int64_t r = r_lo + (perm ? perm[(int64_t)e * n + s0 + b] : s0 + b);
                float v = a.x[r * T_in + k];
                if (sc) v = fmaf(v, sc[k], mn[k]);
                hbuf[k * Bs + b] = v;
            }
            if (tid == 0) s_argmax_hits = 0;
            __syncthreads();
            // ---- forward
            float l1_sum = 0.0f;
            {
                int ho = 0; int64_t po = 0;
                for (int l = 0; l < L; ++l) {
                    const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
                    const int code = a.arch.acts[l];
                    const float* Wl = W + po; const float* bl = Wl + (int64_t)win * wout;
                    const float* hin = hbuf + ho; float* hout = hbuf + ho + win * Bs;
                    const float c1 = a.arch.l1[l];
                    for (int i = tid; i < wout * 32; i += FIT_THREADS) {
                        const int nn = i >> 5, b = i & 31;
                        for (int bb = b; bb < nb; bb += 32) {
                            float acc = bl[nn];
                            for (int k = 0; k < win; ++k) acc = fmaf(hin[k * Bs + bb], Wl[k * wout + nn], acc);
                            const float h = gb_act(code, acc);
                            hout[nn * Bs + bb] = h;
                            if (c1 != 0.0f) l1_sum += c1 * fabsf(h);
                        }
                    }
                    ho += win * Bs; po += (int64_t)win * wout + wout;
                    __syncthreads();
                }
            }
            // ---- loss, dLoss/dz_L, accuracy
            const float* hL;
            {
                int ho = 0;
                for (int l = 0; l < L; ++l) ho += a.arch.widths[l] * Bs;
                hL = hbuf + ho;
            }
            float* Dcur = D0; float* Dnext = D1;
            float sq = 0.0f;
            {
                const float inv = 2.0f / (float)(nb * T_out);
                const int code = a.arch.acts[L - 1];
                const float cL = a.arch.l1[L - 1] * (a.l1_mean ? 1.0f / (float)nb : 1.0f);
                for (int i = tid; i < nb * T_out; i += FIT_THREADS) {
                    const int b = i / T_out, nn = i - b * T_out;
                    const int64_t r = r_lo + (perm ? perm[(int64_t)e * n + s0 + b] : s0 + b);
                    const float yh = hL[nn * Bs + b];
                    const float diff = yh - ysrc[r * T_out + nn];
                    sq = fmaf(diff, diff, sq);
                    float dh = inv * diff;
                    if (cL != 0.0f) dh += cL * (yh > 0.0f ? 1.0f : (yh < 0.0f ? -1.0f : 0.0f));
                    Dcur[b * ldd + nn] = dh * act_grad_from_h(code, yh);
                }
                if (tid < nb && a.hist_acc) {
                    // Keras 'accuracy' on a float [B,T] target = categorical accuracy; binary at T_out == 1
                    const int64_t r = r_lo + (perm ? perm[(int64_t)e * n + s0 + tid] : s0 + tid);
                    int hit;
                    if (T_out == 1) {
                        hit = ((hL[tid] > 0.5f ? 1.0f : 0.0f) == ysrc[r]) ? 1 : 0;
                    } else {
                        int ay = 0, ap = 0; float by = ysrc[r * T_out], bp = hL[tid];
                        for (int j = 1; j < T_out; ++j) {
                            const float vy = ysrc[r * T_out + j], vp = hL[j * Bs + tid];
                            if (vy > by) { by = vy; ay = j; }
                            if (vp > bp) { bp = vp; ap = j; }
                        }
                        hit = ay == ap;
                    }
                    if (hit) atomicAdd(&s_argmax_hits, 1);
                }
            }
            if (a.hist_loss) {
                const float tot_sq = block_sum(sq, red);
                const float tot_l1 = block_sum(l1_sum, red);
                if (tid == 0) {
                    const float batch_loss = tot_sq / (float)(nb * T_out) + (a.l1_mean ? tot_l1 / (float)nb : tot_l1);
                    loss_acc += (double)batch_loss * nb;
                    hits_acc += s_argmax_hits;
                }
            } else {
                __syncthreads();
            }
            // ---- Adam step size (Keras: alpha = lr*sqrt(1-b2^t)/(1-b1^t), eps outside)
            t += 1;
            const float tf = (float)t;
            const float alpha = lr * sqrtf(1.0f - powf(b2, tf)) / (1.0f - powf(b1, tf));
            const float l1_scale = a.l1_mean ? 1.0f / (float)nb : 1.0f;
            // ---- backward
            {
                int ho = 0; int64_t po = 0;
                int hoffs[GB200_MAX_LAYERS + 1]; int64_t poffs[GB200_MAX_LAYERS];
                for (int l = 0; l < L; ++l) {
                    hoffs[l] = ho; poffs[l] = po;
                    ho += a.arch.widths[l] * Bs; po += (int64_t)a.arch.widths[l] * a.arch.widths[l + 1] + a.arch.widths[l + 1];
                }
                hoffs[L] = ho;
                for (int l = L - 1; l >= 0; --l) {
                    const int win = a.arch.widths[l], wout = a.arch.widths[l + 1];
                    float* Wl = W + poffs[l];
                    const float* hin = hbuf + hoffs[l];
                    if (l > 0) {
                        // A: dLoss/dz_{l-1}  (reads the not-yet-updated W_l)
                        const int pcode = a.arch.acts[l - 1];
                        const float c1 = a.arch.l1[l - 1] * l1_scale;
                        for (int i = tid; i < win * 32; i += FIT_THREADS) {
                            const int k = i >> 5, b = i & 31;
                            for (int bb = b; bb < nb; bb += 32) {
                                float acc = 0.0f;
                                for (int nn = 0; nn < wout; ++nn) acc = fmaf(Dcur[bb * ldd + nn], Wl[k * wout + nn], acc);
                                const float h = hin[k * Bs + bb];
                                if (c1 != 0.0f) acc += c1 * (h > 0.0f ? 1.0f : (h < 0.0f ? -1.0f : 0.0f));
                                Dnext[bb * ldd + k] = acc * act_grad_from_h(pcode, h);
                            }
                        }
                        __syncthreads();
                    }
                    // B: gradients of W_l, b_l and their Adam update
                    const int n_w = win * wout;
                    for (int i = tid; i < n_w + wout; i += FIT_THREADS) {
                        float g = 0.0f;
                        if (i < n_w) {
                            const int k = i / wout, nn = i - k * wout;
                            for (int b = 0; b < nb; ++b) g = fmaf(hin[k * Bs + b], Dcur[b * ldd + nn], g);
                        } else {
                            const int nn = i - n_w;
                            for (int b = 0; b < nb; ++b) g += Dcur[b * ldd + nn];
                        }
                        const int64_t p = poffs[l] + i;
                        float m = Mo[p], v = Vo[p];
                        m += (g - m) * (1.0f - b1);
                        v += (g * g - v) * (1.0f - b2);
                        Mo[p] = m; Vo[p] = v;
                        Wl[i] -= alpha * m / (sqrtf(v) + eps);
                    }
                    __syncthreads();
                    float* tmp = Dcur; Dcur = Dnext; Dnext = tmp;
                }
            }
        }
        if (tid == 0) {
            if (a.hist_loss) a.hist_loss[(size_t)job * a.epochs + e] = n > 0 ? (float)(loss_acc / n) : NAN;
            if (a.hist_acc) a.hist_acc[(size_t)job * a.epochs + e] = n > 0 ? (float)hits_acc / (float)n : NAN;
        }
    }
    __syncthreads();
    if (a.state_in_smem >= 1) for (int64_t i = tid; i < P; i += FIT_THREADS) gW[i] = W[i];
    if (a.state_in_smem >= 2) for (int64_t i = tid; i < P; i += FIT_THREADS) { gM[i] = Mo[i]; gV[i] = Vo[i]; }
    if (tid == 0 && a.adam_t) a.adam_t[job] = t;
}

}  // namespace

int gb_launch_ff_fit(const gb200_ff_arch* arch, const gb200_adam* adam, int n_jobs,
                     const int64_t* lo, const int64_t* hi, const int32_t* scale_slot,
                     const float* in_scale, const float* in_min, const float* x, const float* y,
                     const int64_t* perm_off, const int32_t* perm_pool, int epochs, int batch_size,
                     int l1_mean, float* params, float* adam_mv, int64_t* adam_t, float* hist_loss,
                     float* hist_acc, cudaStream_t stream) {
    if (n_jobs <= 0) return GB_OK;
    FitArgs a{};
    a.arch = *arch; a.adam = *adam;
    a.lo = lo; a.hi = hi; a.scale_slot = scale_slot; a.in_scale = in_scale; a.in_min = in_min;
    a.x = x; a.y = y; a.perm_off = perm_off; a.perm_pool = perm_pool;
    a.epochs = epochs; a.batch = batch_size; a.l1_mean = l1_mean;
    a.params = params; a.adam_mv = adam_mv; a.adam_t = adam_t; a.hist_loss = hist_loss; a.hist_acc = hist_acc;
    a.n_params = gb200_ff_param_count(arch);
    int sum_w = 0, max_w = 0;
    for (int l = 0; l <= arch->n_layers; ++l) { sum_w += arch->widths[l]; if (arch->widths[l] > max_w) max_w = arch->widths[l]; }
    a.Bs = batch_size | 1;
    a.ldd = max_w | 1;
    a.h_floats = sum_w * a.Bs;
    const size_t base = ((size_t)a.h_floats + 2 * (size_t)batch_size * a.ldd) * sizeof(float);
    const size_t cap = 227 * 1024 - 256;
    GB_REQUIRE(base <= cap, "ff_fit: batch_size %d x widths do not fit in shared memory", batch_size);
    const size_t pbytes = (size_t)a.n_params * sizeof(float);
    a.state_in_smem = base + 3 * pbytes <= cap ? 2 : (base + pbytes <= cap ? 1 : 0);
    const size_t smem = base + (a.state_in_smem == 2 ? 3 * pbytes : a.state_in_smem == 1 ? pbytes : 0);
    GB_CUDA_CHECK(cudaFuncSetAttribute(ff_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ff_fit_kernel<<<n_jobs, FIT_THREADS, smem, stream>>>(a);
    GB_CUDA_CHECK(cudaGetLastError());
    return GB_OK;
}
