// HBM ceiling for the scorer's traffic MIX: 1 stream read, 4 streams written (+3 thin row vectors),
// no arithmetic.  The driver's MEASURED_PEAKS.json figure is a 1:1 copy; this says what a pure
// streaming kernel reaches with the scorer's 1:4 read:write ratio, i.e. the practical ceiling
// `roofline.frac` should be read against.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// tools/hbm_mix.cu -o tools/_bin/hbm_mix ; run: tools/_bin/hbm_mix [rows] [T] [n_out]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NOUT>
__global__ void mix_kernel(const float4* __restrict__ x, float4* o0, float4* o1, float4* o2, float4* o3,
                           float* t0, float* t1, float* t2, size_t n4, size_t rows) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldcs(x + i);
        if (NOUT > 0) __stcs(o0 + i, v);
        if (NOUT > 1) __stcs(o1 + i, v);
        if (NOUT > 2) __stcs(o2 + i, v);
        if (NOUT > 3) __stcs(o3 + i, v);
        if (i < rows) { t0[i] = v.x; t1[i] = v.y; t2[i] = v.z; }
    }
}

int main(int argc, char** argv) {
    const size_t rows = argc > 1 ? atoll(argv[1]) : 12800000;
    const int T = argc > 2 ? atoi(argv[2]) : 50;
    const int nout = argc > 3 ? atoi(argv[3]) : 4;
    const size_t n = rows * T, n4 = n / 4;
    float4 *x, *o[4]; float* t[3];
    CK(cudaMalloc(&x, n * 4));
    for (int i = 0; i < 4; ++i) CK(cudaMalloc(&o[i], n * 4));
    for (int i = 0; i < 3; ++i) CK(cudaMalloc(&t[i], rows * 4));
    CK(cudaMemset(x, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)n * 4 * (1 + nout) + (double)rows * 12;
    for (int blocks_per_sm = 1; blocks_per_sm <= 8; blocks_per_sm *= 2) {
        for (int threads = 256; threads <= 1024; threads *= 2) {
            const int grid = 148 * blocks_per_sm;
            float best = 1e30f;
            for (int it = 0; it < 6; ++it) {
                cudaEventRecord(e0);
                switch (nout) {
                    case 1: mix_kernel<1><<<grid, threads>>>(x, o[0], o[1], o[2], o[3], t[0], t[1], t[2], n4, rows); break;
                    case 2: mix_kernel<2><<<grid, threads>>>(x, o[0], o[1], o[2], o[3], t[0], t[1], t[2], n4, rows); break;
                    case 3: mix_kernel<3><<<grid, threads>>>(x, o[0], o[1], o[2], o[3], t[0], t[1], t[2], n4, rows); break;
                    default: mix_kernel<4><<<grid, threads>>>(x, o[0], o[1], o[2], o[3], t[0], t[1], t[2], n4, rows); break;
                }
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 2 && ms < best) best = ms;
            }
            printf("nout=%d grid=%d threads=%d  %.3f ms  %.1f GB/s\n", nout, grid, threads, best, bytes / best * 1e-6);
        }
    }
    return 0;
}
