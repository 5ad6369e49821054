// fw_gae.cu — GAE(lambda) over a time-major [T, N] rollout (RolloutBuffer.compute_returns_and_advantage,
// stable_baselines3/common/buffers.py:304-333) and the vector-pipe FMA peak micro-benchmarks.
//
// GAE is HBM bound: 12 B read + 8 B written per transition, ~8 flop.  One thread owns one env column and walks
// time backwards; at every t the warp touches 32 consecutive floats of each array (coalesced 128 B lines), the
// reverse walk is a pure streaming pattern and the f64 carry stays in a register.
// Arithmetic restates numpy's dtype promotion exactly: f32 deltas for t < T-1, an f64 delta at
// t = T-1 (bool `dones` -> float64 `1.0 - dones`), an f64 carried accumulator, f32 stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fwb200.h"
#include "fw_math.cuh"

namespace {

__global__ void gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ done,
                           const float* __restrict__ last_val, const uint8_t* __restrict__ last_done,
                           float* __restrict__ adv, float* __restrict__ ret, int T, int N, float gamma, float gl) {
    const int nidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (nidx >= N) return;
    double last_gae;
    {
        const size_t o = (size_t)(T - 1) * N + nidx;
        const double nnt = 1.0 - (double)last_done[nidx];
        const float gv = __fmul_rn(gamma, last_val[nidx]);
        const float v = val[o];
        // no contraction: the reference evaluates r + (g*v)*nnt - v with separate roundings
        const double delta = __dsub_rn(__dadd_rn((double)rew[o], __dmul_rn((double)gv, nnt)), (double)v);
        last_gae = delta;   // + coef * 0
        const float a = (float)last_gae;
        adv[o] = a;
        ret[o] = __fadd_rn(a, v);
    }
    float v_next = val[(size_t)(T - 1) * N + nidx];
    for (int t = T - 2; t >= 0; --t) {
        const size_t o = (size_t)t * N + nidx;
        const float nnt = __fsub_rn(1.0f, done[o + N]);
        const float v = val[o];
        const float d = __fsub_rn(__fadd_rn(rew[o], __fmul_rn(__fmul_rn(gamma, v_next), nnt)), v);
        const float coef = __fmul_rn(gl, nnt);
        last_gae = __dadd_rn((double)d, __dmul_rn((double)coef, last_gae));
        const float a = (float)last_gae;
        adv[o] = a;
        ret[o] = __fadd_rn(a, v);
        v_next = v;
    }
}

// diagnostic: evaluates the hot-loop math kernels of fw_math.cuh elementwise (tests/test_gpu_math.py)
__global__ void debug_math_kernel(int op, const double* x, const double* y, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) out[i] = fw::exp_bf(x[i]);
    else if (op == 1) out[i] = fw::asin_bf(x[i]);
    else out[i] = fw::atan2_bf(y[i], x[i]);
}

template <typename T, int ILP>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = (T)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = acc[k] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    if (s == (T)123456789) out[0] = s;   // never true; keeps the loop alive
}

template <typename T>
int measure_peak(double* tflops_out) {
    T* d = nullptr;
    if (cudaMalloc(&d, sizeof(T)) != cudaSuccess) return FW_ECUDA;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ILP = 8, threads = 256, blocks = sms * 8, iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<T, ILP><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return FW_ECUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * ILP * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return FW_OK;
}

}  // namespace

extern "C" {

int fw_gae(const float* rew_dev, const float* val_dev, const float* done_dev, const float* last_val_dev,
           const uint8_t* last_done_dev, float* adv_dev, float* ret_dev, int32_t T, int32_t N, float gamma,
           float gae_lambda, void* stream) {
    if (!rew_dev || !val_dev || !done_dev || !last_val_dev || !last_done_dev || !adv_dev || !ret_dev || T <= 0 || N <= 0)
        return FW_EINVAL;
    // python-float gamma * gae_lambda is formed in double, then cast to the float32 array dtype (numpy weak scalar)
    const float gl = (float)((double)gamma * (double)gae_lambda);
    const int bs = 128, grid = (N + bs - 1) / bs;
    gae_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(rew_dev, val_dev, done_dev, last_val_dev, last_done_dev, adv_dev,
                                                     ret_dev, T, N, gamma, gl);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

int fw_debug_math(int32_t op, const double* x_dev, const double* y_dev, double* out_dev, int32_t n, void* stream) {
    if (!x_dev || !out_dev || n <= 0 || op < 0 || op > 2 || (op == 2 && !y_dev)) return FW_EINVAL;
    debug_math_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, x_dev, y_dev, out_dev, n);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

int fw_measure_fma_peak(int32_t device, int32_t precision, double* tflops_out) {
    if (!tflops_out) return FW_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return FW_ENODEVICE;
    return precision == FW_F64 ? measure_peak<double>(tflops_out) : measure_peak<float>(tflops_out);
}

}  // extern "C"
