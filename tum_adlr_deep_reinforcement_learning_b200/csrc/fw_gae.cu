// fw_gae.cu — GAE(lambda) over a time-major [T, N] rollout (RolloutBuffer.compute_returns_and_advantage,
// stable_baselines3/common/buffers.py:304-333) and the vector-pipe FMA peak micro-benchmarks.
//
// GAE is HBM bound: 12 B read + 8 B written per transition, ~8 flop.  The recurrence A[t] = d[t] + c[t] A[t+1] has to
// be evaluated sequentially per env to stay bit-exact (a parallel scan re-associates the f64 carry), so the
// parallelism is over envs only — 8192 dependent walks of T steps — and the bandwidth has to come from memory-level
// parallelism INSIDE each walk: gae_stream_kernel gives every one-warp block a strip of SW env columns and streams
// [TT rows x SW columns] tiles of rew / val / done through a STAGES-deep cp.async ring in shared memory (the whole
// warp issues the 16-byte copies, lanes < SW run the scan out of shared memory), so ~10-20 MB of loads are in flight
// across the GPU while every lane's carry chain (one DMUL + one DADD per step) runs at its own pace.  gae_kernel (one
// thread per column, plain loads: one exposed DRAM round trip per step) remains for shapes the tiles do not fit.
// Arithmetic restates numpy's dtype promotion exactly: f32 deltas for t < T-1, an f64 delta at
// t = T-1 (bool `dones` -> float64 `1.0 - dones`), an f64 carried accumulator, f32 stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// One warp per strip of SW <= 32 env columns (N % SW == 0, 16-byte aligned rows).  Tile k holds the rows t = T-2-k*TT
// downwards: rew[t], val[t] and done[t+1] (the flag the step at t multiplies with).  The arithmetic is gae_kernel's,
// statement for statement.  A single warp's instruction stream sets the pace of a strip, so the scan is software
// pipelined by hand in chunks of R rows: `front` does everything that does not depend on the carry (shared-memory
// reads, the f32 delta, the conversions) for R rows, `chain` then runs one DMUL + one DADD per row with the stores
// hanging off it, and the front of chunk c+1 sits in the same basic block as the chain of chunk c, so the compiler
// fills the chain's dependency stalls with it.  No branch and no 64-bit multiply per row.
template <int SW, int TT, int STAGES, int R>
__global__ void __launch_bounds__(32) gae_stream_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                                        const float* __restrict__ done, const float* __restrict__ last_val,
                                                        const uint8_t* __restrict__ last_done, float* __restrict__ adv,
                                                        float* __restrict__ ret, int T, int N, float gamma, float gl) {
    static_assert(SW <= 32 && TT % (2 * R) == 0, "strip / tile shape");
    extern __shared__ float4 gae_smem4[];
    float* sm = reinterpret_cast<float*>(gae_smem4);
    constexpr int CPR = SW / 4;                    // 16-byte chunks per tile row
    constexpr int STAGE_FLOATS = 3 * TT * SW;
    const int lane = threadIdx.x;
    const int n0 = blockIdx.x * SW;
    const int ntiles = (T - 1 + TT - 1) / TT;      // rows T-2 .. 0

    auto issue = [&](int k) {
        if (k < ntiles) {
            const int t_hi = T - 2 - k * TT;
            const int rows = t_hi + 1 < TT ? t_hi + 1 : TT;
            float* base = sm + (k % STAGES) * STAGE_FLOATS;
            for (int c = lane; c < rows * CPR; c += 32) {
                const int r = c / CPR, q = c - r * CPR;
                const size_t g = (size_t)(t_hi - r) * N + n0 + q * 4;
                cp_async16(base + (0 * TT + r) * SW + q * 4, rew + g);
                cp_async16(base + (1 * TT + r) * SW + q * 4, val + g);
                cp_async16(base + (2 * TT + r) * SW + q * 4, done + g + N);
            }
        }
        cp_async_commit();                         // an empty group keeps the group count uniform
    };
#pragma unroll
    for (int k = 0; k < STAGES - 1; ++k) issue(k);

    const bool mine = lane < SW;
    const int col0 = mine ? lane : 0;
    double last_gae;
    float v_next;
    {
        const int nidx = n0 + col0;
        const size_t o = (size_t)(T - 1) * N + nidx;
        const double nnt = 1.0 - (double)last_done[nidx];
        const float gv = __fmul_rn(gamma, last_val[nidx]);
        const float v = val[o];
        const double delta = __dsub_rn(__dadd_rn((double)rew[o], __dmul_rn((double)gv, nnt)), (double)v);
        last_gae = delta;
        const float a = (float)delta;
        if (mine) { adv[o] = a; ret[o] = __fadd_rn(a, v); }
        v_next = v;
    }
    for (int k = 0; k < ntiles; ++k) {
        issue(k + STAGES - 1);
        cp_async_wait<STAGES - 1>();               // this thread's copies of tile k have landed ...
        __syncwarp();                              // ... and so have the other lanes'
        const int t_hi = T - 2 - k * TT;
        const int rows = t_hi + 1 < TT ? t_hi + 1 : TT;
        const float* base = sm + (k % STAGES) * STAGE_FLOATS + col0;
        if (mine) {
            float* pa = adv + (size_t)t_hi * N + n0 + col0;      // row r of the tile sits at pa[-r * N]
            float* pr = ret + (size_t)t_hi * N + n0 + col0;
            auto front = [&](int r0, float vprev, float (&vv)[R], double (&dd)[R], double (&cc)[R]) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int r = r0 + i;
                    const float nnt = __fsub_rn(1.0f, base[(2 * TT + r) * SW]);
                    vv[i] = base[(1 * TT + r) * SW];
                    const float vn = (i == 0) ? vprev : vv[i > 0 ? i - 1 : 0];
                    const float d = __fsub_rn(__fadd_rn(base[r * SW], __fmul_rn(__fmul_rn(gamma, vn), nnt)), vv[i]);
                    dd[i] = (double)d;
                    cc[i] = (double)__fmul_rn(gl, nnt);
                }
            };
            auto chain = [&](int off, const float (&vv)[R], const double (&dd)[R], const double (&cc)[R]) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    last_gae = __dadd_rn(dd[i], __dmul_rn(cc[i], last_gae));
                    const float a = (float)last_gae;
                    pa[off] = a;
                    pr[off] = __fadd_rn(a, vv[i]);
                    off -= N;
                }
            };
            if (rows == TT) {
                float vvA[R], vvB[R];
                double ddA[R], ccA[R], ddB[R], ccB[R];
                front(0, v_next, vvA, ddA, ccA);
#pragma unroll
                for (int c = 0; c < TT / R; c += 2) {
                    front((c + 1) * R, vvA[R - 1], vvB, ddB, ccB);
                    chain(-(c * R) * N, vvA, ddA, ccA);
                    if (c + 2 < TT / R) front((c + 2) * R, vvB[R - 1], vvA, ddA, ccA);
                    chain(-((c + 1) * R) * N, vvB, ddB, ccB);
                }
                v_next = vvB[R - 1];
            } else {                               // the ragged last tile (rows 0 .. of the rollout)
                int off = 0;
#pragma unroll 1
                for (int r = 0; r < rows; ++r) {
                    const float nnt = __fsub_rn(1.0f, base[(2 * TT + r) * SW]);
                    const float v = base[(1 * TT + r) * SW];
                    const float d = __fsub_rn(__fadd_rn(base[r * SW], __fmul_rn(__fmul_rn(gamma, v_next), nnt)), v);
                    const float coef = __fmul_rn(gl, nnt);
                    last_gae = __dadd_rn((double)d, __dmul_rn((double)coef, last_gae));
                    const float a = (float)last_gae;
                    pa[off] = a;
                    pr[off] = __fadd_rn(a, v);
                    off -= N;
                    v_next = v;
                }
            }
        }
        __syncwarp();                              // the next issue() refills the stage just consumed
    }
}

template <int SW, int TT, int STAGES, int R = 8>
static cudaError_t launch_gae_stream(const float* rew, const float* val, const float* done, const float* last_val,
                                     const uint8_t* last_done, float* adv, float* ret, int T, int N, float gamma, float gl,
                                     cudaStream_t st) {
    const size_t smem = (size_t)STAGES * 3 * TT * SW * sizeof(float);
    auto k = gae_stream_kernel<SW, TT, STAGES, R>;
    static bool attr_set = false;                  // per process; the value never changes
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    k<<<N / SW, 32, smem, st>>>(rew, val, done, last_val, last_done, adv, ret, T, N, gamma, gl);
    return cudaGetLastError();
}

// diagnostic: evaluates the hot-loop math kernels of fw_math.cuh elementwise (tests/test_gpu_math.py)
__global__ void debug_math_kernel(int op, const double* x, const double* y, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) out[i] = fw::exp_bf(x[i]);
    else if (op == 1) out[i] = fw::asin_bf(x[i]);
    else if (op == 3) out[i] = fw::rsqrt_fast(x[i]);
    else if (op == 4) out[i] = fw::exp_bf<false>(x[i]);            // 4..6: the immediate-coefficient instantiations
    else if (op == 5) out[i] = fw::asin_bf<false>(x[i]);
    else if (op == 6) out[i] = fw::atan2_bf<false>(y[i], x[i]);
    else if (op == 7) out[i] = fw::log_bf<true>(x[i]);
    else if (op == 8) out[i] = fw::pow_hot_bf<true>(x[i], y[i]);              // x^y
    else if (op == 9) out[i] = fw::pow_hot_bf<false>(x[i], y[i]);
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
    cudaStream_t st = (cudaStream_t)stream;
    // strip width: the widest of 32 / 16 / 8 columns that still yields >= 2.5 one-warp blocks per SM (each strip is one
    // sequential walk whose pace a single warp's instruction stream sets, so the bandwidth comes from the number of
    // strips in flight; wider rows give longer DRAM bursts).  Measured at T = 2048 on one B200: N = 8192 -> 16 columns
    // 68 us (4.9 TB/s), 32 columns 74 us, 8 columns 101 us; N = 65536 -> 32 columns 5.7 TB/s.  FWB200_GAE_SW = 8 | 16 | 32
    // forces a width, 1 the plain one-thread-per-column kernel (experiments).
    static int sms = 0, forced = -1;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char* e = getenv("FWB200_GAE_SW");
        forced = e ? atoi(e) : 0;
    }
    const bool aligned = (N % 4 == 0) && T >= 2 &&
                         ((((uintptr_t)rew_dev | (uintptr_t)val_dev | (uintptr_t)done_dev) & 15) == 0);
    int sw = 0;
    if (aligned) {
        for (int cand = 32; cand >= 8; cand >>= 1)
            if (N % cand == 0 && (2 * (N / cand) >= 5 * sms || cand == 8)) { sw = cand; break; }
        if (forced == 1) sw = 0;
        else if ((forced == 8 || forced == 16 || forced == 32) && N % forced == 0) sw = forced;
    }
    cudaError_t e;
#define GAE_ARGS rew_dev, val_dev, done_dev, last_val_dev, last_done_dev, adv_dev, ret_dev, T, N, gamma, gl, st
    if (sw == 32) e = launch_gae_stream<32, 32, 4, 8>(GAE_ARGS);
    else if (sw == 16) e = launch_gae_stream<16, 64, 4, 8>(GAE_ARGS);
    else if (sw == 8) e = launch_gae_stream<8, 64, 4, 8>(GAE_ARGS);
    else {
        const int bs = 128, grid = (N + bs - 1) / bs;
        gae_kernel<<<grid, bs, 0, st>>>(rew_dev, val_dev, done_dev, last_val_dev, last_done_dev, adv_dev, ret_dev, T, N, gamma, gl);
        e = cudaGetLastError();
    }
#undef GAE_ARGS
    return e == cudaSuccess ? FW_OK : FW_ECUDA;
}

int fw_debug_math(int32_t op, const double* x_dev, const double* y_dev, double* out_dev, int32_t n, void* stream) {
    if (!x_dev || !out_dev || n <= 0 || op < 0 || op > 9 || ((op == 2 || op == 6 || op == 8 || op == 9) && !y_dev)) return FW_EINVAL;
    debug_math_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, x_dev, y_dev, out_dev, n);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

int fw_measure_fma_peak(int32_t device, int32_t precision, double* tflops_out) {
    if (!tflops_out) return FW_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return FW_ENODEVICE;
    return precision == FW_F64 ? measure_peak<double>(tflops_out) : measure_peak<float>(tflops_out);
}

}  // extern "C"
