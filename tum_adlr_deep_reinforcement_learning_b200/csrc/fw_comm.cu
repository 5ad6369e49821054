// fw_comm.cu — the PPO gradient all-reduce fused with clip_grad_norm_ + Adam in ONE kernel over NVLink peer memory.
//
// the only collective of the path is the mean of the flattened policy gradient (10.5 k floats = 42 KB) once
// per optimiser step — 80 times per PPO iteration.  At that size an NCCL all-reduce is pure latency (ring set-up,
// several kernel phases: ~25-30 us inside the captured update graph, 9 % of an iteration on 8 GPUs).  Here every rank
// owns one cudaMalloc'ed buffer that its peers map through CUDA IPC (NVLink 5 / NVSwitch: every peer is one hop away at
// full bandwidth), and the optimiser step is one launch of one block per rank:
//   1. PUSH the local gradient into slot [rank] of every peer's buffer — posted 16-byte stores over NVLink, no round
//      trip per element (a first version pulled the peers' slots with loads: every load is a ~2 us NVLink round trip and
//      the kernel took 37 us at 2 GPUs against 22 us for NCCL + divide + Adam).  Two slot sets alternate per call, so a
//      peer that runs ahead cannot overwrite what a slower rank still reads,
//   2. one system-scope fence, then release-store the call number into every peer's flag word; acquire-spin on the own
//      flag words until every peer has signalled this call,
//   3. add the world slots of the LOCAL buffer in RANK ORDER (every rank forms bit-identical sums, so the replicas never
//      drift), scale by 1 / world,
//   4. gradient-norm clip and Adam update as in adam_clip_kernel (fw_ppo.cu; stable_baselines3/ppo/ppo.py:212-214).
// No host involvement, no second launch, capturable in a CUDA graph (the call number lives in device memory).
// A rank that does not see its peers within ~2 s sets an error word instead of spinning forever.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fwb200.h"

#define FW_COMM_HEADER_BYTES 512

struct FwCommHeader {                       // at offset 0 of every rank's shared buffer
    unsigned long long flags[FW_COMM_MAX_WORLD];     // flags[r]: number of the last call rank r has signalled to this rank
    unsigned long long epoch;                        // number of calls this rank has made
    int error;                                       // 1: a peer did not arrive in time
};

struct FwComm {
    int world, rank, n, device;
    size_t slot_floats, bytes;
    char* local;                            // this rank's buffer (cudaMalloc): header | 2 sets of `world` slots
    char* peer[FW_COMM_MAX_WORLD];          // mapped buffers of every rank (peer[rank] == local)
    int connected;
};

namespace {

struct PeerTable { char* p[FW_COMM_MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {      // peer data is only valid after the acquire above
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(1024) allreduce_adam_kernel(const PeerTable peers, int world, int rank, size_t slot_floats,
                                                              float* __restrict__ param, float* __restrict__ grad,
                                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                              float* step_dev, int n, float lr, float beta1, float beta2,
                                                              float eps, float max_norm) {
    __shared__ unsigned long long s_epoch;
    __shared__ double sh[32];
    __shared__ float s_coef, s_bc1, s_bc2;
    FwCommHeader* me = reinterpret_cast<FwCommHeader*>(peers.p[rank]);
    if (threadIdx.x == 0) s_epoch = ++me->epoch;
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    // slot set (epoch & 1), inside it one slot per source rank
    const size_t set_off = FW_COMM_HEADER_BYTES + (size_t)(epoch & 1ull) * (size_t)world * slot_floats * sizeof(float);
    const int n4 = (n + 3) / 4;                       // slots are padded to 64 floats: whole float4s
    constexpr int U = 3;                               // float4s per thread per pass: all loads of a pass in flight at once
    for (int j0 = threadIdx.x; j0 < n4; j0 += U * blockDim.x) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u * blockDim.x;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n4) {
                v[u].x = grad[4 * j];
                v[u].y = 4 * j + 1 < n ? grad[4 * j + 1] : 0.f;
                v[u].z = 4 * j + 2 < n ? grad[4 * j + 2] : 0.f;
                v[u].w = 4 * j + 3 < n ? grad[4 * j + 3] : 0.f;
            }
        }
        for (int r = 0; r < world; ++r) {
            float4* dst = reinterpret_cast<float4*>(peers.p[r] + set_off + (size_t)rank * slot_floats * sizeof(float));
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (j0 + u * blockDim.x < n4) dst[j0 + u * blockDim.x] = v[u];
        }
    }
    __threadfence_system();                            // the pushed gradient is visible to the peers before the flag is
    __syncthreads();
    if ((int)threadIdx.x < world) {
        st_release_sys(&reinterpret_cast<FwCommHeader*>(peers.p[threadIdx.x])->flags[rank], epoch);
        const long long t0 = clock64();
        while (ld_acquire_sys(&me->flags[threadIdx.x]) < epoch) {
            if (clock64() - t0 > 4000000000ll) { me->error = 1; break; }      // ~2 s: report, do not hang the GPU
        }
    }
    __syncthreads();
    // mean gradient from the LOCAL slots, ranks added in rank order (identical bits on every rank); it replaces the
    // local gradient.  L1 may hold these addresses from two calls ago: read through L2 (ld.cg).
    const float inv_world = 1.f / (float)world;
    const float4* local_set = reinterpret_cast<const float4*>(peers.p[rank] + set_off);
    const size_t slot4 = slot_floats / 4;
    double ss = 0.0;
    for (int j0 = threadIdx.x; j0 < n4; j0 += U * blockDim.x) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            float4 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                t[u] = (j0 + u * blockDim.x < n4) ? __ldcg(local_set + (size_t)r * slot4 + j0 + u * blockDim.x) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) { acc[u].x += t[u].x; acc[u].y += t[u].y; acc[u].z += t[u].z; acc[u].w += t[u].w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u * blockDim.x;
            if (j >= n4) continue;
            const float g4[4] = {acc[u].x * inv_world, acc[u].y * inv_world, acc[u].z * inv_world, acc[u].w * inv_world};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * j + k < n) { grad[4 * j + k] = g4[k]; ss += (double)g4[k] * (double)g4[k]; }
        }
    }
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) {
            const float total = (float)sqrt(v);
            float coef = max_norm > 0.f ? max_norm / (total + 1e-6f) : 1.f;
            s_coef = coef < 1.f ? coef : 1.f;
            const float step = *step_dev + 1.f;
            *step_dev = step;
            s_bc1 = 1.f - powf(beta1, step);
            s_bc2 = 1.f - powf(beta2, step);
        }
    }
    __syncthreads();
    const float coef = s_coef, step_size = lr / s_bc1, inv_sqrt_bc2 = rsqrtf(s_bc2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float g = grad[i] * coef;
        const float m = beta1 * exp_avg[i] + (1.f - beta1) * g;
        const float v = beta2 * exp_avg_sq[i] + (1.f - beta2) * g * g;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
        param[i] -= step_size * (m / denom);
    }
}

thread_local char g_comm_err[256] = "";

}  // namespace

extern "C" {

const char* fw_comm_last_error(void) { return g_comm_err; }

int fw_comm_create(int32_t n_floats, int32_t world, int32_t rank, int32_t device, FwComm** out) {
    if (!out || n_floats <= 0 || world < 1 || world > FW_COMM_MAX_WORLD || rank < 0 || rank >= world) return FW_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return FW_ENODEVICE;
    FwComm* c = new FwComm();
    memset(c, 0, sizeof(*c));
    c->world = world; c->rank = rank; c->n = n_floats; c->device = device;
    c->slot_floats = ((size_t)n_floats + 63) / 64 * 64;
    c->bytes = FW_COMM_HEADER_BYTES + 2 * (size_t)world * c->slot_floats * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&c->local, c->bytes);
    if (e != cudaSuccess) { snprintf(g_comm_err, sizeof(g_comm_err), "cudaMalloc: %s", cudaGetErrorString(e)); delete c; return FW_ENOMEM; }
    cudaMemset(c->local, 0, c->bytes);
    cudaDeviceSynchronize();
    c->peer[rank] = c->local;
    *out = c;
    return FW_OK;
}

int fw_comm_export(FwComm* c, void* handle64) {
    if (!c || !handle64) return FW_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == FW_COMM_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) { snprintf(g_comm_err, sizeof(g_comm_err), "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return FW_ECUDA; }
    memcpy(handle64, &h, sizeof(h));
    return FW_OK;
}

int fw_comm_connect(FwComm* c, const void* handles) {
    if (!c || !handles) return FW_EINVAL;
    cudaSetDevice(c->device);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * FW_COMM_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            snprintf(g_comm_err, sizeof(g_comm_err), "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            return FW_ECUDA;
        }
        c->peer[r] = (char*)p;
    }
    c->connected = 1;
    return FW_OK;
}

int fw_comm_allreduce_adam(FwComm* c, float* param_dev, float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                           float* step_dev, int32_t n, float lr, float beta1, float beta2, float eps, float max_norm,
                           void* stream) {
    if (!c || !c->connected || n != c->n || !param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !step_dev) return FW_EINVAL;
    PeerTable t;
    for (int r = 0; r < FW_COMM_MAX_WORLD; ++r) t.p[r] = r < c->world ? c->peer[r] : nullptr;
    allreduce_adam_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(t, c->world, c->rank, c->slot_floats, param_dev, grad_dev,
                                                                exp_avg_dev, exp_avg_sq_dev, step_dev, n, lr, beta1, beta2, eps,
                                                                max_norm);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

int fw_comm_error(FwComm* c) {
    if (!c) return FW_EINVAL;
    FwCommHeader h;
    if (cudaMemcpy(&h, c->local, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return FW_ECUDA;
    return h.error;
}

int fw_comm_destroy(FwComm* c) {
    if (!c) return FW_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    delete c;
    return FW_OK;
}

}  // extern "C"
