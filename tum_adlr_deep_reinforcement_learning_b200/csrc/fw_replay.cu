// fw_replay.cu — the off-policy replay ring of the SAC path in HBM: batched insert of one transition
// per env and Philox-sampled minibatch gather with normalisation at SAMPLE time.
//
// Reference (stable_baselines3/common/buffers.py:146-256, single env): five numpy arrays observations / next_observations
// / actions / rewards / dones of [buffer_size, 1, ...]; add() writes row `pos` and advances it (wrapping, `full` flag);
// sample() draws batch_size uniform indices in [0, size) and returns the rows, observations and rewards passed through
// VecNormalize.normalize_obs / normalize_reward with the statistics CURRENT at sample time (:245-254 — the ring holds the
// original observations and rewards, off_policy_algorithm.py:430-436).
//
// Here a transition is ONE contiguous row  obs[D] | next_obs[D] | action[A] | reward | done  padded to a multiple of four
// floats, so a sampled transition is one 16-byte-aligned burst (144 B for D = 14, A = 3) instead of five scattered
// sectors, and an insert of N envs is N consecutive rows.  Ring head, fill level and the sample counter live on the
// device, so both operations can sit inside a captured CUDA graph.  Both are HBM / latency bound: 4 B moved per float, no
// arithmetic to speak of.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fwb200.h"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// one thread per float4 of the n new rows; row i goes to ring slot (head + i) mod capacity
__global__ void __launch_bounds__(256) replay_insert_kernel(const FwReplay rb, const float* __restrict__ obs,
                                                            const float* __restrict__ next_obs, const float* __restrict__ act,
                                                            const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                            int n) {
    const int W4 = rb.row_floats / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * W4) return;
    const int i = (int)(t / W4), q = (int)(t - (long long)i * W4);
    const int D = rb.obs_dim, A = rb.act_dim;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = q * 4 + k;                 // column of the packed row
        float x = 0.f;
        if (c < D) x = obs[(size_t)i * D + c];
        else if (c < 2 * D) x = next_obs[(size_t)i * D + (c - D)];
        else if (c < 2 * D + A) x = act[(size_t)i * A + (c - 2 * D)];
        else if (c == 2 * D + A) x = rew[i];
        else if (c == 2 * D + A + 1) x = done[i] ? 1.f : 0.f;
        v[k] = x;
    }
    const long long slot = (*rb.head_dev + i) % rb.capacity;
    reinterpret_cast<float4*>(rb.rows + (size_t)slot * rb.row_floats)[q] = make_float4(v[0], v[1], v[2], v[3]);
}

// after the insert (stream order): head += n (mod capacity), size = min(size + n, capacity)
__global__ void replay_advance_kernel(const FwReplay rb, int n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *rb.head_dev = (*rb.head_dev + n) % rb.capacity;
    const long long s = *rb.size_dev + n;
    *rb.size_dev = s < rb.capacity ? s : rb.capacity;
}

// one thread per float4 of the batch rows; sample b reads ring row floor(u_b * size), u_b from Philox(seed; b, call number)
__global__ void __launch_bounds__(256) replay_sample_kernel(const FwReplay rb, const FwReplayNorm nm, int batch,
                                                            unsigned long long seed, float* __restrict__ obs_out,
                                                            float* __restrict__ act_out, float* __restrict__ next_obs_out,
                                                            float* __restrict__ done_out, float* __restrict__ rew_out,
                                                            long long* __restrict__ idx_out) {
    const int W4 = rb.row_floats / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * W4) return;
    const int b = (int)(t / W4), q = (int)(t - (long long)b * W4);
    const unsigned long long call = (unsigned long long)*rb.sample_calls_dev;
    uint32_t c0 = (uint32_t)b, c1 = (uint32_t)call, c2 = (uint32_t)(call >> 32), c3 = 0x5A4Du;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const long long size = *rb.size_dev;
    // 53-bit uniform in [0, 1) -> index in [0, size): unbiased to 2^-53 (np.random.randint(0, size), buffers.py:98-100)
    const double u = (double)((((unsigned long long)c0) << 21) ^ (((unsigned long long)c1) >> 11)) * (1.0 / 9007199254740992.0);
    long long idx = (long long)(u * (double)size);
    if (idx >= size) idx = size - 1;
    if (idx < 0) idx = 0;
    if (q == 0 && idx_out) idx_out[b] = idx;
    const float4 r4 = reinterpret_cast<const float4*>(rb.rows + (size_t)idx * rb.row_floats)[q];
    const float v[4] = {r4.x, r4.y, r4.z, r4.w};
    const int D = rb.obs_dim, A = rb.act_dim;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = q * 4 + k;
        float x = v[k];
        if (c < 2 * D) {
            const int j = c < D ? c : c - D;
            if (nm.norm_obs) {        // VecNormalize.normalize_obs (vec_normalize.py:150-163): float64 statistics, clip, float32
                double y = ((double)x - nm.obs_mean[j]) / sqrt(nm.obs_var[j] + (double)nm.epsilon);
                y = fmin(fmax(y, -(double)nm.clip_obs), (double)nm.clip_obs);
                x = (float)y;
            }
            (c < D ? obs_out : next_obs_out)[(size_t)b * D + j] = x;
        } else if (c < 2 * D + A) {
            act_out[(size_t)b * A + (c - 2 * D)] = x;
        } else if (c == 2 * D + A) {
            if (nm.norm_reward) {     // normalize_reward (vec_normalize.py:165-172)
                double y = (double)x / sqrt(*nm.ret_var + (double)nm.epsilon);
                y = fmin(fmax(y, -(double)nm.clip_reward), (double)nm.clip_reward);
                x = (float)y;
            }
            rew_out[b] = x;
        } else if (c == 2 * D + A + 1) {
            done_out[b] = x;
        }
    }
}

__global__ void replay_count_kernel(const FwReplay rb) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *rb.sample_calls_dev += 1;
}

bool replay_ok(const FwReplay* rb) {
    return rb && rb->rows && rb->head_dev && rb->size_dev && rb->sample_calls_dev && rb->capacity > 0 && rb->obs_dim > 0 &&
           rb->act_dim > 0 && rb->row_floats == FW_REPLAY_ROW_FLOATS(rb->obs_dim, rb->act_dim);
}

}  // namespace

extern "C" int fw_replay_size(void) { return (int)sizeof(FwReplay); }
extern "C" int fw_replay_norm_size(void) { return (int)sizeof(FwReplayNorm); }

extern "C" int fw_replay_insert(const FwReplay* rb, const float* obs_dev, const float* next_obs_dev, const float* actions_dev,
                                const float* rewards_dev, const uint8_t* dones_dev, int32_t n, void* stream) {
    if (!replay_ok(rb) || !obs_dev || !next_obs_dev || !actions_dev || !rewards_dev || !dones_dev || n <= 0 || n > rb->capacity)
        return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long work = (long long)n * (rb->row_floats / 4);
    replay_insert_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(*rb, obs_dev, next_obs_dev, actions_dev, rewards_dev,
                                                                          dones_dev, n);
    replay_advance_kernel<<<1, 32, 0, st>>>(*rb, n);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

extern "C" int fw_replay_sample(const FwReplay* rb, const FwReplayNorm* norm, int32_t batch, uint64_t seed, float* obs_out_dev,
                                float* actions_out_dev, float* next_obs_out_dev, float* dones_out_dev, float* rewards_out_dev,
                                int64_t* indices_out_dev, void* stream) {
    if (!replay_ok(rb) || !norm || batch <= 0 || !obs_out_dev || !actions_out_dev || !next_obs_out_dev || !dones_out_dev ||
        !rewards_out_dev)
        return FW_EINVAL;
    if ((norm->norm_obs && (!norm->obs_mean || !norm->obs_var)) || (norm->norm_reward && !norm->ret_var)) return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const long long work = (long long)batch * (rb->row_floats / 4);
    replay_sample_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(*rb, *norm, batch, (unsigned long long)seed,
                                                                          obs_out_dev, actions_out_dev, next_obs_out_dev,
                                                                          dones_out_dev, rewards_out_dev,
                                                                          (long long*)indices_out_dev);
    replay_count_kernel<<<1, 32, 0, st>>>(*rb);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}
