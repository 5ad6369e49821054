// fw_ppo.cu — the PPO minibatch loss and its gradient with respect to the policy/value network OUTPUTS, fused.
//
// The forked SB3 PPO (stable_baselines3/ppo/ppo.py:163-218) builds the clipped surrogate loss out of ~25 elementwise /
// reduction ops on [B] and [B, 3] tensors, and autograd adds twice as many for the backward pass: at B = 32 768 every
// one of them is a launch that moves 128-400 KB, i.e. pure launch latency (the update was 88 % of a PPO iteration).
// Here the whole block between "network outputs" and "gradients of the network outputs" is three small launches:
//   ppo_adv_stats_kernel   sum and sum of squares of the minibatch advantages (f64 accumulators)
//   ppo_loss_kernel        per row: Gaussian log-prob, ratio, clipped surrogate, value error; writes dLoss/dmean [B,3]
//                          and dLoss/dvalue [B]; block-reduces the loss terms and dLoss/dlog_std
//   ppo_finish_kernel      means, entropy term, the three reported scalars
// The MLPs themselves (and their backward) stay in PyTorch, as the north star asks; ppo.FusedPPOLoss hands the
// gradients written here to autograd.  Reference formulas, with ppo.py line numbers:
//   advantages = (adv - mean) / (std + 1e-8)                      :170   (torch.std: unbiased)
//   ratio = exp(log_prob - old_log_prob)                          :173
//   policy_loss = -mean(min(adv * ratio, adv * clamp(ratio, 1 - c, 1 + c)))   :176-178
//   value_loss = mse(returns, values)                             :196
//   entropy_loss = -mean(entropy)                                 :203
//   loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss       :207
// with log_prob / entropy of the diagonal Gaussian (common/distributions.py:130-175, state-independent log_std).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fwb200.h"

namespace {

constexpr int ACT = 3;
constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid-wide sum of K doubles per thread: fixed shuffle trees inside the block, one partial per block in
// global memory, and the LAST block to arrive (a ticket counter) adds the partials in block order and stores out[0..K).
// No floating-point atomics anywhere: the same inputs give the same bits on every run, which checkpoint / resume and
// run-to-run reproducibility of training rely on.  `ticket` must be 0 on entry and is left at 0.
template <int K>
__device__ __forceinline__ void grid_accumulate(const double (&v)[K], double* partial, unsigned* ticket, double* out) {
    __shared__ double sh[K][TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) sh[k][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = lane < TPB / 32 ? sh[k][lane] : 0.0;
            s = warp_sum(s);
            if (lane == 0) partial[(size_t)blockIdx.x * K + k] = s;
        }
        if (lane == 0) {
            __threadfence();
            is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (threadIdx.x < K) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partial + (size_t)b * K + threadIdx.x);
            out[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// scratch layout of fw_ppo_loss (FW_PPO_SCRATCH_DOUBLES): acc[8] | two tickets | partials of the two reductions
constexpr int PPO_MAX_GRID = 592;
__device__ __forceinline__ unsigned* ppo_ticket(double* scratch, int which) { return reinterpret_cast<unsigned*>(scratch + 8) + which; }
__device__ __forceinline__ double* ppo_partial_a(double* scratch) { return scratch + 9; }
__device__ __forceinline__ double* ppo_partial_b(double* scratch) { return scratch + 9 + 2 * PPO_MAX_GRID; }

// acc[0] = sum adv, acc[1] = sum adv^2
__global__ void __launch_bounds__(TPB) ppo_adv_stats_kernel(const float* __restrict__ adv, int B, double* acc) {
    double v[2] = {0.0, 0.0};
    for (int i = blockIdx.x * TPB + threadIdx.x; i < B; i += gridDim.x * TPB) {
        const double a = (double)adv[i];
        v[0] += a;
        v[1] += a * a;
    }
    grid_accumulate<2>(v, ppo_partial_a(acc), ppo_ticket(acc, 0), acc);
}

// acc[2] = sum min(pl1, pl2), acc[3] = sum (ret - v)^2, acc[4..6] = sum_i g_i * ((a - m)^2 / var - 1) per action dim
__global__ void __launch_bounds__(TPB) ppo_loss_kernel(const float* __restrict__ mean, const float* __restrict__ values,
                                                       const float* __restrict__ log_std, const float* __restrict__ actions,
                                                       const float* __restrict__ old_log_prob, const float* __restrict__ adv,
                                                       const float* __restrict__ returns, int B, float clip_range,
                                                       float vf_coef, double* acc, float* __restrict__ grad_mean,
                                                       float* __restrict__ grad_values) {
    const double n = (double)B;
    const double mu = acc[0] / n;
    // unbiased variance like torch.std(); a one-row batch gives NaN there as well
    const double var_a = (acc[1] - n * mu * mu) / (n - 1.0);
    const float a_mean = (float)mu, a_den = (float)sqrt(var_a > 0.0 ? var_a : 0.0) + 1e-8f;
    float ls[ACT], inv_var[ACT];
#pragma unroll
    for (int j = 0; j < ACT; ++j) { ls[j] = log_std[j]; inv_var[j] = expf(-2.f * ls[j]); }
    const float inv_b = 1.f / (float)B;
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * TPB + threadIdx.x; i < B; i += gridDim.x * TPB) {
        float d[ACT], lp = 0.f;
#pragma unroll
        for (int j = 0; j < ACT; ++j) {
            d[j] = actions[(size_t)i * ACT + j] - mean[(size_t)i * ACT + j];
            lp += -(d[j] * d[j]) * (0.5f * inv_var[j]) - ls[j] - 0.91893853320467274178f;      // 0.5 log(2 pi)
        }
        const float an = (adv[i] - a_mean) / a_den;
        const float ratio = expf(lp - old_log_prob[i]);
        const float lo = 1.f - clip_range, hi = 1.f + clip_range;
        const float rc = fminf(fmaxf(ratio, lo), hi);
        const float pl1 = an * ratio, pl2 = an * rc;
        v[0] += (double)fminf(pl1, pl2);
        // d min(pl1, pl2) / d ratio as autograd has it: inside the clip range both branches carry the gradient (a tie,
        // split in halves, both lead to `an`); outside only the unclipped branch has one, and only where it is the min
        const bool inside = ratio >= lo && ratio <= hi;
        float g_ratio = 0.f;
        if (inside) g_ratio = an;
        else if (pl1 < pl2) g_ratio = an;
        else if (pl1 == pl2) g_ratio = 0.5f * an;
        const float g_lp = -inv_b * g_ratio * ratio;              // d policy_loss / d log_prob_i
#pragma unroll
        for (int j = 0; j < ACT; ++j) {
            grad_mean[(size_t)i * ACT + j] = g_lp * d[j] * inv_var[j];
            v[2 + j] += (double)(g_lp * (d[j] * d[j] * inv_var[j] - 1.f));
        }
        const float ve = values[i] - returns[i];
        v[1] += (double)(ve * ve);
        grad_values[i] = vf_coef * 2.f * inv_b * ve;
    }
    grid_accumulate<5>(v, ppo_partial_b(acc), ppo_ticket(acc, 1), acc + 2);
}

// out[0] = loss, out[1] = policy_loss, out[2] = value_loss; grad_log_std[j]
__global__ void ppo_finish_kernel(const double* acc, const float* log_std, int B, float ent_coef, float vf_coef,
                                  float* out, float* grad_log_std) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double n = (double)B;
    const double policy_loss = -acc[2] / n, value_loss = acc[3] / n;
    double entropy = 0.0;
    for (int j = 0; j < ACT; ++j) {
        entropy += 0.5 + 0.91893853320467274178 + (double)log_std[j];
        grad_log_std[j] = (float)(acc[4 + j] - (double)ent_coef);   // entropy_loss = -entropy: d/d log_std_j = -1
    }
    out[0] = (float)(policy_loss - (double)ent_coef * entropy + (double)vf_coef * value_loss);
    out[1] = (float)policy_loss;
    out[2] = (float)value_loss;
}

}  // namespace

extern "C" int fw_ppo_loss(const float* mean_dev, const float* values_dev, const float* log_std_dev,
                           const float* actions_dev, const float* old_log_prob_dev, const float* adv_dev,
                           const float* returns_dev, int32_t batch, float clip_range, float ent_coef, float vf_coef,
                           double* scratch_dev, float* grad_mean_dev, float* grad_values_dev, float* grad_log_std_dev,
                           float* losses_dev, void* stream) {
    if (!mean_dev || !values_dev || !log_std_dev || !actions_dev || !old_log_prob_dev || !adv_dev || !returns_dev ||
        !scratch_dev || !grad_mean_dev || !grad_values_dev || !grad_log_std_dev || !losses_dev || batch <= 0)
        return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch_dev, 0, sizeof(double) * 9, st) != cudaSuccess) return FW_ECUDA;
    int grid = (batch + TPB - 1) / TPB;
    if (grid > PPO_MAX_GRID) grid = PPO_MAX_GRID;
    ppo_adv_stats_kernel<<<grid, TPB, 0, st>>>(adv_dev, batch, scratch_dev);
    ppo_loss_kernel<<<grid, TPB, 0, st>>>(mean_dev, values_dev, log_std_dev, actions_dev, old_log_prob_dev, adv_dev,
                                          returns_dev, batch, clip_range, vf_coef, scratch_dev, grad_mean_dev,
                                          grad_values_dev);
    ppo_finish_kernel<<<1, 32, 0, st>>>(scratch_dev, log_std_dev, batch, ent_coef, vf_coef, losses_dev, grad_log_std_dev);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// Rollout glue of one env step: VecNormalize.step_wait + RunningMeanStd.update (vec_normalize.py:106-127,
// running_mean_std.py:19-39), RolloutBuffer.add (buffers.py:292-302) and the Monitor-style episode bookkeeping
// (monitor.py:99-113), which as device-tensor ops were ~70 launches per step (launch latency: 3x the simulator's
// own time at 8192 envs).  Three launches:
//   rollout_stats_kernel    ret = ret * gamma + r; batch sums of obs, obs^2, ret, ret^2; episode trackers
//   rollout_moments_kernel  Chan's parallel update of the running mean / var / count (one block)
//   rollout_apply_kernel    normalise obs and reward with the UPDATED moments, write rollout-buffer row t
//                           (previous obs / dones, this step's action, value, log-prob, reward), roll last_obs /
//                           last_dones forward, zero the finished envs' return accumulators
// All statistics are float64 like the reference; the batch variance is E[x^2] - mean^2 from f64 sums.
namespace {

// scratch layout: [0, D) sum obs_j, [D, 2D) sum obs_j^2, 2D sum ret, 2D+1 sum ret^2 | [2D+2, 3D+3) the scales of
// rollout_scales_kernel | 3D+3 the ticket | from 3D+4: FW_ROLLOUT_BLOCKS partial rows of 2D+5 sums.
// Deterministic like grid_accumulate: every warp adds its groups into its own shared-memory row in a fixed order, the
// block adds its warps in order, the last block to arrive adds the blocks in order.  No floating-point atomics.
constexpr int ROLLOUT_BLOCKS = FW_ROLLOUT_BLOCKS;
__global__ void __launch_bounds__(TPB) rollout_stats_kernel(const FwRolloutPost p) {
    extern __shared__ double sh[];                       // [TPB / 32][2 D + 5] partial sums of the warps
    __shared__ bool is_last;
    const int D = p.obs_dim, NS = 2 * D + 5;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* mine = sh + (size_t)w * NS;
    for (int k = lane; k < NS; k += 32) mine[k] = 0.0;
    __syncwarp();
    // every warp walks whole groups of 32 envs so that the shuffles below are full-warp
    for (int base = (blockIdx.x * TPB + threadIdx.x - lane); base < p.n; base += gridDim.x * TPB) {
        const int i = base + lane;
        const bool on = i < p.n;
        double r_new = 0.0, e_ret = 0.0, e_len = 0.0, e_cnt = 0.0;
        if (on) {
            const double rw = (double)p.rew_raw[i];
            r_new = p.ret[i] * (double)p.gamma + rw;
            p.ret[i] = r_new;
            const double rr = p.run_ret[i] + rw, rl = p.run_len[i] + 1.0;
            const bool d = p.done[i] != 0;
            if (d) { e_ret = rr; e_len = rl; e_cnt = 1.0; }
            p.run_ret[i] = d ? 0.0 : rr;
            p.run_len[i] = d ? 0.0 : rl;
        }
        double s;
        s = warp_sum(r_new); if (lane == 0) mine[2 * D] += s;
        s = warp_sum(r_new * r_new); if (lane == 0) mine[2 * D + 1] += s;
        s = warp_sum(e_ret); if (lane == 0) mine[2 * D + 2] += s;
        s = warp_sum(e_len); if (lane == 0) mine[2 * D + 3] += s;
        s = warp_sum(e_cnt); if (lane == 0) mine[2 * D + 4] += s;
        if (p.training) {
            for (int j = 0; j < D; ++j) {
                const double x = on ? (double)p.obs_raw[(size_t)i * D + j] : 0.0;
                s = warp_sum(x); if (lane == 0) mine[j] += s;
                s = warp_sum(x * x); if (lane == 0) mine[D + j] += s;
            }
        }
    }
    __syncthreads();
    double* partial = p.scratch + 3 * D + 4;
    unsigned* ticket = reinterpret_cast<unsigned*>(p.scratch + 3 * D + 3);
    for (int k = threadIdx.x; k < NS; k += TPB) {
        double s = 0.0;
        for (int q = 0; q < TPB / 32; ++q) s += sh[(size_t)q * NS + k];
        partial[(size_t)blockIdx.x * NS + k] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int k = threadIdx.x; k < NS; k += TPB) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partial + (size_t)b * NS + k);
            if (k < 2 * D + 2) p.scratch[k] = s;
            else p.ep_stats[k - 2 * D - 2] += s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

__device__ __forceinline__ void chan_update(double* mean, double* var, double count, double bsum, double bsumsq, double nb) {
    const double bm = bsum / nb;
    double bv = bsumsq / nb - bm * bm;
    if (bv < 0.0) bv = 0.0;
    const double delta = bm - *mean, tot = count + nb;
    const double m2 = *var * count + bv * nb + delta * delta * count * nb / tot;
    *mean = *mean + delta * nb / tot;
    *var = m2 / tot;
}

__global__ void rollout_moments_kernel(const FwRolloutPost p) {
    const int D = p.obs_dim;
    const double nb = (double)p.n;
    if (!p.training) return;
    const int j = threadIdx.x;
    const double oc = *p.obs_count, rc = *p.ret_count;
    if (j < D) chan_update(p.obs_mean + j, p.obs_var + j, oc, p.scratch[j], p.scratch[D + j], nb);
    if (j == 0) chan_update(p.ret_mean, p.ret_var, rc, p.scratch[2 * D], p.scratch[2 * D + 1], nb);
    __syncthreads();
    if (j == 0) {
        *p.obs_count = oc + nb;
        *p.ret_count = rc + nb;
    }
}

// scratch[2 D + 2 + j] = 1 / sqrt(obs_var_j + eps), scratch[3 D + 2] = 1 / sqrt(ret_var + eps): one square root per
// statistic instead of one per element (runs after the moments update, also when training is off)
__global__ void rollout_scales_kernel(const FwRolloutPost p) {
    const int D = p.obs_dim, j = threadIdx.x;
    if (j < D) p.scratch[2 * D + 2 + j] = 1.0 / sqrt(p.obs_var[j] + (double)p.epsilon);
    if (j == 0) p.scratch[3 * D + 2] = 1.0 / sqrt(*p.ret_var + (double)p.epsilon);
}

// one thread per observation element (coalesced); the threads of column 0 also move the per-env scalars
__global__ void __launch_bounds__(TPB) rollout_apply_kernel(const FwRolloutPost p) {
    const int D = p.obs_dim, A = p.act_dim;
    const size_t o = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (o >= (size_t)p.n * D) return;
    const int i = (int)(o / D), j = (int)(o - (size_t)i * D);
    p.buf_obs[o] = p.last_obs[o];
    const float raw = p.obs_raw[o];
    float v = raw;
    if (p.norm_obs) {
        double x = ((double)raw - p.obs_mean[j]) * p.scratch[2 * D + 2 + j];
        x = fmin(fmax(x, -(double)p.clip_obs), (double)p.clip_obs);
        v = (float)x;
    }
    p.last_obs[o] = v;
    if (j != 0) return;
    for (int k = 0; k < A; ++k) p.buf_actions[(size_t)i * A + k] = p.actions[(size_t)i * A + k];
    float rw = p.rew_raw[i];
    if (p.norm_reward) {
        double x = (double)rw * p.scratch[3 * D + 2];
        x = fmin(fmax(x, -(double)p.clip_reward), (double)p.clip_reward);
        rw = (float)x;
    }
    p.buf_rewards[i] = rw;
    p.buf_dones[i] = p.last_dones[i];
    p.buf_values[i] = p.values[i];
    p.buf_log_probs[i] = p.log_probs[i];
    const bool d = p.done[i] != 0;
    p.last_dones[i] = d ? 1.f : 0.f;
    if (d) p.ret[i] = 0.0;
}

}  // namespace

extern "C" int fw_rollout_post_size(void) { return (int)sizeof(FwRolloutPost); }

extern "C" int fw_rollout_post_step(const FwRolloutPost* p, void* stream) {
    if (!p || p->n <= 0 || p->obs_dim <= 0 || p->obs_dim > 256 || p->act_dim <= 0 || !p->obs_raw || !p->rew_raw ||
        !p->done || !p->actions || !p->values || !p->log_probs || !p->last_obs || !p->last_dones || !p->ret ||
        !p->obs_mean || !p->obs_var || !p->obs_count || !p->ret_mean || !p->ret_var || !p->ret_count || !p->run_ret ||
        !p->run_len || !p->ep_stats || !p->buf_obs || !p->buf_actions || !p->buf_rewards || !p->buf_dones ||
        !p->buf_values || !p->buf_log_probs || !p->scratch)
        return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = p->obs_dim;
    int grid = (p->n + TPB - 1) / TPB;
    if (grid > ROLLOUT_BLOCKS) grid = ROLLOUT_BLOCKS;
    // the ticket starts at 0 (a scratch buffer allocated zeroed) and every launch leaves it at 0
    rollout_stats_kernel<<<grid, TPB, sizeof(double) * (TPB / 32) * (2 * D + 5), st>>>(*p);
    rollout_moments_kernel<<<1, 256, 0, st>>>(*p);
    rollout_scales_kernel<<<1, 256, 0, st>>>(*p);
    rollout_apply_kernel<<<(unsigned)(((size_t)p->n * D + TPB - 1) / TPB), TPB, 0, st>>>(*p);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}

// ---------------------------------------------------------------------------------------------------------------
// clip_grad_norm_(max_norm) + Adam.step() over ONE flat parameter vector (ppo.py:212-214; torch.optim.Adam defaults
// as SB3 sets them: betas (0.9, 0.999), eps 1e-5, no weight decay).  The policy has ~10.5 k parameters in 13 tensors:
// as foreach ops that was ~25 launches per optimiser step; here one block does both (norm in f64, update in f32).
// `step_dev` is the device-resident step counter (capturable in a CUDA graph).
namespace {

__global__ void __launch_bounds__(1024) adam_clip_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                         float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                         float* step_dev, int n, float lr, float beta1, float beta2,
                                                         float eps, float max_norm) {
    __shared__ double sh[32];
    __shared__ float s_coef, s_bc1, s_bc2;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double g = (double)grad[i]; ss += g * g; }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
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

}  // namespace

extern "C" int fw_adam_clip_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                                 float* step_dev, int32_t n, float lr, float beta1, float beta2, float eps,
                                 float max_norm, void* stream) {
    if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || !step_dev || n <= 0) return FW_EINVAL;
    adam_clip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, step_dev, n,
                                                           lr, beta1, beta2, eps, max_norm);
    return cudaGetLastError() == cudaSuccess ? FW_OK : FW_ECUDA;
}
