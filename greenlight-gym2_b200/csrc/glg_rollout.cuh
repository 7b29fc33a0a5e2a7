// glg_rollout.cuh -- what stands between the env step and a PPO update in the reference's training loop, as kernels
// (SURVEY 8f-3): VecNormalize's running observation / return statistics and normalisation (gl_gym/RL/utils.py:60-67,
// RL/experiment_manager.py:142-147, stable-baselines3 2.6.0 common/vec_env/vec_normalize.py + running_mean_std.py), the
// on-policy rollout buffer, and generalised advantage estimation (SB3 common/buffers.py RolloutBuffer, configs/agents/ppo.yml
// gamma / gae_lambda).  The env step writes raw observations [B][obs_dim] f32 and rewards [B] f64; one glg_rollout_store then
//   1. glg_roll_moments_kernel : per-CTA partial sums of (x - mean) and (x - mean)^2 per observation column over the CTA's rows
//                                (shifted by the running mean: no cancellation), plus the same for the discounted returns
//                                ret <- ret * gamma + reward of the CTA's rows                      [reads obs once]
//   2. glg_roll_finish_kernel  : per column, the partials summed in a fixed order (deterministic) and merged into the running
//                                (mean, var, count) by Chan's parallel update; writes mean and 1/sqrt(var + eps)
//   3. glg_roll_apply_kernel   : normalise + clip the observations straight into the rollout buffer slot [t+1][B][obs_dim],
//                                normalise + clip the rewards into [t][B], episode_starts[t+1] = done, ret[done] = 0
//                                                                                                     [reads obs once, writes once]
// HBM traffic per store at B = 65 536, obs_dim = 263: 2 x 69 MB read + 69 MB written = 32 us at the measured 6.4 TB/s; the eager
// torch version of round 1 took 4.5 ms per step.
//   glg_roll_gae_kernel        : one thread per env, backwards over the T steps (coalesced across envs):
//                                delta = r_t + gamma V_{t+1} nnt - V_t ; A_t = delta + gamma lambda nnt A_{t+1} ; R_t = A_t + V_t
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int GLG_ROLL_ROWS = 64;   // rows (envs) per CTA in the moments / apply kernels
constexpr int GLG_ROLL_NT = 256;

struct GlgRollArgs {
    int B, obs_dim, n_blocks, training, norm_obs, norm_reward;
    double gamma, clip_obs, clip_reward, epsilon;
    const float *obs;           // [B][obs_dim] raw observations of the step
    const double *reward;       // [B] raw rewards (NULL right after reset: no reward yet)
    const unsigned char *done;  // [B]
    double *ret;                // [B] discounted return accumulators (VecNormalize.returns)
    double *stat;               // [obs_dim + 1][3]: running mean, var, count per column; last row = return statistics
    double *partial;            // [n_blocks][obs_dim + 1][2]
    double *norm64;             // [obs_dim + 1][2]: mean, 1/sqrt(var + eps) as the apply kernel reads them
    float *obs_out;             // rollout slot [B][obs_dim]
    float *reward_out;          // rollout slot [B] (NULL right after reset)
    float *starts_out;          // episode_starts slot [B] (1.0 where the env just finished / was reset)
};

__global__ void __launch_bounds__(GLG_ROLL_NT) glg_roll_moments_kernel(const GlgRollArgs A) {
    const int r0 = blockIdx.x * GLG_ROLL_ROWS, r1 = min(r0 + GLG_ROLL_ROWS, A.B);
    for (int c = threadIdx.x; c < A.obs_dim; c += GLG_ROLL_NT) {
        const double m = A.stat[3 * c];
        double s1 = 0.0, s2 = 0.0;
        const float *col = A.obs + (size_t)r0 * A.obs_dim + c;
        int r = r0;
        for (; r + 8 <= r1; r += 8, col += 8 * (size_t)A.obs_dim) {  // eight independent loads in flight per thread
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(col + (size_t)k * A.obs_dim);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double d = (double)v[k] - m;
                s1 += d;
                s2 = fma(d, d, s2);
            }
        }
        for (; r < r1; ++r, col += A.obs_dim) {
            const double d = (double)__ldg(col) - m;
            s1 += d;
            s2 = fma(d, d, s2);
        }
        A.partial[((size_t)blockIdx.x * (A.obs_dim + 1) + c) * 2] = s1;
        A.partial[((size_t)blockIdx.x * (A.obs_dim + 1) + c) * 2 + 1] = s2;
    }
    // discounted returns of this CTA's rows: warp 0 updates them and reduces their shifted moments in lane order
    if (threadIdx.x < 32) {
        double s1 = 0.0, s2 = 0.0;
        if (A.reward && A.training) {  // VecNormalize._update_reward runs only while training
            const double m = A.stat[3 * A.obs_dim];
            for (int r = r0 + (int)threadIdx.x; r < r1; r += 32) {
                const double v = A.ret[r] * A.gamma + A.reward[r];
                A.ret[r] = v;
                const double d = v - m;
                s1 += d;
                s2 = fma(d, d, s2);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        if (threadIdx.x == 0) {
            A.partial[((size_t)blockIdx.x * (A.obs_dim + 1) + A.obs_dim) * 2] = s1;
            A.partial[((size_t)blockIdx.x * (A.obs_dim + 1) + A.obs_dim) * 2 + 1] = s2;
        }
    }
}

// one warp per column (+ the return statistic): lane l sums the partials of CTAs l, l + 32, ... in order, the lanes are combined
// by a fixed shuffle tree (deterministic), lane 0 does the Chan update (running_mean_std.py).  (One THREAD per column walking
// all partials serially was 250 us of dependent L2 loads at 1024 CTAs.)
__global__ void __launch_bounds__(GLG_ROLL_NT) glg_roll_finish_kernel(const GlgRollArgs A) {
    const int c = blockIdx.x * (GLG_ROLL_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c > A.obs_dim) return;
    const bool is_ret = c == A.obs_dim;
    double mean = A.stat[3 * c], var = A.stat[3 * c + 1], count = A.stat[3 * c + 2];
    const bool update = A.training && (is_ret ? A.reward != nullptr : A.norm_obs != 0);
    if (update) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll 4
        for (int b = lane; b < A.n_blocks; b += 32) {
            s1 += A.partial[((size_t)b * (A.obs_dim + 1) + c) * 2];
            s2 += A.partial[((size_t)b * (A.obs_dim + 1) + c) * 2 + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, o);
            s2 += __shfl_down_sync(0xffffffffu, s2, o);
        }
        const double n = (double)A.B;
        const double dmean = s1 / n;                            // batch mean - running mean
        const double bvar = fmax(s2 / n - dmean * dmean, 0.0);  // population variance of the batch
        const double tot = count + n;
        const double m2 = var * count + bvar * n + dmean * dmean * count * n / tot;
        mean = mean + dmean * n / tot;
        var = m2 / tot;
        count = tot;
        if (lane == 0) {
            A.stat[3 * c] = mean;
            A.stat[3 * c + 1] = var;
            A.stat[3 * c + 2] = count;
        }
    }
    if (lane == 0) {
        A.norm64[2 * c] = mean;
        A.norm64[2 * c + 1] = 1.0 / sqrt(var + A.epsilon);
    }
}

__global__ void __launch_bounds__(GLG_ROLL_NT) glg_roll_apply_kernel(const GlgRollArgs A) {
    const int r0 = blockIdx.x * GLG_ROLL_ROWS, r1 = min(r0 + GLG_ROLL_ROWS, A.B);
    for (int c = threadIdx.x; c < A.obs_dim; c += GLG_ROLL_NT) {
        const double m = A.norm64[2 * c], is = A.norm64[2 * c + 1];
        const float *col = A.obs + (size_t)r0 * A.obs_dim + c;
        float *out = A.obs_out + (size_t)r0 * A.obs_dim + c;
        int r = r0;
        for (; r + 8 <= r1; r += 8, col += 8 * (size_t)A.obs_dim, out += 8 * (size_t)A.obs_dim) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(col + (size_t)k * A.obs_dim);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (A.norm_obs) v[k] = (float)fmin(fmax(((double)v[k] - m) * is, -A.clip_obs), A.clip_obs);
                out[(size_t)k * A.obs_dim] = v[k];
            }
        }
        for (; r < r1; ++r, col += A.obs_dim, out += A.obs_dim) {
            float y = __ldg(col);
            if (A.norm_obs) y = (float)fmin(fmax(((double)y - m) * is, -A.clip_obs), A.clip_obs);
            *out = y;
        }
    }
    if (threadIdx.x < GLG_ROLL_ROWS && r0 + (int)threadIdx.x < r1) {
        const int r = r0 + threadIdx.x;
        const bool dn = A.done && A.done[r];
        if (A.reward && A.reward_out) {
            double v = A.reward[r];
            if (A.norm_reward) v = fmin(fmax(v * A.norm64[2 * A.obs_dim + 1], -A.clip_reward), A.clip_reward);
            A.reward_out[r] = (float)v;
        }
        if (A.starts_out) A.starts_out[r] = (dn || !A.reward) ? 1.0f : 0.0f;  // after reset every env starts an episode
        if (dn || !A.reward) A.ret[r] = 0.0;
    }
}

// RolloutBuffer.compute_returns_and_advantage (SB3 common/buffers.py): values [T+1][B] (row T = value of the observation after the
// last step), episode_starts [T+1][B] (row t+1 = done flags of step t), rewards [T][B] -> advantages, returns [T][B]
__global__ void glg_roll_gae_kernel(int T, int B, double gamma, double lam, const float *rewards, const float *values, const float *starts,
                                    float *adv, float *ret) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float last = 0.0f;
    const float g = (float)gamma, gl = (float)(gamma * lam);
    for (int t = T - 1; t >= 0; --t) {
        const float nnt = 1.0f - starts[(size_t)(t + 1) * B + b];
        const float v = values[(size_t)t * B + b];
        const float delta = rewards[(size_t)t * B + b] + g * values[(size_t)(t + 1) * B + b] * nnt - v;
        last = delta + gl * nnt * last;
        adv[(size_t)t * B + b] = last;
        ret[(size_t)t * B + b] = last + v;
    }
}
