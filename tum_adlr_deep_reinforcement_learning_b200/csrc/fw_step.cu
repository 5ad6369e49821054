// fw_step.cu — the env-step kernels, the reset kernels and the C ABI of libfwb200.so (see include/fwb200.h).
//
// One VecEnv.step over every env of the handle = three launches on the caller's stream (exact mode):
//   rk45_init_kernel     action scaling -> command constraint -> f0 = fun(t0, y0) and scipy's select_initial_step
//   rk45_attempt_kernel  the RK45 attempt loop; persistent lanes pull envs from a queue, so every lane of every warp
//                        executes the same RHS row all the time (8 warps/SM: 240 registers + 720 B of stage storage)
//   head_kernel          post-step commit (quaternion renormalisation, Euler angles, Va/alpha/beta, constraint
//                        checks) -> Dryden filter advance -> goal ring / streak test -> reward -> target law ->
//                        observation -> termination -> streamed episode metrics -> episode-end rows for the host ->
//                        auto-reset = warp-cooperative copy of the env's precomputed next-episode row
// plus one on the handle's high-priority side stream:
//   refill_kernel        FixedWingAircraft.reset (Philox) for the rows consumed at this step; overlaps the next step
// The once-per-step head code is large and cold; in its own kernel it runs at high occupancy instead of stalling
// the register-heavy integrator warps on instruction fetch (profiles/r01_*.txt).  Fixed-step modes use
// rk4_kernel -> head_kernel.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>

#include "fw_device.cuh"

namespace fw {

struct StepIO {
    const void* actions;   // [n,3] f32 or f64
    int actions_f64;
    float* obs;            // [n,14]
    float* rew;            // [n]
    uint8_t* done;         // [n]
    float* term_obs;       // [n,14] nullable
    double* obs64;         // nullable
    double* rew64;         // nullable
    int auto_reset;
    int random_actions;    // 1: draw U(-1,1)^3 from Philox (fw_step_random)
    unsigned long long action_seed;
    unsigned long long action_step;
    double* info;          // fw_set_info_rows: [0] = int32 count of this step, then cap rows of FW_INFO_HEAD + obs_dim
    int info_cap;
};

// Precomputed next-episode rows.  FixedWingAircraft.reset depends only on (seed, env id, episode number) — never on the
// trajectory that just ended — so the complete post-reset row of every env (state SoA row, first error-ring entry, reset
// observation) is computed AHEAD of time into a second SoA.  When an episode ends, head_kernel copies that row into
// place (a few hundred independent loads/stores) instead of running the ~5 k-instruction reset on one lane while the
// 31 other lanes of the warp — and, because one wave of warps runs the kernel, the whole step — wait for it; the
// consumed rows are recomputed by refill_kernel on a side stream while the next step integrates.
template <typename T> struct Spare {
    Soa<T> S2;        // r [n][RF_COUNT], i [n][IF_COUNT], err_ring [n][3] of the precomputed rows: env-major, so that
                      // the copy of one env's row reads contiguous memory (one page, full sectors)
    float* obs;       // [n][obs_dim]
    double* obs64;    // [n][obs_dim]
    int32_t* list;    // [2][n] envs that consumed their row at a step (indexed by step parity)
    int32_t* count;   // [2]
    int on;           // 0: inline reset (waypoint env)
};

// reset_env addresses field f of env e as base[f * S.n + e]; with n = 1 and the base moved by e * (fields - 1) that is
// base[e * fields + f], the env-major spare layout
template <typename T>
__device__ __forceinline__ Soa<T> spare_view(const Spare<T>& P, int env) {
    Soa<T> L = P.S2;
    L.r += (size_t)env * (RF_COUNT - 1);
    L.i += (size_t)env * (IF_COUNT - 1);
    L.err_ring += (size_t)env * 2;
    if (L.err_ring_a) L.err_ring_a += (size_t)env * 2;
    if (L.par) L.par += (size_t)env * (FW_PAR_FIELDS - 1);
    L.n = 1;
    return L;
}

// Warp-cooperative: for every lane whose episode ended (`mine`), all converged lanes of the warp copy that env's
// precomputed row, field f by lane f mod width, every lane's loads in flight at once — one DRAM round trip per
// finished env instead of one per field.  IF_GOAL_ACHIEVED survives a reset (fixed_wing.py keeps goal_achieved across
// episodes); IF_NFEV / IF_NATT keep the diagnostics of the step that ended the episode.
template <typename T>
__device__ __forceinline__ void take_spare_warp(const Soa<T>& S, const Spare<T>& P, bool mine, int env, int odim,
                                                float* obs, double* obs64) {
    // fields in use: the attitude_angular block at the end of both field lists only when the config has such targets
    // (212 real fields are ONE pass of 8 x 32 loads per lane, 257 would be two)
    const int n_rf = S.err_ring_a ? (int)RF_COUNT : (int)RF_ATGT, n_if = S.err_ring_a ? (int)IF_COUNT : (int)IF_AGOAL_RING;
    const unsigned act = __activemask();
    unsigned dm = __ballot_sync(act, mine);
    if (!dm) return;
    __syncwarp(act);                       // the owners' earlier stores to their rows are ordered before the copies
    const int lane = threadIdx.x & 31;
    const int rank = __popc(act & ((1u << lane) - 1u)), width = __popc(act);
    const int n = S.n;
    while (dm) {
        const int owner = __ffs(dm) - 1;
        dm &= dm - 1;
        const int e = __shfl_sync(act, env, owner);
        {
            const T* src = P.S2.r + (size_t)e * RF_COUNT;
            T* dst = S.r + e;
#pragma unroll 1
            for (int f0 = rank; f0 < n_rf; f0 += 8 * width) {
                T tmp[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) tmp[j] = (f0 + j * width < n_rf) ? src[f0 + j * width] : (T)0;
#pragma unroll
                for (int j = 0; j < 8; ++j) if (f0 + j * width < n_rf) dst[(size_t)(f0 + j * width) * n] = tmp[j];
            }
        }
        {
            const int32_t* src = P.S2.i + (size_t)e * IF_COUNT;
            int32_t* dst = S.i + e;
#pragma unroll 1
            for (int f0 = rank; f0 < n_if; f0 += 2 * width) {
                int32_t tmp[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) tmp[j] = (f0 + j * width < n_if) ? src[f0 + j * width] : 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int f = f0 + j * width;
                    if (f < n_if && f != IF_GOAL_ACHIEVED && f != IF_NFEV && f != IF_NATT) dst[(size_t)f * n] = tmp[j];
                }
            }
        }
        if (rank < 3) S.err_ring[(size_t)rank * n + e] = P.S2.err_ring[(size_t)e * 3 + rank];
        if (rank < 3 && S.err_ring_a) S.err_ring_a[(size_t)rank * n + e] = P.S2.err_ring_a[(size_t)e * 3 + rank];
        if (S.par)          // the next episode's aircraft parameters (simulator.model)
            for (int f = rank; f < FW_PAR_FIELDS; f += width) S.par[(size_t)f * n + e] = P.S2.par[(size_t)e * FW_PAR_FIELDS + f];
        for (int q = rank; q < odim; q += width) {
            if (obs) obs[(size_t)e * odim + q] = P.obs[(size_t)e * odim + q];
            if (obs64) obs64[(size_t)e * odim + q] = P.obs64[(size_t)e * odim + q];
        }
    }
}

// the row of the episode after the one env is in now (main S): spare.episode := main.episode, then reset the spare
template <typename T>
__device__ void make_spare(const DCfg<T>& c, const Soa<T>& S, const Spare<T>& P, int env) {
    const Soa<T> L = spare_view(P, env);
    L.i[IF_EPISODE + env] = S.i[IF_EPISODE * S.n + env];
    reset_env<T>(c, L, env, nullptr, nullptr, P.obs, P.obs64);
}

template <typename T>
__global__ void reset_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S, const Spare<T> P, const uint8_t* mask,
                             const double* state_in, const double* target_in, float* obs, double* obs64) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.n) return;
    if (mask && !mask[env]) return;
    if (c.env_kind == FW_ENV_WAYPOINT) { wp_reset_env<T>(c, S, env, obs, obs64); return; }
    reset_env<T>(c, S, env, state_in, target_in, obs, obs64, true);
    if (P.on) make_spare<T>(c, S, P, env);
}

template <typename T>
__global__ void __launch_bounds__(32) refill_kernel(const __grid_constant__ DCfg<T> c, const Spare<T> P, int parity) {
    const int n = P.S2.n;
    const int cnt = P.count[parity];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        // the consumed row carried episode e + 1, which is what the spare's own counter still says: this makes e + 2
        const int env = P.list[(size_t)parity * n + j];
        reset_env<T>(c, spare_view(P, env), env, nullptr, nullptr, P.obs, P.obs64);
    }
}

// fw_set_config: every env's precomputed next-episode row is recomputed from the new reset-time configuration
template <typename T>
__global__ void __launch_bounds__(128) respare_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S, const Spare<T> P) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= S.n) return;
    make_spare<T>(c, S, P, env);
}

// sum |diff| of one column over the trailing window, accumulated in float32 like
// np.sum(np.abs(np.diff(...)), dtype=np.float32) (fixed_wing.py:1198-1228)
template <typename T>
__device__ __forceinline__ T delta_feature(const T cur, const T* ring /* [4][3] strided by 3 */, int col, int n_prev,
                                           bool f32) {
    // chronological order: oldest previous ... most recent previous, current
    float s = 0.f;
    T newer;
#pragma unroll
    for (int age = 4; age >= 1; --age) {
        if (age <= n_prev) {
            const T older = ring[(age - 1) * 3 + col];
            newer = (age == 1) ? cur : ring[(age - 2) * 3 + col];
            const float d = f32 ? fabsf((float)newer - (float)older) : (float)M<T>::fabs(newer - older);
            s += d;
        }
    }
    return (T)s;
}

// scratch SoA between the kernels of one step
template <typename T> struct Scratch {
    T* f0;        // [18][n]  fun(t0, y0)                           (init -> attempt)
    T* hinit;     // [n]      first step size                        (init -> attempt)
    T* cmd;       // [3][n]   constrained elevon_r / elevon_l / throttle commands
    T* turb;      // [6][n]   turbulence sample of this step (lin3, ang3)
    T* ytmp;      // [19][n]  sol.y[:, -1] before the commit          (attempt / rk4 -> head)
    int32_t* fail;     // [n] FwTermCode raised inside the integrator (0 = none)
    int32_t* counter;  // work-queue head of the attempt kernel
};

// raw action -> (a_raw, act_f32, constrained dynamics commands, constrained input commands)
// fixed_wing.py:491-506 + Actuation.set_and_constrain_commands pyfly.py:545-582
template <typename T>
__device__ __forceinline__ void prep_action(const DCfg<T>& c, const StepIO& io, int env, T (&a_raw)[3], bool& act_f32,
                                            T (&cmd_dyn)[3], T (&cmd_in)[3]) {
    if (io.random_actions) {
        const uint4 rr = rng_block(io.action_seed, c.env_id_offset + env, io.action_step >> 32, RNG_ACTION,
                                   (uint32_t)io.action_step);
        const uint32_t w[3] = {rr.x, rr.y, rr.z};
#pragma unroll
        for (int j = 0; j < 3; ++j) a_raw[j] = (T)(float)(((double)w[j] + 0.5) * (2.0 / 4294967296.0) - 1.0);
        act_f32 = true;
    } else if (io.actions_f64) {
#pragma unroll
        for (int j = 0; j < 3; ++j) a_raw[j] = (T)((const double*)io.actions)[(size_t)env * 3 + j];
        act_f32 = false;
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) a_raw[j] = (T)((const float*)io.actions)[(size_t)env * 3 + j];
        act_f32 = true;
    }
    T a_cmd[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (c.scale_actions) {   // linear_action_scaling(clip(action)) (fixed_wing.py:497-504, 630-652)
            const T cl = clip(a_raw[j], c.scale_low, c.scale_high);
            const T num = act_f32 ? (T)((float)cl - (float)c.scale_low) : (cl - c.scale_low);
            a_cmd[j] = (c.act_hi[j] - c.act_lo[j]) * num / (c.scale_high - c.scale_low) + c.act_lo[j];
        } else {
            a_cmd[j] = a_raw[j];
        }
    }
    cmd_dyn[0] = clip(-a_cmd[1] + a_cmd[0], c.elevon_min, c.elevon_max);
    cmd_dyn[1] = clip(a_cmd[1] + a_cmd[0], c.elevon_min, c.elevon_max);
    cmd_dyn[2] = clip(a_cmd[2], c.throttle_min, c.throttle_max);
    cmd_in[0] = clip((cmd_dyn[0] + cmd_dyn[1]) / (T)2, c.act_lo[0], c.act_hi[0]);
    cmd_in[1] = clip((-cmd_dyn[0] + cmd_dyn[1]) / (T)2, c.act_lo[1], c.act_hi[1]);
    cmd_in[2] = cmd_dyn[2];
}

// loads what one integration needs: y, steady wind, this step's turbulence sample, commands
template <typename T, bool TURB>
__device__ __forceinline__ void load_dyn(const DCfg<T>& c, const Soa<T>& S, const StepIO& io, int env, T (&y)[FW_NY],
                                         DynCtx<T>& x, T& elev0, T& ail0) {
    const int n = S.n;
    const T* r = S.r + env;
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) y[i] = r[(RF_Y + i) * n];
#pragma unroll
    for (int k = 0; k < 3; ++k) x.wind[k] = r[(RF_WIND + k) * n];
    if (TURB) {
        T fx[12], fu[4];
#pragma unroll
        for (int i = 0; i < 12; ++i) fx[i] = r[(RF_FX + i) * n];
#pragma unroll
        for (int i = 0; i < 4; ++i) fu[i] = r[(RF_FU + i) * n];
        turb_eval(c, fx, fu, x.tl, x.ta);
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) { x.tl[k] = 0; x.ta[k] = 0; }
    }
    T a_raw[3], cmd_in[3];
    bool act_f32;
    prep_action(c, io, env, a_raw, act_f32, x.cmd, cmd_in);
    // elevator/aileron seen by the t == 0 RHS call: 0 right after a reset (disabled ControlVariable.reset)
    elev0 = 0; ail0 = 0;
    if (S.i[(c.env_kind == FW_ENV_WAYPOINT ? IF_STEPS_TGT : IF_STEPS) * n + env] > 0) {
        const T er = clip(y[13], c.elevon_min, c.elevon_max), el = clip(y[14], c.elevon_min, c.elevon_max);
        elev0 = (er + el) / (T)2;
        ail0 = (-er + el) / (T)2;
    }
}

// ---- kernel A0: RungeKutta.__init__ + select_initial_step for every env, lock step ----
template <typename T, bool TURB, bool PE>
__global__ void __launch_bounds__(128, 4) rk45_init_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S, const StepIO io,
                                                        const Scratch<T> W, int attempt_threads, int32_t* done_count) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (env == 0) {
        *W.counter = attempt_threads;      // queue head: the first attempt_threads envs are pre-assigned
        *done_count = 0;                   // this step's list of envs that take their precomputed reset row
        if (io.info) *reinterpret_cast<int32_t*>(io.info) = 0;
    }
    if (env >= n) return;
    T y[FW_NY], h_abs = 0, elev0, ail0;
    DynCtx<T> x;
    load_dyn<T, TURB>(c, S, io, env, y, x, elev0, ail0);
    const int rc = rk45_init<T, TURB, PE>(c, x, y, S.r + (size_t)RF_Y * n + env, W.f0 + env, n, elev0, ail0, h_abs,
                                          PE ? S.par + env : nullptr);
    W.hinit[env] = h_abs;
#pragma unroll
    for (int k = 0; k < 3; ++k) { W.cmd[k * n + env] = x.cmd[k]; W.turb[k * n + env] = x.tl[k]; W.turb[(3 + k) * n + env] = x.ta[k]; }
    W.fail[env] = rc;
    S.i[IF_NFEV * n + env] = 2;              // a raise can only come from the second evaluation (t > 0)
    S.i[IF_NATT * n + env] = 0;
}

// ---- kernel A1: the attempt loop, persistent lanes pulling envs from a queue ----
template <typename T, bool TURB, int NT, bool PE>
__global__ void __maxnreg__(255) rk45_attempt_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S,
                                                          const Scratch<T> W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Stage storage holds FW_NS = 15 components: position (y[7..9]) never feeds back into the RHS, so its stage
    // inputs are never needed and its two quadratures (sum B_j K_j for y_new, sum E_j K_j for the error estimate) are
    // carried in registers, accumulated in the same j order as the stored components.
    // Layout: components in PAIRS, KV[(stage * FW_NP + pair) * NT + lane] = (component 2 pair, 2 pair + 1): one
    // 16-byte (f64) access per lane moves two values, consecutive lanes are consecutive -> conflict-free LDS.128 / STS.128.
    typedef typename Vec2<T>::type V2;
    V2* KV = reinterpret_cast<V2*>(smem_raw) + threadIdx.x;
    T* K = reinterpret_cast<T*>(KV);
#define KP(s, pp) KV[((s) * FW_NP + (pp)) * NT]
#define KS(s, cc) K[(((s) * FW_NP + ((cc) >> 1)) * NT) * 2 + ((cc) & 1)]
#define KC(cc) ((cc) < 7 ? (cc) : (cc) + 3)       /* stored slot -> ODE component */
    const int n = S.n;
    const T rtol = c.rtol, atol = c.atol, t_bound = c.dt;
    const unsigned lane = threadIdx.x & 31;
    int env = blockIdx.x * NT + threadIdx.x;
    bool need = true, exhausted = false, first_fetch = true;
    T y[FW_NY], ys[FW_NY], dyv[FW_NY];
    T kpos0[3] = {0, 0, 0}, posB[3] = {0, 0, 0}, posE[3] = {0, 0, 0};
    DynCtx<T> x;
    T t = 0, t_new = 0, h = 0, h_abs = 0, min_step = 0;
    bool rejected = false;
    int nfev = 0, natt = 0;
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) { y[i] = 0; ys[i] = 0; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { x.cmd[k] = 0; x.wind[k] = 0; x.tl[k] = 0; x.ta[k] = 0; }
    y[0] = 1; ys[0] = 1; ys[10] = 20; y[10] = 20;      // a benign state for lanes that never get an env

    while (true) {
        // ---------------- fetch: lanes without an env take the next one from the queue ----------------
        while (true) {
            const unsigned want = __ballot_sync(0xffffffffu, need && !exhausted);
            if (want == 0) break;
            if (need && !exhausted) {
                if (!first_fetch) {
                    int base = 0;
                    const unsigned peers = __activemask();
                    const int leader = __ffs(peers) - 1;
                    if ((int)lane == leader) base = atomicAdd(W.counter, __popc(peers));
                    base = __shfl_sync(peers, base, leader);
                    env = base + __popc(peers & ((1u << lane) - 1u));
                }
                first_fetch = false;
                if (env >= n) exhausted = true;
                else if (W.fail[env] != 0) { /* raised during initialisation: nothing to integrate */ }
                else {
#pragma unroll
                    for (int i = 0; i < FW_NY; ++i) y[i] = S.r[(RF_Y + i) * n + env];
#pragma unroll
                    for (int cc = 0; cc < FW_NS; ++cc) KS(0, cc) = W.f0[KC(cc) * n + env];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        kpos0[k] = W.f0[(7 + k) * n + env];
                        x.cmd[k] = W.cmd[k * n + env]; x.wind[k] = S.r[(RF_WIND + k) * n + env];
                        x.tl[k] = W.turb[k * n + env]; x.ta[k] = W.turb[(3 + k) * n + env];
                    }
                    h_abs = W.hinit[env];
                    t = 0; rejected = false; nfev = 2; natt = 0;
                    // top of RungeKutta._step_impl (rk.py:111-127)
                    min_step = ulp10<T>(t);
                    if (h_abs < min_step) h_abs = min_step;
                    need = false;
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted)) break;
        const bool active = !exhausted && !need;
        // ---------------- one attempt: top of the `while not step_accepted` loop (rk.py:129-141) ----------------
        bool finished = false;
        int rc = 0;
        if (active) {
            if (h_abs < min_step) finished = true;     // TOO_SMALL_STEP: solve_ivp status -1, ignored by pyfly
            else {
                t_new = t + h_abs;
                if (t_new - t_bound > (T)0) t_new = t_bound;
                h = t_new - t;
                h_abs = M<T>::fabs(h);
                natt++;
            }
        }
        const bool run = active && !finished;
#pragma unroll
        for (int k = 0; k < 3; ++k) { posB[k] = kpos0[k] * (T)RK_ROW[6][0]; posE[k] = kpos0[k] * (T)RK_E[0]; }
#pragma unroll 1
        for (int row = 1; row <= 6; ++row) {
            {
                // sum_{j < row} A[row][j] K_j, j ascending.  The j loop stays a loop: written out for every j behind
                // warp-uniform tests of `row` it needs 255 registers + 330 B of spills and the kernel is 25 % slower (measured).
                // Two terms per trip put 16 LDS.128 in flight ahead of 30 FMAs (same FMA order per component): 154.6 ->
                // 150.4 us in the A/B; 3 / 4 / 6 terms per trip 151.6 / 152.2 / 152.7, explicit ping-pong prefetch 159-162.
                T acc[2 * FW_NP];
#pragma unroll
                for (int cc = 0; cc < 2 * FW_NP; ++cc) acc[cc] = 0;
#define FW_TERM(j)                                                                      \
                {                                                                       \
                    const T a = (T)RK_ROW[row][j];                                      \
                    _Pragma("unroll") for (int pp = 0; pp < FW_NP; ++pp) {              \
                        const V2 kv = KP(j, pp);                                        \
                        acc[2 * pp] += kv.x * a;                                        \
                        if (2 * pp + 1 < FW_NS) acc[2 * pp + 1] += kv.y * a;            \
                    }                                                                   \
                }
#pragma unroll 2
                for (int j = 0; j < row; ++j) FW_TERM(j)
#undef FW_TERM
#pragma unroll
                for (int cc = 0; cc < FW_NS; ++cc) ys[KC(cc)] = y[KC(cc)] + acc[cc] * h;
#pragma unroll
                for (int k = 0; k < 3; ++k) ys[7 + k] = y[7 + k] + posB[k] * h;     // == y_new at row 6; unused before
                ys[18] = y[18];
            }
            const int r2 = rhs<T, TURB, PE>(c, x, ys, false, (T)0, (T)0, dyv, PE ? S.par + (env < n ? env : 0) : nullptr, n);
            if (run && rc == 0) {
                nfev++;
                if (r2) rc = r2;
                else if (row < 6) {
#pragma unroll
                    for (int pp = 0; pp < FW_NP; ++pp) {
                        V2 kv;
                        kv.x = dyv[KC(2 * pp)];
                        kv.y = (2 * pp + 1 < FW_NS) ? dyv[KC(2 * pp + 1)] : (T)0;
                        KP(row, pp) = kv;
                    }
                    const T bj = (T)RK_ROW[6][row], ej = (T)RK_E[row];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { posB[k] += dyv[7 + k] * bj; posE[k] += dyv[7 + k] * ej; }
                }
            }
        }
        if (run) {
            if (rc) finished = true;
            else {
                // ys == y_new, dyv == f_new: error estimate and step-size control (rk.py:100-104, 139-166)
                T errsq = 0;
#pragma unroll
                for (int i = 0; i < FW_NK; ++i) {
                    T acc = 0;
                    if (i >= 7 && i <= 9) acc = posE[i - 7];
                    else {
                        const int cc = i < 7 ? i : i - 3;
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc += KS(j, cc) * (T)RK_E[j];
                    }
                    acc += dyv[i] * (T)RK_E[6];
                    const T scl = atol + M<T>::fmax(M<T>::fabs(y[i]), M<T>::fabs(ys[i])) * rtol;
                    const T ei = acc * h * M<T>::rcp_hot(scl);      // scl >= atol: normal range
                    errsq += ei * ei;
                }
                const T err = M<T>::sqrt(errsq) / M<T>::sqrt((T)FW_NY);
                // err == 0: pow(0, -0.2) = inf -> min(10, inf) = MAX_FACTOR, the same as scipy's special case
                const T pf = (T)0.9 * M<T>::template pow_hot<true>(err, (T)-0.2);
                if (err < (T)1) {
                    T factor = (err == (T)0) ? (T)10 : M<T>::fmin((T)10, pf);
                    if (rejected) factor = M<T>::fmin((T)1, factor);
                    h_abs *= factor;
                    t = t_new;
#pragma unroll
                    for (int i = 0; i < FW_NK; ++i) y[i] = ys[i];
#pragma unroll
                    for (int cc = 0; cc < FW_NS; ++cc) KS(0, cc) = dyv[KC(cc)];              // FSAL
#pragma unroll
                    for (int k = 0; k < 3; ++k) kpos0[k] = dyv[7 + k];
                    if (!(t < t_bound)) finished = true;            // solver.status == 'finished'
                    else {
                        rejected = false;
                        min_step = ulp10<T>(t);
                        if (h_abs < min_step) h_abs = min_step;
                    }
                } else {
                    h_abs *= M<T>::fmax((T)0.2, pf);
                    rejected = true;
                }
            }
        }
        if (active && finished) {
#pragma unroll
            for (int i = 0; i < FW_NY; ++i) W.ytmp[i * n + env] = y[i];
            W.fail[env] = rc;
            S.i[IF_NFEV * n + env] = nfev;
            S.i[IF_NATT * n + env] = natt;
            need = true;
        }
    }
#undef KC
#undef KS
#undef KP
}

// ---- fixed-step modes: lock-step integrate kernel (no adaptivity, no divergence to rebalance) ----
template <typename T, bool TURB, bool PE>
__global__ void __launch_bounds__(128) rk4_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S, const StepIO io,
                                                  const Scratch<T> W, int32_t* done_count) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (env == 0) {
        *done_count = 0;
        if (io.info) *reinterpret_cast<int32_t*>(io.info) = 0;
    }
    if (env >= n) return;
    T y[FW_NY], elev0, ail0;
    DynCtx<T> x;
    load_dyn<T, TURB>(c, S, io, env, y, x, elev0, ail0);
    int nfev = 0, natt = 0;
    const int rc = solve_rk4<T, TURB, PE>(c, x, y, elev0, ail0, nfev, natt, PE ? S.par + env : nullptr, n);
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) W.ytmp[i * n + env] = y[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) { W.turb[k * n + env] = x.tl[k]; W.turb[(3 + k) * n + env] = x.ta[k]; }
    W.fail[env] = rc;
    S.i[IF_NFEV * n + env] = nfev;
    S.i[IF_NATT * n + env] = natt;
}

// PyFly.step after solve_ivp (pyfly.py:1396-1406, 1852-1881): quaternion renormalisation, Euler angles, constraint checks
// on p, q, r, actuator clips, Va / alpha / beta.  roll .. om_obs enter holding the previous committed values and leave
// holding what each variable's `.history[-1]` is after a (possibly partial) commit.  Returns the FwTermCode.
template <typename T>
__device__ __forceinline__ int post_step_commit(const DCfg<T>& c, const DynCtx<T>& x, T (&y)[FW_NY], int fail, T& roll,
                                                T& pitch, T& Va, T& alpha, T& beta, T (&om_obs)[3]) {
    if (!fail) {
        const T nrm = M<T>::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = y[i] / nrm;
        const T e0 = y[0], e1 = y[1], e2 = y[2], e3 = y[3];
        roll = M<T>::atan2_hot((T)2 * (e0 * e1 + e2 * e3), e0 * e0 + e3 * e3 - e1 * e1 - e2 * e2);
        pitch = M<T>::asin_hot((T)2 * (e0 * e2 - e1 * e3));
        // yaw (pyfly.py:704-706) only feeds the Euler-angle rotation below and is not observed: it is not materialised
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (!fail) {
                if (y[4 + i] < c.omega_con_min[i] || y[4 + i] > c.omega_con_max[i]) fail = FW_TERM_OMEGA_P + i;
                else om_obs[i] = y[4 + i];
            }
        }
        if (!fail) {
            y[13] = clip(y[13], c.elevon_min, c.elevon_max);
            y[14] = clip(y[14], c.elevon_min, c.elevon_max);
            y[15] = clip(y[15], c.throttle_min, c.throttle_max);
            y[16] = clip(y[16], -c.elevon_dot_max, c.elevon_dot_max);
            y[17] = clip(y[17], -c.elevon_dot_max, c.elevon_dot_max);
            // pyfly.py:1398-1406 rotates the steady wind with the matrix built from (roll, pitch, yaw); for the
            // normalised quaternion those Euler angles came from, that matrix equals the quaternion form of
            // _rot_b_v (pyfly.py:1782-1800) up to rounding, so no angle -> sin/cos round trip is needed.
            T wb[3];
            wb[0] = ((T)-1 + (T)2 * (e0 * e0 + e1 * e1)) * x.wind[0] + (T)2 * (e1 * e2 + e3 * e0) * x.wind[1] + (T)2 * (e1 * e3 - e2 * e0) * x.wind[2];
            wb[1] = (T)2 * (e1 * e2 - e3 * e0) * x.wind[0] + ((T)-1 + (T)2 * (e0 * e0 + e2 * e2)) * x.wind[1] + (T)2 * (e2 * e3 + e1 * e0) * x.wind[2];
            wb[2] = (T)2 * (e1 * e3 + e2 * e0) * x.wind[0] + (T)2 * (e2 * e3 - e1 * e0) * x.wind[1] + ((T)-1 + (T)2 * (e0 * e0 + e3 * e3)) * x.wind[2];
            const T a0 = y[10] - (wb[0] + x.tl[0]), a1 = y[11] - (wb[1] + x.tl[1]), a2 = y[12] - (wb[2] + x.tl[2]);
            T Van = M<T>::sqrt(a0 * a0 + a1 * a1 + a2 * a2);
            const T al = M<T>::atan2_hot(a2, a0), be = M<T>::asin_hot(a1 / Van);
            if (c.va_con_max > (T)0 && Van > c.va_con_max) fail = FW_TERM_VA;
            else {
                if (Van < c.va_value_min) Van = c.va_value_min;
                Va = Van; alpha = al; beta = be;
            }
        }
    }
    return fail;
}

// sum of the last `cnt` entries of history["error"] per target state (end_error, fixed_wing.py:1690-1700), added in
// chronological order.  The ring rows are read ten entries x three states at a time: one load per addition was one
// exposed DRAM round trip per addition, 150 in a row, and — one wave of warps running the kernel — stretched every
// step in which any episode ended by 70 us.
template <typename T>
__device__ __noinline__ void end_error_sums(const Soa<T>& S, int env, int n_err, int cnt, bool have_new,
                                            const T (&e_new)[3], T (&s50)[3]) {
    const int n = S.n;
    const int newest = (n_err - 1) % FW_END_ERR_WINDOW;
    s50[0] = s50[1] = s50[2] = 0;
#pragma unroll 1
    for (int q0 = 0; q0 < cnt; q0 += 10) {
        T v[3][10];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const int q = q0 + j;
            const int slot = (n_err - cnt + (q < cnt ? q : 0)) % FW_END_ERR_WINDOW;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                v[k][j] = (q >= cnt) ? (T)0 : (slot == newest && have_new) ? e_new[k]
                                                                           : S.err_ring[(size_t)(slot * 3 + k) * n + env];
        }
#pragma unroll
        for (int j = 0; j < 10; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) s50[k] += v[k][j];
    }
}

// ---- kernel B: everything after the integrator (once per step, high occupancy) ----
// One wave of warps runs this kernel (1024 blocks of 64 threads, 8 per SM), so its duration is the critical path of ONE
// warp: every exposed DRAM round trip and every long dependent chain adds to it in full.  Hence: the integrator
// output is loaded without waiting for the `fail` flag; the rare episode-end work is warp-cooperative (end-error ring
// sums, the copy of the precomputed next-episode row) instead of one lane walking hundreds of dependent loads while the
// whole grid waits (that was +70 us on every step in which any episode ended); the episode-end rows for the host are
// written here so that they travel with the step outputs.  Two restructurings were measured and dropped: strict
// load -> compute -> store phases per subsystem (one round trip per phase: 1.5x slower) and hoisting every load to the
// top (1.5 KB of spills, 10 % slower).

// The same sums with the whole warp fetching: every lane loads 5 of the (up to) 150 ring entries of a finished env —
// one DRAM round trip instead of five — and the owner adds them up in chronological order through shuffles.  The
// entry of the current step is already in the ring (the owner stored it; __syncwarp orders that store before the loads).
template <typename T>
__device__ __forceinline__ void end_error_sums_warp(const Soa<T>& S, bool mine, int env, int n_err, T (&s50)[3]) {
    const unsigned act = __activemask();
    s50[0] = s50[1] = s50[2] = 0;
    if (act != 0xffffffffu) {              // ragged last block: every finished env sums for itself
        if (mine) {
            const T none[3] = {0, 0, 0};
            end_error_sums<T>(S, env, n_err, n_err < FW_END_ERR_WINDOW ? n_err : FW_END_ERR_WINDOW, false, none, s50);
        }
        return;
    }
    unsigned dm = __ballot_sync(act, mine);
    if (!dm) return;
    __syncwarp(act);
    const int lane = threadIdx.x & 31, n = S.n;
    while (dm) {
        const int owner = __ffs(dm) - 1;
        dm &= dm - 1;
        const int e = __shfl_sync(act, env, owner), ne = __shfl_sync(act, n_err, owner);
        const int cnt = ne < FW_END_ERR_WINDOW ? ne : FW_END_ERR_WINDOW;
        T v[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int t = j * 32 + lane, q = t / 3, k = t - 3 * q;          // entry q (chronological), state k
            const int slot = (ne - cnt + (q < cnt ? q : 0)) % FW_END_ERR_WINDOW;
            v[j] = (q < cnt) ? S.err_ring[(size_t)(slot * 3 + k) * n + e] : (T)0;
        }
        T a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
#pragma unroll 1
            for (int l = 0; l < 32; l += 3) {                              // t = j * 32 + l; state k = t % 3 rotates
                const int k0 = (j * 32 + l) % 3;
                const T x0 = __shfl_sync(act, v[j], l), x1 = __shfl_sync(act, v[j], (l + 1) & 31),
                        x2 = __shfl_sync(act, v[j], (l + 2) & 31);
                const bool h1 = l + 1 < 32, h2 = l + 2 < 32;
                // entries beyond 3 * cnt were loaded as 0; adding 0 leaves the sums as they are
                if (k0 == 0) { a0 += x0; if (h1) a1 += x1; if (h2) a2 += x2; }
                else if (k0 == 1) { a1 += x0; if (h1) a2 += x1; if (h2) a0 += x2; }
                else { a2 += x0; if (h1) a0 += x1; if (h2) a1 += x2; }
            }
        }
        if (lane == owner) { s50[0] = a0; s50[1] = a1; s50[2] = a2; }
    }
}

template <typename T, bool TURB, bool GENERIC>
__global__ void __launch_bounds__(64, 8) head_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S, const StepIO io,
                                                  const Scratch<T> W, const Spare<T> P, int parity) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (env >= n) return;
    T* r = S.r + env;
    int32_t* ii = S.i + env;
    // The goal-ring words are read late, behind stores the compiler may not hoist them above: their lines are fetched
    // into L1 now, without holding registers for them.  (The episode-statistics fields were prefetched here too until
    // their loads moved ahead of the observation section, below.)
#pragma unroll
    for (int f = IF_GOAL_RING; f < IF_GOAL_RING + 28; ++f) prefetch_l1(ii + (size_t)f * n);
    int fail = W.fail[env];
    T y[FW_NY];
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) y[i] = W.ytmp[i * n + env];     // not waiting for `fail`: a raise is the rare case
    if (fail) {
#pragma unroll
        for (int i = 0; i < FW_NY; ++i) y[i] = r[(RF_Y + i) * n];
    }
    const T roll_prev = r[RF_ROLL * n], pitch_prev = r[RF_PITCH * n], Va_prev = r[RF_VA * n];
    const T alpha_prev = r[RF_ALPHA * n], beta_prev = r[RF_BETA * n];
    const T omega_prev[3] = {r[(RF_Y + 4) * n], r[(RF_Y + 5) * n], r[(RF_Y + 6) * n]};
    DynCtx<T> x;
    T tgt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x.wind[k] = r[(RF_WIND + k) * n]; tgt[k] = r[(RF_TGT + k) * n];
        x.tl[k] = TURB ? W.turb[k * n + env] : (T)0; x.ta[k] = TURB ? W.turb[(3 + k) * n + env] : (T)0;
    }
    int steps = ii[IF_STEPS * n], steps_tgt = ii[IF_STEPS_TGT * n], sim_step = ii[IF_SIM_STEP * n];
    const unsigned long long episode = (unsigned long long)(uint32_t)ii[IF_EPISODE * n];
    const int nfev = ii[IF_NFEV * n], natt = ii[IF_NATT * n];
    T a_raw[3], cmd_in[3];
    bool act_f32;
    prep_action(c, io, env, a_raw, act_f32, x.cmd, cmd_in);
    T fx[12], fu[4];
    if (TURB) {
#pragma unroll
        for (int i = 0; i < 12; ++i) fx[i] = r[(RF_FX + i) * n];
#pragma unroll
        for (int i = 0; i < 4; ++i) fu[i] = r[(RF_FU + i) * n];
    }

    // ---------------- post-step commit (pyfly.py:1396-1406, 1852-1881) ----------------
    T roll = roll_prev, pitch = pitch_prev, Va = Va_prev, alpha = alpha_prev, beta = beta_prev;
    T om_obs[3] = {omega_prev[0], omega_prev[1], omega_prev[2]};   // .history[-1] view for a terminal observation
    fail = post_step_commit<T>(c, x, y, fail, roll, pitch, Va, alpha, beta, om_obs);
    sim_step += 1;

    // ---------------- gym head (fixed_wing.py:512-628) ----------------
    const int steps_before = steps;
    steps += 1;
    steps_tgt += 1;
    bool done = false;
    int term = FW_TERM_NONE;
    if (c.steps_max > 0 && steps >= c.steps_max) { done = true; term = FW_TERM_STEPS; }

    // action / command rings: previous entries (age 1 = most recent)
    T aring[12], cring[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) aring[i] = r[(RF_ACT_RING + i) * n];
    const bool need_cring = !c.scale_actions;
#pragma unroll
    for (int i = 0; i < 12; ++i) cring[i] = (need_cring || i < 3) ? r[(RF_CMD_RING + i) * n] : (T)0;
    T cv_sum = r[RF_CV_SUM * n];
    if (steps_before >= 1) {
#pragma unroll
        for (int j = 0; j < 3; ++j) cv_sum += M<T>::fabs(cmd_in[j] - cring[j]);
    }
    const int n_prev = steps_before < 4 ? steps_before : 4;   // valid previous ring entries

    T e_new[3] = {0, 0, 0};
    T reward;
    T obs_v[FW_NOBS];
    int gbits[7] = {0, 0, 0, 0, 1, 1, 1};       // roll pitch Va all | omega_p omega_q omega_r (attitude_angular)
    T atgt[3] = {0, 0, 0}, ea_new[3] = {0, 0, 0};
    const bool ang = GENERIC && c.ang_on;
    if (ang) {
#pragma unroll
        for (int a = 0; a < 3; ++a) atgt[a] = r[(RF_ATGT + a) * n];
    }
    int gcnt[4], gtot[4], settle[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gcnt[k] = ii[(IF_GOAL_CNT + k) * n]; gtot[k] = ii[(IF_GOAL_TOTAL + k) * n]; settle[k] = ii[(IF_SETTLE + k) * n];
    }
    if (!fail) {
        // goal status with the CURRENT target (fixed_wing.py:536-560)
        const T eg[6] = {err_roll(tgt[0], roll), tgt[1] - pitch, tgt[2] - Va, atgt[0] - y[4], atgt[1] - y[5], atgt[2] - y[6]};
        bool resample = false, success_on_step = false, resampled = false;
        if (c.streak_req > 0) {
            gbits[3] = 1;
#pragma unroll
            for (int k = 0; k < 3; ++k) { gbits[k] = M<T>::fabs(eg[k]) <= c.tgt_bound[k]; gbits[3] &= gbits[k]; }
            if (ang) {              // bounded rate targets take part in "all" (fixed_wing.py:1346-1361)
#pragma unroll
                for (int a = 0; a < 3; ++a) { gbits[4 + a] = M<T>::fabs(eg[3 + a]) <= c.ang_bound[a]; gbits[3] &= gbits[4 + a]; }
                const int ga[3] = {gbits[4], gbits[5], gbits[6]};
                angular_goal_update<T>(c, S, env, steps, ga);
            }
            const int idx = steps;                 // index of this entry in history["goal"] (entry 0 = reset)
            const int w = (idx & 127) >> 5, b = idx & 31;
            const int idx_old = idx - c.streak_req;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int32_t* ring = ii + (IF_GOAL_RING + 4 * k) * n;
                if (idx_old >= 0) gcnt[k] -= (ring[((idx_old & 127) >> 5) * n] >> (idx_old & 31)) & 1;
                uint32_t word = (uint32_t)ring[w * n];
                word = (word & ~(1u << b)) | ((uint32_t)gbits[k] << b);
                ring[w * n] = (int32_t)word;
                gcnt[k] += gbits[k];
                gtot[k] += gbits[k];
                if (settle[k] < 0 && idx + 1 >= c.streak_req &&
                    (double)gcnt[k] / (double)c.streak_req >= (double)c.streak_fraction) settle[k] = idx;
            }
            if (steps_tgt >= c.streak_req && (double)gcnt[3] / (double)c.streak_req >= (double)c.streak_fraction) {
                if (GENERIC && c.rew_generic) {            // goal_achieved_on_step (fixed_wing.py:546-547)
                    success_on_step = ii[IF_GOAL_ACHIEVED * n] == 0;
                    ii[IF_GOAL_ACHIEVED * n] = 1;
                }
                if (c.on_success == FW_SUCCESS_DONE) { done = true; term = FW_TERM_SUCCESS; }
                else if (c.on_success == FW_SUCCESS_NEW) resample = true;
            }
        }
        // reward (fixed_wing.py:941-1111, default factor family)
        T val = 0;
        if (GENERIC && c.rew_generic) {
            const T st8[8] = {roll, pitch, Va, y[4], y[5], y[6], alpha, beta};
            val = generic_reward<T>(c, S, env, eg, st8, a_raw, act_f32, aring, n_prev, steps, gbits, success_on_step);
        } else {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (c.rew_err_scaling[k] > (T)0) val -= clip(M<T>::fabs(eg[k]) / c.rew_err_scaling[k], (T)0, c.rew_err_max[k]);
        if (c.rew_delta_scaling > (T)0 && steps > 1) {
            const int np_ = (c.rew_delta_window - 1) < n_prev ? (c.rew_delta_window - 1) : n_prev;
            T dv;
            if (act_f32) {
                const float sd = ring_delta_sum_np<float, T>(np_, a_raw, aring);
                dv = (T)fminf(fmaxf(sd / (float)c.rew_delta_scaling, 0.f), (float)c.rew_delta_max);
            } else {
                dv = clip(ring_delta_sum_np<T, T>(np_, a_raw, aring) / c.rew_delta_scaling, (T)0, c.rew_delta_max);
            }
            val -= dv;
        }
        if (c.rew_bound_scaling > (T)0 && c.has_action_bounds) {
            T hi = 0, lo = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (a_raw[j] > c.action_bounds_max[j]) hi += M<T>::fabs(a_raw[j] - c.action_bounds_max[j]);
                if (a_raw[j] < c.action_bounds_min[j]) lo += M<T>::fabs(a_raw[j] - c.action_bounds_min[j]);
            }
            val -= clip(M<T>::fabs(hi + lo) / c.rew_bound_scaling, (T)0, c.rew_bound_max);
        }
        }
        reward = val;
        // target resample / advance (fixed_wing.py:569-580, 1363-1471)
        int tcls[3] = {c.tgt_class[0], c.tgt_class[1], c.tgt_class[2]};
        T tp[15];
        if (GENERIC && c.tgt_moving) {
#pragma unroll
            for (int k = 0; k < 15; ++k) tp[k] = r[(RF_TPROP + k) * n];
#pragma unroll
            for (int k = 0; k < 3; ++k) tcls[k] = ii[(IF_TCLS + k) * n];
        }
        if (GENERIC && (resample || (c.resample_every > 0 && steps_tgt >= c.resample_every))) {
            T u12[12];
            target_draws<T>(c, env_seed(S, env), c.env_id_offset + env, episode, RNG_RESAMPLE, (uint32_t)(steps * 8), u12);
            sample_target<T>(c, *S.rc, roll, pitch, Va, steps, u12, tgt, tcls, tp);
            steps_tgt = 0;
            resampled = true;
            if (c.tgt_moving) {
#pragma unroll
                for (int k = 0; k < 15; ++k) r[(RF_TPROP + k) * n] = tp[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) ii[(IF_TCLS + k) * n] = tcls[k];
            }
        }
        if (ang) {
            // rate targets: re-derived from zero by a resample (sample_target), then advanced — every value from the
            // targets BEFORE this step's advance (fixed_wing.py:574-580, 1455-1460)
            const T er = err_roll(tgt[0], roll), ep = tgt[1] - pitch;
            if (resampled) angular_targets<T>(c, roll, pitch, er, ep, true, atgt);
            angular_targets<T>(c, roll, pitch, er, ep, false, atgt);
        }
        const T tgt_pitch_cur = tgt[1];
        if (GENERIC && c.tgt_moving) {
            const T TWO_PI = (T)6.283185307179586476925286766559;
#pragma unroll
            for (int k = 0; k < 2; ++k) {          // roll, pitch (Va is constant or compensate)
                if (tcls[k] == FW_TGT_LINEAR) tgt[k] = tgt[k] + tp[k] * c.dt;
                else if (tcls[k] == FW_TGT_SINUSOIDAL)
                    tgt[k] = tp[3 + k] * M<T>::sin(TWO_PI / tp[6 + k] * ((T)steps + tp[9 + k])) + tp[12 + k];
            }
            if (tcls[2] == FW_TGT_LINEAR) tgt[2] = tgt[2] + tp[2] * c.dt;
            else if (tcls[2] == FW_TGT_SINUSOIDAL)
                tgt[2] = tp[5] * M<T>::sin(TWO_PI / tp[8] * ((T)steps + tp[11])) + tp[14];
        }
        if (tcls[2] == FW_TGT_COMPENSATE) {
            // the Va law sees the pitch target itself, or the bias of a sinusoidal one (fixed_wing.py:1381-1384); every
            // new value is computed from the targets BEFORE this step's advance
            const T pitch_tar = (GENERIC && c.tgt_moving && tcls[1] == FW_TGT_SINUSOIDAL) ? tp[13] : tgt_pitch_cur, va_t = tgt[2];
            const T D2R = (T)(3.141592653589793238462643383279502884 / 180.0);
            if (pitch_tar <= (T)-2.5 * D2R) {
                const T va_end = (T)28.434 - (T)40.0841 * pitch_tar;
                T slope = 0;
                if (va_t <= va_end) {
                    const T s = (va_t < va_end * (T)0.95) ? (T)1 : (T)1 - va_t / (va_end * (T)1.5);
                    slope = (T)7 * M<T>::fmax((T)0, s);
                }
                tgt[2] = va_t + (slope * (-tgt_pitch_cur) - (T)0.25) * c.dt;
            } else if (pitch_tar >= (T)5 * D2R) {
                const T va_end = (T)26.27 - (T)41.2529 * pitch_tar;
                if (va_t > va_end) tgt[2] = (steps_tgt < 750) ? va_t + (va_end - va_t) * (T)1 / (T)150 : va_end;
            }
        }
        {
            const T PI = (T)3.141592653589793238462643383279502884;
            if (M<T>::fabs(tgt[0]) > PI) tgt[0] = sgn(tgt[0]) * (py_mod(M<T>::fabs(tgt[0]), PI) - PI);
        }
        e_new[0] = err_roll(tgt[0], roll); e_new[1] = tgt[1] - pitch; e_new[2] = tgt[2] - Va;
        if (ang) { ea_new[0] = atgt[0] - y[4]; ea_new[1] = atgt[1] - y[5]; ea_new[2] = atgt[2] - y[6]; }
    } else {
        done = true;
        term = fail;
        reward = c.step_fail_timesteps ? (T)(steps - c.steps_max) : c.step_fail_value;
    }

    // The episode-statistics fields are loaded HERE, ahead of the observation section: behind its stores their round trip
    // was exposed in full (11 % of the kernel's samples on the first use, profiles/r02_v9_stall_lines.txt)
    T esum[3], eabs[3], emin[3], emax[3], e0v[3], eprev[3];
    int rise_lo[3], rise_hi[3];
    const T ep_ret_prev = r[RF_EP_RET * n];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        e0v[k] = r[(RF_E0 + k) * n]; esum[k] = r[(RF_ESUM + k) * n]; eabs[k] = r[(RF_EABS + k) * n];
        emin[k] = r[(RF_EMIN + k) * n]; emax[k] = r[(RF_EMAX + k) * n]; eprev[k] = r[(RF_EPREV + k) * n];
        rise_lo[k] = ii[(IF_RISE_LO + k) * n]; rise_hi[k] = ii[(IF_RISE_HI + k) * n];
    }
    // ---------------- observation (fixed_wing.py:1113-1262) ----------------
    obs_v[0] = roll; obs_v[1] = pitch; obs_v[2] = Va;
    obs_v[3] = om_obs[0]; obs_v[4] = om_obs[1]; obs_v[5] = om_obs[2];
    obs_v[6] = tgt[0]; obs_v[7] = tgt[1]; obs_v[8] = tgt[2];
    obs_v[9] = alpha; obs_v[10] = beta;
    {
        const int np_ = (c.obs_act_window - 1) < n_prev ? (c.obs_act_window - 1) : n_prev;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            obs_v[11 + j] = c.scale_actions ? delta_feature<T>(a_raw[j], aring, j, np_, act_f32)
                                            : delta_feature<T>(cmd_in[j], cring, j, np_, false);
    }

    if (!(GENERIC && c.obs_generic) && (c.obs_noise_std > (T)0 || c.obs_noise_mean != (T)0))
        add_obs_noise<T>(c, env_seed(S, env), c.env_id_offset + env, episode, steps, obs_v, FW_NOBS);
    T og[GENERIC ? FW_NOBS_MAX : 1];
    const T* obs_out = obs_v;
    int odim = FW_NOBS;
    if (GENERIC && c.obs_generic) {
        const T e_cur[3] = {fail ? err_roll(tgt[0], roll) : e_new[0], fail ? tgt[1] - pitch : e_new[1], fail ? tgt[2] - Va : e_new[2]};
        const T cur[14] = {roll, pitch, Va, om_obs[0], om_obs[1], om_obs[2], alpha, beta, tgt[0], tgt[1], tgt[2],
                           e_cur[0], e_cur[1], e_cur[2]};
        const T actval[3] = {(y[13] + y[14]) / (T)2, (-y[13] + y[14]) / (T)2, y[15]};
        const T acur[6] = {atgt[0], atgt[1], atgt[2], fail ? atgt[0] - om_obs[0] : ea_new[0], fail ? atgt[1] - om_obs[1] : ea_new[1],
                           fail ? atgt[2] - om_obs[2] : ea_new[2]};
        generic_observation<T>(c, S, env, steps, !fail, cur, a_raw, act_f32, cmd_in, actval, episode, og, (const T*)nullptr,
                               ang ? acur : (const T*)nullptr);
        obs_out = og;
        odim = c.obs_len * c.obs_n;
    }

    // ---------------- streamed episode statistics (fixed_wing.py:1644-1736) ----------------
    const T ep_ret = ep_ret_prev + reward;
    int n_err = steps_before + 1;           // entries in history["error"] before this step
    if (!fail) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T ea = M<T>::fabs(e_new[k]), prev = eprev[k];
            const T low_lim = M<T>::fabs(c.rise_low * e0v[k]), high_lim = M<T>::fabs(c.rise_high * e0v[k]);
            if (rise_lo[k] < 0 && prev >= low_lim && ea < low_lim) rise_lo[k] = n_err - 1;
            if (rise_hi[k] < 0 && prev >= high_lim && ea < high_lim) rise_hi[k] = n_err - 1;
            esum[k] += e_new[k]; eabs[k] += ea;
            emin[k] = M<T>::fmin(emin[k], e_new[k]); emax[k] = M<T>::fmax(emax[k], e_new[k]);
            r[(RF_EPREV + k) * n] = ea;
            S.err_ring[(size_t)((n_err % FW_END_ERR_WINDOW) * 3 + k) * n + env] = e_new[k];
        }
        if (ang) angular_stats_update<T>(c, S, env, n_err, ea_new);
        n_err += 1;
    }

    T s50[3];
    end_error_sums_warp<T>(S, done, env, n_err, s50);
    if (done) {
        double* m = S.metrics + (size_t)env * FW_NMETRIC;
        const int off = steps - (n_err - 1);     // rise index offset: 0 normally, 1 after a failed step
        const int end_cnt = n_err < FW_END_ERR_WINDOW ? n_err : FW_END_ERR_WINDOW;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T e0 = e0v[k];
            m[FW_M_AVG_ERROR + k] = (M<T>::fabs(e0) >= (T)0.01) ? (double)M<T>::fabs((esum[k] / (T)n_err) / e0) : CUDART_NAN;
            m[FW_M_TOTAL_ERROR + k] = (double)eabs[k];
            m[FW_M_END_ERROR + k] = (double)M<T>::fabs(s50[k] / (T)end_cnt);
            const double re = rise_lo[k] >= 0 ? (double)(rise_lo[k] + off) : CUDART_NAN;
            const double rs = rise_hi[k] >= 0 ? (double)(rise_hi[k] + off) : CUDART_NAN;
            m[FW_M_RISE_TIME + k] = re - rs;
            const T opp = (e0 > (T)0) ? emin[k] : emax[k];
            m[FW_M_OVERSHOOT + k] = (sgn(opp) == sgn(e0)) ? CUDART_NAN : (double)M<T>::fabs(opp / e0);
        }
        m[FW_M_CONTROL_VARIATION] = (double)cv_sum / (3.0 * (double)c.dt * (double)(steps - 1));
        const int n_goal = fail ? steps : steps + 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m[FW_M_SUCCESS + k] = settle[k] >= 0 ? 1.0 : 0.0;
            m[FW_M_SETTLING_TIME + k] = settle[k] >= 0 ? (double)settle[k] : CUDART_NAN;
            m[FW_M_SUCCESS_TIME_FRAC + k] = (double)gtot[k] / (double)n_goal;
        }
        S.ep_ret[env] = (double)ep_ret;
        S.ep_len[env] = steps;
        if (ang) angular_metrics<T>(c, S, env, n_err, n_goal, off);
    }
    S.ep_term[env] = term;
    if (io.rew) io.rew[env] = (float)reward;
    if (io.rew64) io.rew64[env] = (double)reward;
    if (io.done) io.done[env] = done ? 1 : 0;

    // an episode that ended takes the next episode's row, computed ahead of time (refill_kernel recomputes it on the
    // side stream while the next step integrates)
    const bool take = done && io.auto_reset;
    if (take && io.term_obs) write_obs(obs_out, odim, env, io.term_obs, (double*)nullptr);
    // "integrator" observation entries of the RESET observation read the error history of the episode that just ended
    // (fixed_wing.py:453-460, 1165-1180): the precomputed row holds them as 0 and they are added here from the live ring
    T int_reset[6] = {0, 0, 0, 0, 0, 0};
    if constexpr (GENERIC) {
        if (take && c.obs_has_int) {
#pragma unroll 1
            for (int k = 0; k < (c.ang_on ? 6 : 3); ++k)
                int_reset[k] = integrator_reset_value<T>(c, S, env, k, n_err, k < 3 ? e0v[k < 3 ? k : 0] : r[(RF_AE0 + k - 3) * n]);
        }
    }
    take_spare_warp<T>(S, P, take, env, odim, io.obs, io.obs64);
    if constexpr (GENERIC) {
        if (c.obs_has_int) {
            __syncwarp(__activemask());          // the cooperative copy of the reset observation is done
            if (take) patch_integrator_reset_obs<T>(c, P, env, odim, int_reset, io.obs, io.obs64);
        }
    }
    if (take) {
        P.list[(size_t)parity * n + atomicAdd(P.count + parity, 1)] = env;
        if (io.info) {
            // packed episode-end row for the host (fw_set_info_rows): env, termination, length, return, 28 metrics,
            // terminal observation — what VecEnv infos need, fetched together with the step outputs
            const int slot = atomicAdd(reinterpret_cast<int32_t*>(io.info), 1);
            if (slot < io.info_cap) {
                double* row = io.info + 1 + (size_t)slot * (FW_INFO_HEAD + odim);
                const double* m = S.metrics + (size_t)env * FW_NMETRIC;
                row[0] = (double)env; row[1] = (double)term; row[2] = (double)steps; row[3] = (double)ep_ret;
                for (int k = 0; k < FW_NMETRIC; ++k) row[4 + k] = m[k];
                for (int q = 0; q < odim; ++q) row[FW_INFO_HEAD + q] = (double)(float)obs_out[q];
            }
        }
        return;
    }
    write_obs(obs_out, odim, env, io.obs, io.obs64);

    // ---------------- store ----------------
    if (TURB && !fail) {
        T un[4];
        noise_sample(c, S, env, episode, sim_step, un);
        turb_step(c, fx, fu, un, sim_step);
#pragma unroll
        for (int i = 0; i < 12; ++i) r[(RF_FX + i) * n] = fx[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[(RF_FU + i) * n] = fu[i];
    }
    if (!fail) {
#pragma unroll
        for (int i = 0; i < FW_NY; ++i) r[(RF_Y + i) * n] = y[i];
        r[RF_ROLL * n] = roll; r[RF_PITCH * n] = pitch; r[RF_VA * n] = Va; r[RF_ALPHA * n] = alpha; r[RF_BETA * n] = beta;
    }
    if (ang) {
#pragma unroll
        for (int a = 0; a < 3; ++a) r[(RF_ATGT + a) * n] = atgt[a];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r[(RF_TGT + k) * n] = tgt[k];
        r[(RF_ESUM + k) * n] = esum[k]; r[(RF_EABS + k) * n] = eabs[k];
        r[(RF_EMIN + k) * n] = emin[k]; r[(RF_EMAX + k) * n] = emax[k];
        ii[(IF_RISE_LO + k) * n] = rise_lo[k]; ii[(IF_RISE_HI + k) * n] = rise_hi[k];
    }
    // rings shift: age k -> age k+1, current -> age 1
#pragma unroll
    for (int i = 11; i >= 3; --i) r[(RF_ACT_RING + i) * n] = aring[i - 3];
#pragma unroll
    for (int j = 0; j < 3; ++j) r[(RF_ACT_RING + j) * n] = a_raw[j];
    if (need_cring) {
#pragma unroll
        for (int i = 11; i >= 3; --i) r[(RF_CMD_RING + i) * n] = cring[i - 3];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) r[(RF_CMD_RING + j) * n] = cmd_in[j];
    r[RF_CV_SUM * n] = cv_sum;
    r[RF_EP_RET * n] = ep_ret;
    ii[IF_STEPS * n] = steps; ii[IF_STEPS_TGT * n] = steps_tgt; ii[IF_SIM_STEP * n] = sim_step;
    ii[IF_ACT_F32 * n] = act_f32 ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ii[(IF_GOAL_CNT + k) * n] = gcnt[k]; ii[(IF_GOAL_TOTAL + k) * n] = gtot[k]; ii[(IF_SETTLE + k) * n] = settle[k];
    }
}

// ---- waypoint head (FixedWingAircraft_simple.step, simple_train.py:410-509) ----
template <typename T, bool TURB>
__global__ void __launch_bounds__(128, 4) waypoint_head_kernel(const __grid_constant__ DCfg<T> c, const Soa<T> S,
                                                               const StepIO io, const Scratch<T> W) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (env >= n) return;
    T* r = S.r + env;
    int32_t* ii = S.i + env;
    int fail = W.fail[env];
    T y[FW_NY];
#pragma unroll
    for (int i = 0; i < FW_NY; ++i) y[i] = fail ? r[(RF_Y + i) * n] : W.ytmp[i * n + env];
    T roll = r[RF_ROLL * n], pitch = r[RF_PITCH * n], Va = r[RF_VA * n], alpha = r[RF_ALPHA * n], beta = r[RF_BETA * n];
    T om_obs[3] = {r[(RF_Y + 4) * n], r[(RF_Y + 5) * n], r[(RF_Y + 6) * n]};
    DynCtx<T> x;
    T goal[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x.wind[k] = r[(RF_WIND + k) * n]; goal[k] = r[(RF_TGT + k) * n];
        x.tl[k] = TURB ? W.turb[k * n + env] : (T)0; x.ta[k] = TURB ? W.turb[(3 + k) * n + env] : (T)0;
        x.cmd[k] = 0;
    }
    int steps = ii[IF_STEPS * n], sim_step = ii[IF_SIM_STEP * n];
    const unsigned long long episode = (unsigned long long)(uint32_t)ii[IF_EPISODE * n];
    fail = post_step_commit<T>(c, x, y, fail, roll, pitch, Va, alpha, beta, om_obs);
    sim_step += 1;
    steps += 1;
    bool done = false;
    int term = FW_TERM_NONE;
    if (c.steps_max > 0 && steps >= c.steps_max) { done = true; term = FW_TERM_STEPS; }
    T reward;
    bool teleport = false;
    if (!fail) {
        // commit the simulator state
#pragma unroll
        for (int i = 0; i < FW_NY; ++i) r[(RF_Y + i) * n] = y[i];
        r[RF_ROLL * n] = roll; r[RF_PITCH * n] = pitch; r[RF_VA * n] = Va; r[RF_ALPHA * n] = alpha; r[RF_BETA * n] = beta;
        ii[IF_SIM_STEP * n] = sim_step;
        ii[IF_STEPS_TGT * n] = ii[IF_STEPS_TGT * n] + 1;
        bool all = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) all = all && (M<T>::fabs(goal[k] - y[7 + k]) <= c.wp_goal_bound[k]);
        if (all) {                                   // sample_task(idx): next leg, wrapping (simple_train.py:346-355)
            int pos = ii[IF_WP_POS * n];
            pos = (pos < S.wp_len - 2) ? pos + 1 : 0;
            wp_start_leg<T>(c, S, env, pos);
            teleport = true;
#pragma unroll
            for (int k = 0; k < 3; ++k) { goal[k] = r[(RF_TGT + k) * n]; y[7 + k] = r[(RF_Y + 7 + k) * n]; }
        }
        T sacc = 0;                                  // get_reward (simple_train.py:673-690), after the teleport
#pragma unroll
        for (int k = 0; k < 3; ++k) sacc += (T)1 * (M<T>::fabs(goal[k] - y[7 + k]) / c.wp_rew_range[k]);
        reward = (T)1 / M<T>::exp(sacc);
        if (TURB && !teleport) {
            T fx[12], fu[4], un[4];
#pragma unroll
            for (int i = 0; i < 12; ++i) fx[i] = r[(RF_FX + i) * n];
#pragma unroll
            for (int i = 0; i < 4; ++i) fu[i] = r[(RF_FU + i) * n];
            noise_sample(c, S, env, episode, sim_step, un);
            turb_step(c, fx, fu, un, sim_step);
#pragma unroll
            for (int i = 0; i < 12; ++i) r[(RF_FX + i) * n] = fx[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[(RF_FU + i) * n] = fu[i];
        }
    } else {
        done = true;
        term = fail;
        reward = (T)(steps - c.steps_max);
    }
    const T ep_ret = r[RF_EP_RET * n] + reward;
    r[RF_EP_RET * n] = ep_ret;
    ii[IF_STEPS * n] = steps;
    T o[FW_NOBS_WAYPOINT];
    wp_observation<T>(S, env, o);
    if (done) {
        double* m = S.metrics + (size_t)env * FW_NMETRIC;
        for (int k = 0; k < FW_NMETRIC; ++k) m[k] = CUDART_NAN;     // this env computes no metrics (simple_train.py:501-503)
        S.ep_ret[env] = (double)ep_ret;
        S.ep_len[env] = steps;
    }
    S.ep_term[env] = term;
    if (io.rew) io.rew[env] = (float)reward;
    if (io.rew64) io.rew64[env] = (double)reward;
    if (io.done) io.done[env] = done ? 1 : 0;
    if (done && io.auto_reset) {
        if (io.term_obs) write_obs(o, FW_NOBS_WAYPOINT, env, io.term_obs, (double*)nullptr);
        if (io.info) {      // packed episode-end row for the host (fw_set_info_rows), same layout as head_kernel's
            const int slot = atomicAdd(reinterpret_cast<int32_t*>(io.info), 1);
            if (slot < io.info_cap) {
                double* row = io.info + 1 + (size_t)slot * (FW_INFO_HEAD + FW_NOBS_WAYPOINT);
                row[0] = (double)env; row[1] = (double)term; row[2] = (double)steps; row[3] = (double)ep_ret;
                for (int k = 0; k < FW_NMETRIC; ++k) row[4 + k] = CUDART_NAN;
                for (int q = 0; q < FW_NOBS_WAYPOINT; ++q) row[FW_INFO_HEAD + q] = (double)(float)o[q];
            }
        }
        wp_reset_env<T>(c, S, env, io.obs, io.obs64);
        return;
    }
    write_obs(o, FW_NOBS_WAYPOINT, env, io.obs, io.obs64);
}

// gather / scatter between SoA fields and row-major [n, width] buffers
template <typename T>
__global__ void field_copy_kernel(const __grid_constant__ DCfg<T> c, Soa<T> S, int field, void* buf, int to_soa) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = S.n;
    if (env >= n) return;
    double* d = (double*)buf;
    int32_t* di = (int32_t*)buf;
    T* r = S.r + env;
    int32_t* ii = S.i + env;
    auto rw = [&](int rf, int width, int col0, int stride) {
        for (int k = 0; k < width; ++k) {
            if (to_soa) r[(rf + k) * n] = (T)d[(size_t)env * stride + col0 + k];
            else d[(size_t)env * stride + col0 + k] = (double)r[(rf + k) * n];
        }
    };
    switch (field) {
        case FW_FIELD_Y: rw(RF_Y, FW_NY, 0, FW_NY); break;
        case FW_FIELD_EULER:
            // yaw is not stored (recomputed every step); report roll, pitch, NaN
            rw(RF_ROLL, 2, 0, 3);
            if (!to_soa) d[(size_t)env * 3 + 2] = CUDART_NAN;
            break;
        case FW_FIELD_VAB: rw(RF_VA, 3, 0, 3); break;
        case FW_FIELD_WIND: rw(RF_WIND, 3, 0, 3); break;
        case FW_FIELD_TARGET: rw(RF_TGT, 3, 0, 3); break;
        case FW_FIELD_ATARGET: rw(RF_ATGT, 3, 0, 3); break;
        case FW_FIELD_CMD: rw(RF_CMD_RING, 3, 0, 3); break;
        case FW_FIELD_TURB:
            if (!to_soa) {
                T fx[12], fu[4], tl[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
                for (int i = 0; i < 12; ++i) fx[i] = r[(RF_FX + i) * n];
                for (int i = 0; i < 4; ++i) fu[i] = r[(RF_FU + i) * n];
                if (c.turbulence) turb_eval(c, fx, fu, tl, ta);
                for (int i = 0; i < 3; ++i) { d[(size_t)env * 6 + i] = (double)tl[i]; d[(size_t)env * 6 + 3 + i] = (double)ta[i]; }
            }
            break;
        case FW_FIELD_COUNTERS:
            for (int k = 0; k < 4; ++k) {
                const int f = (k == 0) ? IF_STEPS : (k == 1) ? IF_STEPS_TGT : (k == 2) ? IF_SIM_STEP : IF_EPISODE;
                if (to_soa) ii[f * n] = di[(size_t)env * 4 + k];
                else di[(size_t)env * 4 + k] = ii[f * n];
            }
            break;
        case FW_FIELD_NFEV:
            if (!to_soa) { di[(size_t)env * 2] = ii[IF_NFEV * n]; di[(size_t)env * 2 + 1] = ii[IF_NATT * n]; }
            break;
        case FW_FIELD_PARAMS:
            if (S.par) {
                if (to_soa) {
                    double bp[FW_NPARAM];
                    for (int k = 0; k < FW_NPARAM; ++k) bp[k] = d[(size_t)env * FW_NPARAM + k];
                    write_env_params<T>(c, S, env, bp);
                } else {
                    for (int k = 0; k < FW_NPARAM; ++k) d[(size_t)env * FW_NPARAM + k] = (double)S.par[(size_t)k * n + env];
                }
            }
            break;
        default: break;
    }
}

__global__ void episode_info_kernel(int n, const int32_t* term, const double* metrics, const double* ep_ret,
                                    const int32_t* ep_len, int32_t* term_out, double* metrics_out, double* ret_out,
                                    int32_t* len_out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    if (term_out) term_out[env] = term[env];
    if (ret_out) ret_out[env] = ep_ret[env];
    if (len_out) len_out[env] = ep_len[env];
    if (metrics_out)
        for (int k = 0; k < FW_NMETRIC; ++k) metrics_out[(size_t)env * FW_NMETRIC + k] = metrics[(size_t)env * FW_NMETRIC + k];
}

// ---------------------------------------------------------------------------------------------------------------
// host side
static thread_local char g_err[512] = "";
static void set_err(const char* what, cudaError_t e) { snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e)); }
#define CK(call)                                   \
    do {                                           \
        cudaError_t e_ = (call);                   \
        if (e_ != cudaSuccess) { set_err(#call, e_); return FW_ECUDA; } \
    } while (0)

template <typename T> static void convert_cfg(const FwConfig& f, DCfg<T>& d) {
    memset(&d, 0, sizeof(d));
    d.integrator = f.integrator; d.rk4_substeps = f.rk4_substeps; d.turbulence = f.turbulence;
    d.steps_max = f.steps_max; d.scale_actions = f.scale_actions; d.has_action_bounds = f.has_action_bounds;
    for (int k = 0; k < 3; ++k) d.tgt_class[k] = f.tgt_class[k];
    d.on_success = f.on_success; d.streak_req = f.streak_req; d.resample_every = f.resample_every;
    d.rew_delta_window = f.rew_delta_window; d.obs_act_window = f.obs_act_window;
    d.step_fail_timesteps = f.step_fail_timesteps;
    d.rtol = (T)f.rtol; d.atol = (T)f.atol;
#define CP(x) d.x = (T)f.x
    CP(mass); CP(Jy); CP(S_wing); CP(b); CP(c); CP(k_motor); CP(k_T_P); CP(k_Omega); CP(a_0);
    d.M_ = (T)f.M;
    d.ar = (T)(f.b * f.b / f.S_wing);
    CP(C_L_0); CP(C_L_alpha); CP(C_L_q); CP(C_L_delta_e); CP(C_D_p); CP(C_D_q); CP(C_D_beta1); CP(C_D_beta2); CP(C_D_delta_e);
    CP(C_m_0); CP(C_m_alpha); CP(C_m_q); CP(C_m_delta_e); CP(C_m_fp);
    CP(C_Y_0); CP(C_Y_beta); CP(C_Y_p); CP(C_Y_r); CP(C_Y_delta_a); CP(C_Y_delta_r);
    CP(C_l_0); CP(C_l_beta); CP(C_l_p); CP(C_l_r); CP(C_l_delta_a); CP(C_l_delta_r);
    CP(C_n_0); CP(C_n_beta); CP(C_n_p); CP(C_n_r); CP(C_n_delta_a); CP(C_n_delta_r);
    {   // gammas (pyfly.py:1099-1116), evaluated in double on the host
        const double I00 = f.Jx, I11 = f.Jy, I22 = f.Jz, I02 = -f.Jxz;
        double g[9];
        g[0] = I00 * I22 - I02 * I02;
        g[1] = (fabs(I02) * (I00 - I11 + I22)) / g[0];
        g[2] = (I22 * (I22 - I11) + I02 * I02) / g[0];
        g[3] = I22 / g[0];
        g[4] = fabs(I02) / g[0];
        g[5] = (I22 - I00) / I11;
        g[6] = fabs(I02) / I11;
        g[7] = ((I00 - I11) * I00 + I02 * I02) / g[0];
        g[8] = I00 / g[0];
        for (int k = 0; k < 9; ++k) d.gam[k] = (T)g[k];
    }
    d.half_rho = (T)(0.5 * f.rho);
    d.mg = (T)(f.mass * f.g);
    d.g_ = (T)f.g;
    d.model_on = f.model_on;
    d.prop_k = (T)(0.5 * f.rho * f.S_prop * f.C_prop);
    d.inv_pi_e_ar = (T)(1.0 / (3.14159265358979323846 * f.e_oswald * (f.b * f.b / f.S_wing)));
    d.inv_Jy = (T)(1.0 / f.Jy);
    d.inv_mass = (T)(1.0 / f.mass);
    d.exp_M_a0 = (T)exp(f.M * f.a_0);
    CP(dt); CP(elevon_min); CP(elevon_max); CP(elevon_dot_max); CP(throttle_min); CP(throttle_max);
    d.w0sq = (T)(f.elevon_omega0 * f.elevon_omega0);
    d.two_zeta_w0 = (T)(2 * f.elevon_zeta * f.elevon_omega0);
    d.inv_tau = (T)(1 / f.throttle_tau);
    for (int k = 0; k < 3; ++k) { CP(omega_con_min[k]); CP(omega_con_max[k]); }
    CP(va_value_min); CP(va_con_max);
    for (int k = 0; k < 12; ++k) { CP(init_lo[k]); CP(init_hi[k]); }
    CP(wind_mag_min); CP(wind_mag_max); CP(turb_noise_scale);
    for (int fi = 0; fi < 6; ++fi) {
        d.filt[fi].order = f.filt[fi].order; d.filt[fi].noise_row = f.filt[fi].noise_row;
        for (int k = 0; k < 9; ++k) { d.filt[fi].Ad[k] = (T)f.filt[fi].Ad[k]; d.filt[fi].Ablk[k] = (T)f.filt[fi].Ablk[k]; }
        for (int k = 0; k < 3; ++k) { d.filt[fi].Bd0[k] = (T)f.filt[fi].Bd0[k]; d.filt[fi].Bd1[k] = (T)f.filt[fi].Bd1[k]; d.filt[fi].C[k] = (T)f.filt[fi].C[k]; }
        d.filt[fi].D = (T)f.filt[fi].D;
    }
    CP(scale_low); CP(scale_high);
    for (int k = 0; k < 3; ++k) {
        CP(act_lo[k]); CP(act_hi[k]); CP(action_bounds_min[k]); CP(action_bounds_max[k]);
        CP(tgt_low[k]); CP(tgt_high[k]); CP(tgt_delta[k]); CP(tgt_bound[k]); CP(rew_err_scaling[k]); CP(rew_err_max[k]);
        d.tgt_radians[k] = f.tgt_radians[k];
        CP(tgt_slope_low[k]); CP(tgt_slope_high[k]); CP(tgt_amp_low[k]); CP(tgt_amp_high[k]); CP(tgt_period_low[k]); CP(tgt_period_high[k]);
        if (f.tgt_class[k] == FW_TGT_LINEAR || f.tgt_class[k] == FW_TGT_SINUSOIDAL) d.tgt_moving = 1;
    }
    CP(rng_u_override);
    d.env_kind = f.env_kind; d.turb_block_len = f.turb_block_len;
    for (int k = 0; k < 3; ++k) { CP(wp_goal_bound[k]); CP(wp_rew_range[k]); }
    CP(streak_fraction); CP(rew_delta_scaling); CP(rew_delta_max); CP(rew_bound_scaling); CP(rew_bound_max);
    CP(step_fail_value); CP(rise_low); CP(rise_high); CP(obs_noise_mean); CP(obs_noise_std); CP(obs_init_noise);
    d.rew_generic = f.rew_generic; d.rew_n = f.rew_n; d.rew_potential = f.rew_potential; d.rew_nterms = f.rew_nterms;
    for (int k = 0; k < FW_REW_FACTORS_MAX; ++k) {
        d.rew_class[k] = f.rew_class[k]; d.rew_idx[k] = f.rew_idx[k]; d.rew_fclass[k] = f.rew_fclass[k];
        d.rew_shaping[k] = f.rew_shaping[k]; d.rew_window[k] = f.rew_window[k]; d.rew_value_timesteps[k] = f.rew_value_timesteps[k];
        CP(rew_scaling[k]); CP(rew_maxv[k]); CP(rew_sign[k]); CP(rew_value[k]);
    }
    for (int k = 0; k < 4; ++k) { d.term_fclass[k] = f.term_fclass[k]; CP(term_weight[k]); }
    d.obs_generic = f.obs_generic; d.obs_len = f.obs_len; d.obs_n = f.obs_n; d.obs_normalize = f.obs_normalize;
    for (int k = 0; k < FW_OBS_ENTRIES_MAX; ++k) {
        d.obs_kind[k] = f.obs_kind[k]; d.obs_idx[k] = f.obs_idx[k]; d.obs_window[k] = f.obs_window[k];
        d.obs_norm_flag[k] = f.obs_norm_flag[k]; CP(obs_mean[k]); CP(obs_var[k]);
    }
    d.ang_on = f.ang_on;
    for (int k = 0; k < 3; ++k) { CP(ang_max_vel[k]); CP(ang_bound[k]); }
    d.integration_window = f.integration_window;
    d.obs_step = f.obs_step > 0 ? f.obs_step : 1;
    d.obs_has_int = 0;
    for (int k = 0; k < f.obs_n && f.obs_generic; ++k) d.obs_has_int |= (f.obs_kind[k] == FW_OBS_TARGET_INT);
#undef CP
    d.seed = f.seed;
    d.env_id_offset = f.env_id_offset;
}

template <typename T> static void convert_reset_cfg(const FwConfig& f, ResetCfg<T>& d) {
    memset(&d, 0, sizeof(d));
#define CP(x) d.x = (T)f.x
    for (int k = 0; k < 12; ++k) { CP(init_lo[k]); CP(init_hi[k]); }
    CP(wind_mag_min); CP(wind_mag_max);
    for (int k = 0; k < 3; ++k) {
        CP(tgt_low[k]); CP(tgt_high[k]); CP(tgt_delta[k]);
        CP(tgt_slope_low[k]); CP(tgt_slope_high[k]); CP(tgt_amp_low[k]); CP(tgt_amp_high[k]); CP(tgt_period_low[k]); CP(tgt_period_high[k]);
    }
#undef CP
    d.seed = f.seed;
    d.model_uniform = f.model_uniform;
    for (int i = 0; i < FW_NPARAM; ++i) {
        d.par_enabled[i] = f.par_enabled[i]; d.par_orig[i] = f.par_orig[i]; d.par_var[i] = f.par_var[i]; d.par_clip[i] = f.par_clip[i];
    }
}

// the FwConfig fields fw_set_config may change on a live handle (everything ResetCfg carries); used to check that
// nothing else differs
static void blank_reset_fields(FwConfig& f) {
    memset(f.init_lo, 0, sizeof(f.init_lo)); memset(f.init_hi, 0, sizeof(f.init_hi));
    f.wind_mag_min = f.wind_mag_max = 0;
    memset(f.tgt_low, 0, sizeof(f.tgt_low)); memset(f.tgt_high, 0, sizeof(f.tgt_high)); memset(f.tgt_delta, 0, sizeof(f.tgt_delta));
    memset(f.tgt_slope_low, 0, sizeof(f.tgt_slope_low)); memset(f.tgt_slope_high, 0, sizeof(f.tgt_slope_high));
    memset(f.tgt_amp_low, 0, sizeof(f.tgt_amp_low)); memset(f.tgt_amp_high, 0, sizeof(f.tgt_amp_high));
    memset(f.tgt_period_low, 0, sizeof(f.tgt_period_low)); memset(f.tgt_period_high, 0, sizeof(f.tgt_period_high));
    f.seed = 0;
    f.model_uniform = 0;
    memset(f.par_enabled, 0, sizeof(f.par_enabled)); memset(f.par_orig, 0, sizeof(f.par_orig));
    memset(f.par_var, 0, sizeof(f.par_var)); memset(f.par_clip, 0, sizeof(f.par_clip));
}

}  // namespace fw

using namespace fw;

struct FwHandle {
    FwConfig cfg;
    int n, device;
    DCfg<double> c64;
    DCfg<float> c32;
    Soa<double> s64;
    Soa<float> s32;
    void* r_buf; int32_t* i_buf; void* err_ring;
    void* err_ring_a; void* err2a; double* metrics_a;     // attitude_angular targets only (else nullptr)
    double* metrics; double* ep_ret; int32_t* ep_len; int32_t* ep_term;
    void* w_real; int32_t* w_int;          // scratch between the kernels of one step
    double* wp_tasks; int32_t* wp_task_of_env;   // waypoint head: device copies of the task table
    void* rc_dev;                          // ResetCfg<T> in device memory (fw_set_config rewrites it)
    void* par_buf; void* par2_buf;         // per-env aircraft parameters and their next-episode rows (model_on)
    Scratch<double> w64;
    Scratch<float> w32;
    int sm_count, att_blocks_per_sm;
    unsigned long long random_step;
    // precomputed next-episode rows (Spare) and the side stream that refills them
    void* r2_buf; int32_t* i2_buf; void* err2; float* spare_obs; double* spare_obs64; int32_t* done_list;
    Spare<double> p64;
    Spare<float> p32;
    cudaStream_t side;
    cudaEvent_t ev_head, ev_refill;
    int refill_pending, refill_captured, step_parity;
    double* info_rows; int info_cap;       // fw_set_info_rows
    // fw_set_profiling: CUDA events around each kernel of a step (the step then ends with an event synchronise)
    int prof_on;
    cudaEvent_t prof_ev[4];
    double prof_ms[3];
    long long prof_steps;
};

// event i is recorded BEFORE kernel i of the step (and event 3 after the last one)
static inline void prof_mark(FwHandle* h, int i, cudaStream_t st) {
    if (h->prof_on) cudaEventRecord(h->prof_ev[i], st);
}
static void prof_collect(FwHandle* h, bool has_init) {
    if (!h->prof_on) return;
    cudaEventSynchronize(h->prof_ev[3]);
    float ms = 0;
    for (int i = has_init ? 0 : 1; i < 3; ++i) {
        cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]);
        h->prof_ms[i] += ms;
    }
    h->prof_steps += 1;
}

static const int NT_RK45_F64 = 32;     // 6*16*32*8 = 24576 B of stage storage per one-warp block: 8 blocks / SM fit
static const int NT_RK45_F32 = 64;     // 6*16*64*4 = 24576 B
static const int W_REAL_FIELDS = FW_NK + 1 + 3 + 6 + FW_NY;   // f0, hinit, cmd, turb, ytmp

template <typename T> static Scratch<T> make_scratch(void* real, int32_t* ints, size_t n) {
    T* p = (T*)real;
    Scratch<T> w;
    w.f0 = p; p += FW_NK * n;
    w.hinit = p; p += n;
    w.cmd = p; p += 3 * n;
    w.turb = p; p += 6 * n;
    w.ytmp = p;
    w.fail = ints;
    w.counter = ints + n;
    return w;
}

// head_kernel of this step reads the spare rows: the refill of the previous step's consumers must have finished
static inline bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive;
}
static inline void spare_join(FwHandle* h, cudaStream_t st) {
    if (!h->refill_pending) return;
    h->refill_pending = 0;
    // A capturing stream cannot wait for an event recorded before the capture began (and a captured event means
    // nothing outside its graph).  Whoever begins a capture has synchronised with the work before it —
    // torch.cuda.graph() synchronises the device — or has called fw_join; inside a capture every step joins its own
    // refill (spare_refill), so nothing captured is ever pending when the capture ends.
    if (stream_is_capturing(st) != (h->refill_captured != 0)) return;
    cudaStreamWaitEvent(st, h->ev_refill, 0);
}
// after head_kernel: recompute, on the side stream, the rows this step consumed (it overlaps the next step)
template <typename T>
static void spare_refill(FwHandle* h, const DCfg<T>& c, const Spare<T>& P, cudaStream_t st) {
    if (!P.on) return;
    cudaEventRecord(h->ev_head, st);
    cudaStreamWaitEvent(h->side, h->ev_head, 0);
    int grid = (h->n + 31) / 32;
    if (grid > 4 * h->sm_count) grid = 4 * h->sm_count;
    refill_kernel<T><<<grid, 32, 0, h->side>>>(c, P, h->step_parity);
    cudaEventRecord(h->ev_refill, h->side);
    h->refill_pending = 1;
    h->step_parity ^= 1;
    // inside a stream capture the side-stream work is joined right away: a capture must not end with unjoined work,
    // and the caller may end it after any step (the refill then sits between this step and the next in the graph)
    h->refill_captured = stream_is_capturing(st) ? 1 : 0;
    if (h->refill_captured) spare_join(h, st);
}

// head_kernel<.., GENERIC = false> is the straight-line head of the default task family (default observation row, default
// reward family, constant / compensate targets, no resampling); everything else — general observation layout, general
// reward engine, moving targets, periodic or on-success resampling — lives in the GENERIC instantiation only, so that
// its cold code does not cost the default kernel registers (782 -> 1006 B of spills and +5 us when it did).
static inline bool head_generic(const FwConfig& f) {
    bool moving = false;
    for (int k = 0; k < 3; ++k) moving |= (f.tgt_class[k] == FW_TGT_LINEAR || f.tgt_class[k] == FW_TGT_SINUSOIDAL);
    return f.obs_generic || f.rew_generic || moving || f.resample_every > 0 || f.on_success == FW_SUCCESS_NEW || f.ang_on;
}

// One env step = init kernel -> persistent attempt kernel -> head kernel (RK45), or rk4 kernel -> head kernel.
template <typename T, bool TURB, int NT, bool PE>
static int launch_rk45(FwHandle* h, const DCfg<T>& c, const Soa<T>& S, const Scratch<T>& W, const Spare<T>& P, const StepIO& io, cudaStream_t st) {
    const size_t smem = (size_t)6 * FW_NP * 2 * NT * sizeof(T);
    auto k = rk45_attempt_kernel<T, TURB, NT, PE>;
    // per handle: function attributes and occupancy belong to the handle's device (one handle runs one instantiation)
    if (!h->att_blocks_per_sm) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->att_blocks_per_sm, k, NT, smem));
        if (h->att_blocks_per_sm < 1) h->att_blocks_per_sm = 1;
    }
    int grid = h->sm_count * h->att_blocks_per_sm;
    const int need = (h->n + NT - 1) / NT;
    if (grid > need) grid = need;
    const int g0 = (h->n + 127) / 128, g0h = (h->n + 63) / 64;
    prof_mark(h, 0, st);
    const int par = h->step_parity;
    rk45_init_kernel<T, TURB, PE><<<g0, 128, 0, st>>>(c, S, io, W, grid * NT, P.count + par);
    prof_mark(h, 1, st);
    k<<<grid, NT, smem, st>>>(c, S, W);
    prof_mark(h, 2, st);
    spare_join(h, st);
    if (h->cfg.env_kind == FW_ENV_WAYPOINT) waypoint_head_kernel<T, TURB><<<g0, 128, 0, st>>>(c, S, io, W);
    else if (head_generic(h->cfg)) head_kernel<T, TURB, true><<<g0h, 64, 0, st>>>(c, S, io, W, P, par);
    else head_kernel<T, TURB, false><<<g0h, 64, 0, st>>>(c, S, io, W, P, par);
    CK(cudaGetLastError());
    prof_mark(h, 3, st);
    if (io.auto_reset) spare_refill<T>(h, c, P, st);
    CK(cudaGetLastError());
    prof_collect(h, true);
    return FW_OK;
}
template <typename T, bool TURB, bool PE>
static int launch_rk4(FwHandle* h, const DCfg<T>& c, const Soa<T>& S, const Scratch<T>& W, const Spare<T>& P, const StepIO& io, cudaStream_t st) {
    const int g0 = (h->n + 127) / 128, g0h = (h->n + 63) / 64;
    prof_mark(h, 1, st);
    const int par = h->step_parity;
    rk4_kernel<T, TURB, PE><<<g0, 128, 0, st>>>(c, S, io, W, P.count + par);
    prof_mark(h, 2, st);
    spare_join(h, st);
    if (h->cfg.env_kind == FW_ENV_WAYPOINT) waypoint_head_kernel<T, TURB><<<g0, 128, 0, st>>>(c, S, io, W);
    else if (head_generic(h->cfg)) head_kernel<T, TURB, true><<<g0h, 64, 0, st>>>(c, S, io, W, P, par);
    else head_kernel<T, TURB, false><<<g0h, 64, 0, st>>>(c, S, io, W, P, par);
    CK(cudaGetLastError());
    prof_mark(h, 3, st);
    if (io.auto_reset) spare_refill<T>(h, c, P, st);
    CK(cudaGetLastError());
    prof_collect(h, false);
    return FW_OK;
}

static int launch_step(FwHandle* h, const StepIO& io_in, cudaStream_t st) {
    StepIO io = io_in;
    io.info = io.auto_reset ? h->info_rows : nullptr;
    io.info_cap = h->info_cap;
    const bool turb = h->cfg.turbulence != 0, rk45 = h->cfg.integrator == FW_INT_RK45_SCIPY, pe = h->cfg.model_on != 0;
    // per-env aircraft parameters (simulator.model) select their own instantiations; the default path is untouched
#define FW_DISPATCH(T, c, s, w, p, NT)                                                                                  \
    do {                                                                                                                \
        if (rk45) {                                                                                                     \
            if (pe) return turb ? launch_rk45<T, true, NT, true>(h, c, s, w, p, io, st) : launch_rk45<T, false, NT, true>(h, c, s, w, p, io, st);   \
            return turb ? launch_rk45<T, true, NT, false>(h, c, s, w, p, io, st) : launch_rk45<T, false, NT, false>(h, c, s, w, p, io, st);         \
        }                                                                                                               \
        if (pe) return turb ? launch_rk4<T, true, true>(h, c, s, w, p, io, st) : launch_rk4<T, false, true>(h, c, s, w, p, io, st);                 \
        return turb ? launch_rk4<T, true, false>(h, c, s, w, p, io, st) : launch_rk4<T, false, false>(h, c, s, w, p, io, st);                       \
    } while (0)
    if (h->cfg.precision == FW_F64) FW_DISPATCH(double, h->c64, h->s64, h->w64, h->p64, NT_RK45_F64);
    FW_DISPATCH(float, h->c32, h->s32, h->w32, h->p32, NT_RK45_F32);
#undef FW_DISPATCH
}

extern "C" int fw_obs_dim(const FwHandle* h);

// ---- checkpoint / resume: the whole env state of a handle as one contiguous device blob ----
struct BlobPart { void* ptr; size_t bytes; };
static int blob_parts(FwHandle* h, BlobPart* out) {
    const size_t esz = h->cfg.precision == FW_F64 ? 8 : 4, n = (size_t)h->n;
    const size_t odim = (size_t)fw_obs_dim(h);
    int k = 0;
    out[k++] = {h->r_buf, esz * RF_COUNT * n};
    out[k++] = {h->i_buf, sizeof(int32_t) * IF_COUNT * n};
    out[k++] = {h->err_ring, esz * FW_END_ERR_WINDOW * 3 * n};
    out[k++] = {h->metrics, sizeof(double) * FW_NMETRIC * n};
    if (h->err_ring_a) {
        out[k++] = {h->err_ring_a, esz * FW_END_ERR_WINDOW * 3 * n};
        out[k++] = {h->err2a, esz * 3 * n};
        out[k++] = {h->metrics_a, sizeof(double) * FW_NMETRIC_ANG * n};
    }
    out[k++] = {h->ep_ret, sizeof(double) * n};
    out[k++] = {h->ep_len, sizeof(int32_t) * n};
    out[k++] = {h->ep_term, sizeof(int32_t) * n};
    out[k++] = {h->r2_buf, esz * RF_COUNT * n};
    out[k++] = {h->i2_buf, sizeof(int32_t) * IF_COUNT * n};
    out[k++] = {h->err2, esz * 3 * n};
    out[k++] = {h->spare_obs, sizeof(float) * odim * n};
    out[k++] = {h->spare_obs64, sizeof(double) * odim * n};
    out[k++] = {h->rc_dev, sizeof(ResetCfg<double>)};
    if (h->par_buf) { out[k++] = {h->par_buf, esz * FW_PAR_FIELDS * n}; out[k++] = {h->par2_buf, esz * FW_PAR_FIELDS * n}; }
    return k;
}
struct BlobHeader { uint64_t magic; int32_t abi, n, precision, rf_count, if_count, obs_dim; uint64_t random_step; int32_t step_parity, _pad; };
static const uint64_t BLOB_MAGIC = 0x46574232303053ull;   // "FWB200S"

extern "C" {

const char* fw_last_error(void) { return g_err; }
int fw_abi_version(void) { return FW_ABI_VERSION; }
int fw_config_size(void) { return (int)sizeof(FwConfig); }
int fw_obs_dim(const FwHandle* h) {
    if (!h) return FW_EINVAL;
    if (h->cfg.env_kind == FW_ENV_WAYPOINT) return FW_NOBS_WAYPOINT;
    return h->cfg.obs_generic ? h->cfg.obs_len * h->cfg.obs_n : FW_NOBS;
}

int fw_set_info_rows(FwHandle* h, double* rows_dev, int32_t cap) {
    if (!h || (rows_dev && cap <= 0)) return FW_EINVAL;
    h->info_rows = rows_dev;
    h->info_cap = rows_dev ? cap : 0;
    return FW_OK;
}

int fw_join(FwHandle* h, void* stream) {
    if (!h) return FW_EINVAL;
    spare_join(h, (cudaStream_t)stream);
    return FW_OK;
}

int fw_set_profiling(FwHandle* h, int32_t on) {
    if (!h) return FW_EINVAL;
    if (on && !h->prof_ev[0])
        for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&h->prof_ev[i]));
    h->prof_on = on ? 1 : 0;
    h->prof_steps = 0;
    h->prof_ms[0] = h->prof_ms[1] = h->prof_ms[2] = 0;
    return FW_OK;
}

int fw_get_profile(const FwHandle* h, double* ms_sum3, int64_t* steps) {
    if (!h || !ms_sum3 || !steps) return FW_EINVAL;
    for (int i = 0; i < 3; ++i) ms_sum3[i] = h->prof_ms[i];
    *steps = h->prof_steps;
    return FW_OK;
}

int fw_set_waypoint_tasks(FwHandle* h, const double* tasks_dev, int32_t n_tasks, int32_t wp_len,
                          const int32_t* task_of_env_dev, void* stream) {
    if (!h || !tasks_dev || !task_of_env_dev || n_tasks <= 0 || wp_len < 2 || h->cfg.env_kind != FW_ENV_WAYPOINT) {
        snprintf(g_err, sizeof(g_err), "fw_set_waypoint_tasks: needs a waypoint handle, n_tasks > 0, wp_len >= 2");
        return FW_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaFree(h->wp_tasks); cudaFree(h->wp_task_of_env);
    const size_t bytes = sizeof(double) * (size_t)n_tasks * wp_len * FW_WP_ROW;
    CK(cudaMalloc((void**)&h->wp_tasks, bytes));
    CK(cudaMalloc((void**)&h->wp_task_of_env, sizeof(int32_t) * h->n));
    CK(cudaMemcpyAsync(h->wp_tasks, tasks_dev, bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(h->wp_task_of_env, task_of_env_dev, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, st));
    h->s64.wp_tasks = h->s32.wp_tasks = h->wp_tasks;
    h->s64.wp_task_of_env = h->s32.wp_task_of_env = h->wp_task_of_env;
    h->s64.wp_n_tasks = h->s32.wp_n_tasks = n_tasks;
    h->s64.wp_len = h->s32.wp_len = wp_len;
    return FW_OK;
}

int fw_create(const FwConfig* cfg, int32_t n_envs, int32_t device, FwHandle** out) {
    if (!cfg || !out || n_envs <= 0) { snprintf(g_err, sizeof(g_err), "fw_create: bad arguments"); return FW_EINVAL; }
    if (cfg->abi_version != FW_ABI_VERSION) { snprintf(g_err, sizeof(g_err), "fw_create: ABI version mismatch"); return FW_EINVAL; }
    static const int kNoiseRow[6] = {0, 1, 2, 3, 1, 2}, kOrder[6] = {1, 2, 2, 1, 3, 3};
    for (int f = 0; f < 6; ++f)
        if (cfg->turbulence && (cfg->filt[f].noise_row != kNoiseRow[f] || cfg->filt[f].order != kOrder[f])) {
            snprintf(g_err, sizeof(g_err), "fw_create: Dryden filter %d must have order %d and noise row %d", f, kOrder[f], kNoiseRow[f]);
            return FW_EINVAL;
        }
    if (cfg->streak_req > 128 || cfg->rew_delta_window > 5 || cfg->steps_max <= 0 ||
        (!cfg->obs_generic && cfg->obs_act_window > 5) ||
        (cfg->obs_generic && (cfg->obs_len < 1 || cfg->obs_len > FW_OBS_LEN_MAX || cfg->obs_n < 1 || cfg->obs_n > FW_OBS_ENTRIES_MAX ||
                              cfg->obs_act_window + cfg->obs_len - 1 > 9)) ||
        (cfg->precision != FW_F64 && cfg->precision != FW_F32) ||
        (cfg->integrator != FW_INT_RK45_SCIPY && cfg->integrator != FW_INT_RK4_FIXED)) {
        snprintf(g_err, sizeof(g_err), "fw_create: unsupported config value");
        return FW_EINVAL;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        snprintf(g_err, sizeof(g_err), "fw_create: no CUDA device %d (there is no CPU fallback)", device);
        return FW_ENODEVICE;
    }
    CK(cudaSetDevice(device));
    FwHandle* h = new (std::nothrow) FwHandle();
    if (!h) return FW_ENOMEM;
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg; h->n = n_envs; h->device = device;
    convert_cfg<double>(*cfg, h->c64);
    convert_cfg<float>(*cfg, h->c32);
    const size_t esz = cfg->precision == FW_F64 ? 8 : 4, n = (size_t)n_envs;
    CK(cudaMalloc(&h->r_buf, esz * RF_COUNT * n));
    CK(cudaMalloc((void**)&h->i_buf, sizeof(int32_t) * IF_COUNT * n));
    CK(cudaMalloc(&h->err_ring, esz * FW_END_ERR_WINDOW * 3 * n));
    CK(cudaMalloc((void**)&h->metrics, sizeof(double) * FW_NMETRIC * n));
    CK(cudaMalloc((void**)&h->ep_ret, sizeof(double) * n));
    CK(cudaMalloc((void**)&h->ep_len, sizeof(int32_t) * n));
    CK(cudaMalloc((void**)&h->ep_term, sizeof(int32_t) * n));
    CK(cudaMalloc(&h->w_real, esz * W_REAL_FIELDS * n));
    CK(cudaMalloc((void**)&h->w_int, sizeof(int32_t) * (n + 4)));
    CK(cudaMemset(h->w_real, 0, esz * W_REAL_FIELDS * n));
    CK(cudaMemset(h->w_int, 0, sizeof(int32_t) * (n + 4)));
    h->w64 = make_scratch<double>(h->w_real, h->w_int, n);
    h->w32 = make_scratch<float>(h->w_real, h->w_int, n);
    CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaMemset(h->r_buf, 0, esz * RF_COUNT * n));
    CK(cudaMemset(h->i_buf, 0, sizeof(int32_t) * IF_COUNT * n));
    CK(cudaMemset(h->err_ring, 0, esz * FW_END_ERR_WINDOW * 3 * n));
    CK(cudaMemset(h->metrics, 0, sizeof(double) * FW_NMETRIC * n));
    CK(cudaMemset(h->ep_ret, 0, sizeof(double) * n));
    CK(cudaMemset(h->ep_len, 0, sizeof(int32_t) * n));
    CK(cudaMemset(h->ep_term, 0, sizeof(int32_t) * n));
    h->s64 = Soa<double>{(double*)h->r_buf, h->i_buf, (double*)h->err_ring, h->metrics, h->ep_ret, h->ep_len, h->ep_term, nullptr, 0, n_envs, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr};
    h->s32 = Soa<float>{(float*)h->r_buf, h->i_buf, (float*)h->err_ring, h->metrics, h->ep_ret, h->ep_len, h->ep_term, nullptr, 0, n_envs, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr};
    if (cfg->ang_on) {
        CK(cudaMalloc(&h->err_ring_a, esz * FW_END_ERR_WINDOW * 3 * n));
        CK(cudaMalloc(&h->err2a, esz * 3 * n));
        CK(cudaMalloc((void**)&h->metrics_a, sizeof(double) * FW_NMETRIC_ANG * n));
        CK(cudaMemset(h->err_ring_a, 0, esz * FW_END_ERR_WINDOW * 3 * n));
        CK(cudaMemset(h->err2a, 0, esz * 3 * n));
        CK(cudaMemset(h->metrics_a, 0, sizeof(double) * FW_NMETRIC_ANG * n));
        h->s64.err_ring_a = (double*)h->err_ring_a; h->s32.err_ring_a = (float*)h->err_ring_a;
        h->s64.metrics_a = h->s32.metrics_a = h->metrics_a;
    }
    {
        CK(cudaMalloc(&h->rc_dev, sizeof(ResetCfg<double>)));
        if (cfg->precision == FW_F64) {
            ResetCfg<double> rc; convert_reset_cfg<double>(*cfg, rc);
            CK(cudaMemcpy(h->rc_dev, &rc, sizeof(rc), cudaMemcpyHostToDevice));
        } else {
            ResetCfg<float> rc; convert_reset_cfg<float>(*cfg, rc);
            CK(cudaMemcpy(h->rc_dev, &rc, sizeof(rc), cudaMemcpyHostToDevice));
        }
        h->s64.rc = (const ResetCfg<double>*)h->rc_dev;
        h->s32.rc = (const ResetCfg<float>*)h->rc_dev;
        if (cfg->model_on) {
            CK(cudaMalloc(&h->par_buf, esz * FW_PAR_FIELDS * n));
            CK(cudaMalloc(&h->par2_buf, esz * FW_PAR_FIELDS * n));
            CK(cudaMemset(h->par_buf, 0, esz * FW_PAR_FIELDS * n));
            CK(cudaMemset(h->par2_buf, 0, esz * FW_PAR_FIELDS * n));
            h->s64.par = (double*)h->par_buf; h->s32.par = (float*)h->par_buf;
        }
    }
    {
        const int odim = fw_obs_dim(h);
        CK(cudaMalloc(&h->r2_buf, esz * RF_COUNT * n));
        CK(cudaMalloc((void**)&h->i2_buf, sizeof(int32_t) * IF_COUNT * n));
        CK(cudaMalloc(&h->err2, esz * 3 * n));
        CK(cudaMalloc((void**)&h->spare_obs, sizeof(float) * odim * n));
        CK(cudaMalloc((void**)&h->spare_obs64, sizeof(double) * odim * n));
        CK(cudaMalloc((void**)&h->done_list, sizeof(int32_t) * (2 * n + 2)));
        CK(cudaMemset(h->r2_buf, 0, esz * RF_COUNT * n));
        CK(cudaMemset(h->i2_buf, 0, sizeof(int32_t) * IF_COUNT * n));
        CK(cudaMemset(h->err2, 0, esz * 3 * n));
        CK(cudaMemset(h->spare_obs, 0, sizeof(float) * odim * n));
        CK(cudaMemset(h->spare_obs64, 0, sizeof(double) * odim * n));
        CK(cudaMemset(h->done_list, 0, sizeof(int32_t) * (2 * n + 2)));
        h->p64.S2 = h->s64; h->p64.S2.r = (double*)h->r2_buf; h->p64.S2.i = h->i2_buf; h->p64.S2.err_ring = (double*)h->err2;
        h->p64.S2.par = (double*)h->par2_buf;
        h->p64.S2.err_ring_a = (double*)h->err2a;
        h->p32.S2 = h->s32; h->p32.S2.r = (float*)h->r2_buf; h->p32.S2.i = h->i2_buf; h->p32.S2.err_ring = (float*)h->err2;
        h->p32.S2.par = (float*)h->par2_buf;
        h->p32.S2.err_ring_a = (float*)h->err2a;
        h->p64.obs = h->p32.obs = h->spare_obs;
        h->p64.obs64 = h->p32.obs64 = h->spare_obs64;
        h->p64.list = h->p32.list = h->done_list;
        h->p64.count = h->p32.count = h->done_list + 2 * n;
        h->p64.on = h->p32.on = cfg->env_kind == FW_ENV_WAYPOINT ? 0 : 1;
        // highest priority: the few refill blocks must get an SM slot BEFORE the persistent integrator blocks of the
        // next step fill every register file, or they would only run in that kernel's tail and delay its head kernel
        int prio_least = 0, prio_greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_greatest));
        CK(cudaEventCreateWithFlags(&h->ev_head, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_refill, cudaEventDisableTiming));
    }
    CK(cudaDeviceSynchronize());
    *out = h;
    return FW_OK;
}

int fw_destroy(FwHandle* h) {
    if (!h) return FW_EINVAL;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->r_buf); cudaFree(h->i_buf); cudaFree(h->err_ring); cudaFree(h->metrics);
    cudaFree(h->err_ring_a); cudaFree(h->err2a); cudaFree(h->metrics_a);
    cudaFree(h->ep_ret); cudaFree(h->ep_len); cudaFree(h->ep_term); cudaFree(h->w_real); cudaFree(h->w_int);
    cudaFree(h->wp_tasks); cudaFree(h->wp_task_of_env); cudaFree(h->rc_dev); cudaFree(h->par_buf); cudaFree(h->par2_buf);
    cudaFree(h->r2_buf); cudaFree(h->i2_buf); cudaFree(h->err2); cudaFree(h->spare_obs); cudaFree(h->spare_obs64);
    cudaFree(h->done_list);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_head) cudaEventDestroy(h->ev_head);
    if (h->ev_refill) cudaEventDestroy(h->ev_refill);
    for (int i = 0; i < 4; ++i) if (h->prof_ev[i]) cudaEventDestroy(h->prof_ev[i]);
    delete h;
    return FW_OK;
}

int fw_set_config(FwHandle* h, const FwConfig* cfg, void* stream) {
    if (!h || !cfg) { snprintf(g_err, sizeof(g_err), "fw_set_config: null handle/config"); return FW_EINVAL; }
    {
        FwConfig a = h->cfg, b = *cfg;
        blank_reset_fields(a); blank_reset_fields(b);
        if (memcmp(&a, &b, sizeof(FwConfig)) != 0) {
            snprintf(g_err, sizeof(g_err), "fw_set_config: only the reset-time fields (init_lo/hi, wind_mag_min/max, tgt_low/high/delta, "
                                           "tgt_slope/amp/period ranges, seed) may differ from the handle's configuration");
            return FW_EINVAL;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (stream_is_capturing(st)) { snprintf(g_err, sizeof(g_err), "fw_set_config: not allowed during stream capture"); return FW_EINVAL; }
    spare_join(h, st);       // a refill in flight still reads the old values
    // pageable source: cudaMemcpyAsync stages it before returning, so the stack copy may go out of scope
    if (h->cfg.precision == FW_F64) {
        ResetCfg<double> rc; convert_reset_cfg<double>(*cfg, rc);
        CK(cudaMemcpyAsync(h->rc_dev, &rc, sizeof(rc), cudaMemcpyHostToDevice, st));
    } else {
        ResetCfg<float> rc; convert_reset_cfg<float>(*cfg, rc);
        CK(cudaMemcpyAsync(h->rc_dev, &rc, sizeof(rc), cudaMemcpyHostToDevice, st));
    }
    h->cfg = *cfg;
    convert_cfg<double>(*cfg, h->c64);
    convert_cfg<float>(*cfg, h->c32);
    if (h->p64.on) {
        const int bs = 128, grid = (h->n + bs - 1) / bs;
        if (h->cfg.precision == FW_F64) respare_kernel<double><<<grid, bs, 0, st>>>(h->c64, h->s64, h->p64);
        else respare_kernel<float><<<grid, bs, 0, st>>>(h->c32, h->s32, h->p32);
        CK(cudaGetLastError());
    }
    return FW_OK;
}

int64_t fw_state_blob_size(const FwHandle* h) {
    if (!h) return FW_EINVAL;
    BlobPart parts[24];
    const int k = blob_parts(const_cast<FwHandle*>(h), parts);
    size_t total = sizeof(BlobHeader);
    for (int i = 0; i < k; ++i) total += (parts[i].bytes + 15) & ~(size_t)15;
    return (int64_t)total;
}

int fw_get_state_blob(FwHandle* h, void* blob_dev, void* stream) {
    if (!h || !blob_dev) return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (stream_is_capturing(st)) { snprintf(g_err, sizeof(g_err), "fw_get_state_blob: not allowed during stream capture"); return FW_EINVAL; }
    spare_join(h, st);                    // the precomputed rows being refilled belong to the state
    BlobHeader hd;
    memset(&hd, 0, sizeof(hd));
    hd.magic = BLOB_MAGIC; hd.abi = FW_ABI_VERSION; hd.n = h->n; hd.precision = h->cfg.precision;
    hd.rf_count = RF_COUNT; hd.if_count = IF_COUNT; hd.obs_dim = fw_obs_dim(h);
    hd.random_step = h->random_step; hd.step_parity = h->step_parity;
    CK(cudaMemcpyAsync(blob_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, st));     // pageable source: staged before return
    BlobPart parts[24];
    const int k = blob_parts(h, parts);
    char* p = (char*)blob_dev + sizeof(BlobHeader);
    for (int i = 0; i < k; ++i) {
        CK(cudaMemcpyAsync(p, parts[i].ptr, parts[i].bytes, cudaMemcpyDeviceToDevice, st));
        p += (parts[i].bytes + 15) & ~(size_t)15;
    }
    return FW_OK;
}

int fw_set_state_blob(FwHandle* h, const void* blob_dev, void* stream) {
    if (!h || !blob_dev) return FW_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (stream_is_capturing(st)) { snprintf(g_err, sizeof(g_err), "fw_set_state_blob: not allowed during stream capture"); return FW_EINVAL; }
    BlobHeader hd;
    CK(cudaMemcpyAsync(&hd, blob_dev, sizeof(hd), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hd.magic != BLOB_MAGIC || hd.abi != FW_ABI_VERSION || hd.n != h->n || hd.precision != h->cfg.precision ||
        hd.rf_count != RF_COUNT || hd.if_count != IF_COUNT || hd.obs_dim != fw_obs_dim(h)) {
        snprintf(g_err, sizeof(g_err), "fw_set_state_blob: the blob was written by a handle of another shape (n %d, precision %d, "
                                       "obs_dim %d, abi %d)", hd.n, hd.precision, hd.obs_dim, hd.abi);
        return FW_EINVAL;
    }
    spare_join(h, st);
    BlobPart parts[24];
    const int k = blob_parts(h, parts);
    const char* p = (const char*)blob_dev + sizeof(BlobHeader);
    for (int i = 0; i < k; ++i) {
        CK(cudaMemcpyAsync(parts[i].ptr, p, parts[i].bytes, cudaMemcpyDeviceToDevice, st));
        p += (parts[i].bytes + 15) & ~(size_t)15;
    }
    h->random_step = hd.random_step;
    h->step_parity = hd.step_parity;
    // the reset-time configuration travelled with the blob (device copy): bring the host copy in line
    {
        ResetCfg<double> rcd; ResetCfg<float> rcf;
        if (h->cfg.precision == FW_F64) CK(cudaMemcpyAsync(&rcd, h->rc_dev, sizeof(rcd), cudaMemcpyDeviceToHost, st));
        else CK(cudaMemcpyAsync(&rcf, h->rc_dev, sizeof(rcf), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
#define BACK(x) h->cfg.x = (h->cfg.precision == FW_F64) ? (double)rcd.x : (double)rcf.x
        for (int q = 0; q < 12; ++q) { BACK(init_lo[q]); BACK(init_hi[q]); }
        BACK(wind_mag_min); BACK(wind_mag_max);
        for (int q = 0; q < 3; ++q) {
            BACK(tgt_low[q]); BACK(tgt_high[q]); BACK(tgt_delta[q]); BACK(tgt_slope_low[q]); BACK(tgt_slope_high[q]);
            BACK(tgt_amp_low[q]); BACK(tgt_amp_high[q]); BACK(tgt_period_low[q]); BACK(tgt_period_high[q]);
        }
#undef BACK
        h->cfg.seed = (h->cfg.precision == FW_F64) ? rcd.seed : rcf.seed;
        convert_cfg<double>(h->cfg, h->c64);
        convert_cfg<float>(h->cfg, h->c32);
    }
    return FW_OK;
}

int fw_reset(FwHandle* h, const uint8_t* mask_dev, const double* state_dev, const double* target_dev,
             const double* noise_dev, int32_t noise_len, float* obs_dev, double* obs64_dev, void* stream) {
    if (!h) return FW_EINVAL;
    if (noise_dev && noise_len <= 0) { snprintf(g_err, sizeof(g_err), "fw_reset: noise_len must be > 0"); return FW_EINVAL; }
    if (h->cfg.env_kind == FW_ENV_WAYPOINT && !h->wp_tasks) { snprintf(g_err, sizeof(g_err), "fw_reset: call fw_set_waypoint_tasks first"); return FW_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    h->s64.noise = noise_dev; h->s64.noise_len = noise_len;
    h->s32.noise = noise_dev; h->s32.noise_len = noise_len;
    h->p64.S2.noise = noise_dev; h->p64.S2.noise_len = noise_len;
    h->p32.S2.noise = noise_dev; h->p32.S2.noise_len = noise_len;
    spare_join(h, st);        // the reset rewrites the precomputed rows of the envs it touches
    const int bs = 128, grid = (h->n + bs - 1) / bs;
    if (h->cfg.precision == FW_F64) reset_kernel<double><<<grid, bs, 0, st>>>(h->c64, h->s64, h->p64, mask_dev, state_dev, target_dev, obs_dev, obs64_dev);
    else reset_kernel<float><<<grid, bs, 0, st>>>(h->c32, h->s32, h->p32, mask_dev, state_dev, target_dev, obs_dev, obs64_dev);
    CK(cudaGetLastError());
    return FW_OK;
}

int fw_step(FwHandle* h, const void* actions_dev, int32_t actions_f64, float* obs_dev, float* rew_dev,
            uint8_t* done_dev, float* term_obs_dev, double* obs64_dev, double* rew64_dev, int32_t auto_reset,
            void* stream) {
    if (!h || !actions_dev) { snprintf(g_err, sizeof(g_err), "fw_step: null handle/actions"); return FW_EINVAL; }
    StepIO io;
    memset(&io, 0, sizeof(io));
    io.actions = actions_dev; io.actions_f64 = actions_f64; io.obs = obs_dev; io.rew = rew_dev; io.done = done_dev;
    io.term_obs = term_obs_dev; io.obs64 = obs64_dev; io.rew64 = rew64_dev; io.auto_reset = auto_reset;
    return launch_step(h, io, (cudaStream_t)stream);
}

int fw_step_random(FwHandle* h, int32_t k_steps, uint64_t action_seed, float* obs_dev, float* rew_dev,
                   uint8_t* done_dev, void* stream) {
    if (!h || k_steps <= 0) return FW_EINVAL;
    StepIO io;
    memset(&io, 0, sizeof(io));
    io.obs = obs_dev; io.rew = rew_dev; io.done = done_dev; io.auto_reset = 1; io.random_actions = 1;
    io.action_seed = action_seed;
    for (int k = 0; k < k_steps; ++k) {
        io.action_step = h->random_step++;
        int rc = launch_step(h, io, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return FW_OK;
}

int fw_get_episode_info_angular(FwHandle* h, double* metrics_ang_dev, void* stream) {
    if (!h || !metrics_ang_dev) return FW_EINVAL;
    if (!h->metrics_a) { snprintf(g_err, sizeof(g_err), "fw_get_episode_info_angular: the config has no attitude_angular targets"); return FW_EINVAL; }
    CK(cudaMemcpyAsync(metrics_ang_dev, h->metrics_a, sizeof(double) * FW_NMETRIC_ANG * h->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FW_OK;
}

int fw_get_episode_info(FwHandle* h, int32_t* term_code_dev, double* metrics_dev, double* ep_return_dev,
                        int32_t* ep_length_dev, void* stream) {
    if (!h) return FW_EINVAL;
    const int bs = 128, grid = (h->n + bs - 1) / bs;
    episode_info_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(h->n, h->ep_term, h->metrics, h->ep_ret, h->ep_len,
                                                              term_code_dev, metrics_dev, ep_return_dev, ep_length_dev);
    CK(cudaGetLastError());
    return FW_OK;
}

static int field_copy(FwHandle* h, int32_t field, void* buf, int to_soa, void* stream) {
    if (!h || !buf || field < 0 || field >= FW_FIELD_COUNT) return FW_EINVAL;
    if (field == FW_FIELD_PARAMS && !h->par_buf) {
        snprintf(g_err, sizeof(g_err), "FW_FIELD_PARAMS: the handle was created without simulator.model (model_on = 0)");
        return FW_EINVAL;
    }
    const int bs = 128, grid = (h->n + bs - 1) / bs;
    if (h->cfg.precision == FW_F64) field_copy_kernel<double><<<grid, bs, 0, (cudaStream_t)stream>>>(h->c64, h->s64, field, buf, to_soa);
    else field_copy_kernel<float><<<grid, bs, 0, (cudaStream_t)stream>>>(h->c32, h->s32, field, buf, to_soa);
    CK(cudaGetLastError());
    return FW_OK;
}
int fw_get_field(FwHandle* h, int32_t field, void* out_dev, void* stream) { return field_copy(h, field, out_dev, 0, stream); }
int fw_set_field(FwHandle* h, int32_t field, const void* in_dev, void* stream) {
    if (field == FW_FIELD_NFEV || field == FW_FIELD_TURB) return FW_EINVAL;
    return field_copy(h, field, (void*)in_dev, 1, stream);
}

}  // extern "C"
