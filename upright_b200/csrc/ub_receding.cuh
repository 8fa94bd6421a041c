// upright_b200 — receding-horizon bookkeeping on the device for B instances at once:
//   * desired end-effector position at every knot of the new horizon
//     (interpolate_end_effector_pose, upright_control/include/upright_control/reference_trajectory.h:18-47),
//   * warm start = previous primal solution interpolated at the new time grid, tail held
//     (ocs2 MPC warm start as used through manager.py:156-170 [EXT]; mpc.cold_start = false, controller.yaml),
//   * policy evaluation u = u_ff(t) + K(t) (x - x_nom(t)) with linear interpolation between knots
//     (evaluateMpcSolution, upright_control/src/pybindings.cpp:378-381; docs/configuration.md:105-108),
//   * the low-level tracking law and plant of the simulation loop
//     (upright_cmd/scripts/simulations/mpc_sim.py:101-108,148-155) with the model's own exact
//     triple-integrator step (dynamics/system_dynamics.h:15-23) as the plant.
// The host-side numpy mirror of the same rules is upright_b200/manager.py::_RecedingHorizon.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

// target [B, N+1, 3] at knot times t + k dt from waypoints (times [M], pos [B, M, 3])
template <typename T>
__global__ void rh_targets_kernel(int B, int N, double dt, double t, const double* __restrict__ times, int M,
                                  const T* __restrict__ pos, T* __restrict__ target) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * (N + 1)) return;
    const int b = idx / (N + 1), k = idx % (N + 1);
    const double tau = t + dt * k;
    const T* p = pos + size_t(b) * M * 3;
    T* o = target + size_t(idx) * 3;
    if (M == 1 || tau <= times[0]) {
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        return;
    }
    if (tau >= times[M - 1]) {
        const T* q = p + 3 * (M - 1);
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
        return;
    }
    int i = 0;
    while (i + 2 < M && times[i + 1] <= tau) ++i;   // times[i] <= tau < times[i+1]
    const double a = (times[i + 1] - tau) / (times[i + 1] - times[i]);
    for (int c = 0; c < 3; ++c) o[c] = T(a * double(p[3 * i + c]) + (1.0 - a) * double(p[3 * (i + 1) + c]));
}

// previous solution (grid starting at t_old) -> new grid starting at t_new; x_0 is overwritten by the
// observation inside the solve kernel
template <typename T>
__global__ void rh_shift_kernel(int B, int N, int nx, int nu, double dt, double t_old, double t_new,
                                const T* __restrict__ Xo, const T* __restrict__ Uo, T* __restrict__ Xn, T* __restrict__ Un) {
    const int per = (N + 1) * nx + N * nu;
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(B) * per) return;
    const int b = int(idx / per), e = int(idx % per);
    const bool is_x = e < (N + 1) * nx;
    const int k = is_x ? e / nx : (e - (N + 1) * nx) / nu;
    const int c = is_x ? e % nx : (e - (N + 1) * nx) % nu;
    const double tk = t_new + dt * k;
    // interval index on the old grid, clamped to [0, N-1]; weight clamped to [0, 1] (tail held)
    double s = (tk - t_old) / dt;
    int i = int(floor(s + 1e-9));
    i = i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
    double w = s - i;
    w = w < 0.0 ? 0.0 : (w > 1.0 ? 1.0 : w);
    if (is_x) {
        const T* X = Xo + size_t(b) * (N + 1) * nx;
        Xn[size_t(b) * (N + 1) * nx + k * nx + c] = T((1.0 - w) * double(X[i * nx + c]) + w * double(X[(i + 1) * nx + c]));
    } else {
        const T* U = Uo + size_t(b) * N * nu;
        const int j = (i + 1 > N - 1) ? N - 1 : i + 1;   // U extended by its last value
        Un[size_t(b) * N * nu + k * nu + c] = T((1.0 - w) * double(U[i * nu + c]) + w * double(U[j * nu + c]));
    }
}

// per-instance histogram of the solve status over the replans of a rollout
__global__ void rh_count_status_kernel(int B, const int32_t* __restrict__ status, int32_t* __restrict__ counts) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int s = status[b];
    if (s >= 0 && s < 4) counts[4 * b + s] += 1;
}

// Simulated dynamic obstacle of the closed loop (upright_sim/src/upright_sim/simulation.py:300-435 as restated by
// upright_b200/plant.py::BallisticObstacles): free flight under the current mode's acceleration, state reset to the
// next mode's initial values once its time has come (checked on the time at the START of a simulation step).
struct ObstacleMode {
    double time, p[3], v[3], a[3];
};

template <typename T>
struct RolloutArgs {
    int B, N, nq, nx, nu;
    int nxt;              // columns of x / X: robot state + 9 per dynamic obstacle
    int ndyn;             // dynamic obstacles
    double dt;            // knot spacing of the plan
    double t0;            // start time of the current plan
    double t_first;       // time of the first simulation step of this launch
    double sim_dt;
    int n_sub;            // simulation steps in this launch
    int step0;            // global index of the first step (for logging)
    int log_stride, n_log;
    int use_feedback, gain_stages;
    T kp, kv, ka;
    const T* X;           // [B, N+1, nxt] current plan
    const T* U;           // [B, N, nu]
    const T* K;           // [B, gain_stages, nu, nx] or null
    T* x;                 // [B, nxt] plant state (in/out)
    T* xs;                // [B, n_log, nxt] or null
    T* us;                // [B, n_log, nq] or null
    // obstacle plant: modes [ndyn][max_modes], n_modes [ndyn], offsets [B, ndyn, 3] or null, current mode [B, ndyn]
    const ObstacleMode* modes;
    const int* n_modes;
    int max_modes;
    const T* offsets;
    int* mode_idx;
};

// One warp per instance, lane r < nq owns joint r (q_r, v_r, a_r); lane nq + 3 j + c owns axis c of obstacle j.
// n_sub simulation steps per launch.
template <typename T>
__global__ void rh_rollout_kernel(RolloutArgs<T> A) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (warp >= A.B) return;
    const int b = warp, nq = A.nq, nx = A.nx, nxt = A.nxt, nu = A.nu, N = A.N;
    const bool own = lane < nq;
    const int ol = lane - nq, oj = ol / 3, oc = ol % 3;          // obstacle lane: obstacle oj, axis oc
    const bool obs = ol >= 0 && oj < A.ndyn;
    T* xg = A.x + size_t(b) * nxt;
    T q = own ? xg[lane] : T(0), v = own ? xg[nq + lane] : T(0), a = own ? xg[2 * nq + lane] : T(0);
    T orr = obs ? xg[nx + 9 * oj + oc] : T(0), ov = obs ? xg[nx + 9 * oj + 3 + oc] : T(0), oa = obs ? xg[nx + 9 * oj + 6 + oc] : T(0);
    int midx = obs ? A.mode_idx[b * A.ndyn + oj] : 0;
    const T* X = A.X + size_t(b) * (N + 1) * nxt;
    const T* U = A.U + size_t(b) * N * nu;
    const T* K = A.K ? A.K + size_t(b) * A.gain_stages * nu * nx : nullptr;
    const unsigned FULLM = 0xffffffffu;
    for (int sub = 0; sub < A.n_sub; ++sub) {
        const double t = A.t_first + A.sim_dt * sub;
        const double s = (t - A.t0) / A.dt;
        int i = int(floor(s));
        i = i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
        double wd = s - i;
        wd = wd < 0.0 ? 0.0 : (wd > 1.0 ? 1.0 : wd);
        const T w = T(wd), w1 = T(1.0 - wd);
        const int j = (i + 1 > N - 1) ? N - 1 : i + 1;
        T dq = 0, dv = 0, da = 0, u = 0;
        if (own) {
            const T* Xi = X + i * nxt;
            const T* Xj = Xi + nxt;
            dq = q - (w1 * Xi[lane] + w * Xj[lane]);
            dv = v - (w1 * Xi[nq + lane] + w * Xj[nq + lane]);
            da = a - (w1 * Xi[2 * nq + lane] + w * Xj[2 * nq + lane]);
            u = w1 * U[i * nu + lane] + w * U[j * nu + lane];
        }
        if (A.use_feedback && K != nullptr) {
            // gains exist for stages < gain_stages; beyond that the feed-forward input is used alone
            const bool have = j < A.gain_stages;
            const T* Ki = K + (size_t(have ? i : 0) * nu + (own ? lane : 0)) * nx;
            const T* Kj = K + (size_t(have ? j : 0) * nu + (own ? lane : 0)) * nx;
            T fb = 0;
            for (int r = 0; r < nq; ++r) {
                const T eq = __shfl_sync(FULLM, dq, r), ev = __shfl_sync(FULLM, dv, r), ea = __shfl_sync(FULLM, da, r);
                if (own && have) {
                    fb += (w1 * Ki[r] + w * Kj[r]) * eq + (w1 * Ki[nq + r] + w * Kj[nq + r]) * ev +
                          (w1 * Ki[2 * nq + r] + w * Kj[2 * nq + r]) * ea;
                }
            }
            u += fb;
        }
        // u_cmd = Kx (xd - x) + u   (mpc_sim.py:148), xd = x_nom
        const T ucmd = u - (A.kp * dq + A.kv * dv + A.ka * da);
        const int step = A.step0 + sub;
        if (A.xs != nullptr && step % A.log_stride == 0) {
            const int l = step / A.log_stride;
            if (l < A.n_log) {
                T* xo = A.xs + (size_t(b) * A.n_log + l) * nxt;
                if (own) {
                    xo[lane] = q;
                    xo[nq + lane] = v;
                    xo[2 * nq + lane] = a;
                    A.us[(size_t(b) * A.n_log + l) * nq + lane] = ucmd;
                }
                if (obs) {
                    xo[nx + 9 * oj + oc] = orr;
                    xo[nx + 9 * oj + 3 + oc] = ov;
                    xo[nx + 9 * oj + 6 + oc] = oa;
                }
            }
        }
        // exact triple-integrator step over sim_dt with constant jerk
        const T h = T(A.sim_dt);
        const T qn = q + h * v + T(0.5) * h * h * a + h * h * h / T(6) * ucmd;
        const T vn = v + h * a + T(0.5) * h * h * ucmd;
        const T an = a + h * ucmd;
        q = qn; v = vn; a = an;
        if (obs) {
            // mode schedule, then free flight over the step (plant.py::BallisticObstacles.step)
            if (A.modes != nullptr) {
                const ObstacleMode* M = A.modes + oj * A.max_modes;
                if (midx < A.n_modes[oj] - 1 && t >= M[midx + 1].time) {
                    ++midx;
                    orr = T(M[midx].p[oc]) + (A.offsets ? A.offsets[(size_t(b) * A.ndyn + oj) * 3 + oc] : T(0));
                    ov = T(M[midx].v[oc]);
                }
                oa = T(M[midx].a[oc]);
            }
            orr = orr + h * ov + T(0.5) * h * h * oa;
            ov = ov + h * oa;
        }
    }
    if (own) {
        xg[lane] = q;
        xg[nq + lane] = v;
        xg[2 * nq + lane] = a;
    }
    if (obs) {
        xg[nx + 9 * oj + oc] = orr;
        xg[nx + 9 * oj + 3 + oc] = ov;
        xg[nx + 9 * oj + 6 + oc] = oa;
        if (oc == 0) A.mode_idx[b * A.ndyn + oj] = midx;
    }
}

}  // namespace ub
