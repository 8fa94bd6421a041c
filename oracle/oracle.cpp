// TEST INFRASTRUCTURE — CPU oracle (fp64) for the upright MPC hot path.
//
// This is NOT the product.  It restates, on the CPU and in double precision,
// the solve that `ControllerInterface::advanceMpc()` performs in the reference
// (call stack: upright_control/src/upright_control/manager.py:156-170 ->
// upright_control/src/pybindings.cpp:376 -> ocs2::MultipleShootingMpc built at
// upright_control/src/controller_interface.cpp:395-398).  The problem
// functions are in model.h; this file holds the multiple-shooting SQP step:
// Gauss-Newton linearisation, the OCP-QP, the filter line search.
//
// PARITY UNPINNED at the solver boundary: the reference's SQP/QP code lives in
// the un-vendored fork utiasDSL/ocs2@upright (ocs2_sqp + hpipm_catkin, no
// version pinned in upright_control/package.xml) and the repository holds no
// test, golden vector or logged trajectory for it.  The
// oracle therefore documents its choices and is pinned only by
// (i) the reference's set-up golden vectors (tests/test_parsing_golden.py),
// (ii) finite-difference and analytic checks of every model function and
// (iii) an independent dense numpy solve of the same QP (tests/test_oracle_qp.py).
//
// QP method: the OCP-QP with L2-softened rows (HPIPM slack semantics with
// Zl=Zu=Z, zl=zu=0, slack lower bound 0 — wrappers.py:121-143) is the
// piecewise-quadratic programme   min 1/2 z'Hz + g'z + sum_i rho_i/2 dist^2(a_i'z + c_i, [l_i,u_i])
// subject to the linear dynamics; hard rows use the same term inside an
// augmented-Lagrangian loop.  It is solved by a semismooth Newton method whose
// linear systems are Riccati recursions (dense A, B here) with an exact line
// search.  The minimiser is unique (strictly convex), so any converged QP
// method — HPIPM's interior point included — returns the same step.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <thread>
#include <vector>

#include "model.h"

namespace orc {

static const double INF = std::numeric_limits<double>::infinity();

struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d(size_t(r_) * c_, 0.0) {}
    double& operator()(int i, int j) { return d[size_t(i) * c + j]; }
    double operator()(int i, int j) const { return d[size_t(i) * c + j]; }
    void zero() { std::fill(d.begin(), d.end(), 0.0); }
};

struct Dims {
    int nq, nx, nxr, nu, nfc, neq, nfric, nobs, npairs, nproj, nterm, N, nz;   // nx = nxr (robot) + 9 per dynamic obstacle
};

static Dims make_dims(const ub_problem_desc_t& P) {
    Dims D;
    D.nq = P.nq;
    D.nxr = 3 * P.nq;
    // Dynamic obstacles are GENUINE states of the oracle's QP (as in the reference: dimensions.h:36-41,
    // system_dynamics.h:28-38,86-104): uncontrolled constant-acceleration blocks appended to A with zero rows in B.
    // The CUDA kernels eliminate them analytically; agreement of the two formulations is part of the parity tests.
    D.nx = D.nxr + (P.obstacles_enabled ? 9 * P.n_dynamic_obstacles : 0);
    const bool bal = P.balancing_enabled && P.nb > 0;
    D.nfc = bal ? P.nf * P.nc : 0;
    D.nu = P.nq + D.nfc;
    D.neq = bal ? 6 * P.nb : 0;
    D.nfric = (bal && P.nf == 3) ? 5 * P.nc : 0;
    D.npairs = P.obstacles_enabled ? P.n_pairs : 0;
    // end-effector box rows, then inertial-alignment rows, follow the sphere-pair rows
    // projectile-path rows (one per listed link) close the family; they need the obstacle state x.tail(9)
    D.nproj = (P.projectile_enabled && D.nx > D.nxr) ? P.n_projectile_links : 0;
    D.nobs = D.npairs + (P.ee_box_enabled ? 6 : 0) + (P.ia_constraint_enabled ? 5 : 0) + D.nproj;
    D.nterm = 3 + 2 * P.nq;  // stationary_desired_position_constraint.h:39-41
    D.N = P.N;
    D.nz = D.nu + D.nx;
    return D;
}

// Exact discretisation of the triple integrator (system_dynamics.h:15-26).
// A is nilpotent-plus-identity, so the RK4 sensitivity integrator the
// reference forces (controller_interface.cpp:115) reproduces exp(A dt) exactly.
static void discrete_dynamics(const ub_problem_desc_t& P, const Dims& D, Mat& A, Mat& B) {
    const int nq = D.nq;
    const double dt = P.dt;
    A = Mat(D.nx, D.nx);
    B = Mat(D.nx, D.nu);
    for (int i = 0; i < nq; ++i) {
        A(i, i) = A(nq + i, nq + i) = A(2 * nq + i, 2 * nq + i) = 1.0;
        A(i, nq + i) = dt;
        A(i, 2 * nq + i) = 0.5 * dt * dt;
        A(nq + i, 2 * nq + i) = dt;
        B(i, i) = dt * dt * dt / 6.0;
        B(nq + i, i) = 0.5 * dt * dt;
        B(2 * nq + i, i) = dt;
    }
    // ObstacleDynamics::flowmap (system_dynamics.h:28-38): p' = v, v' = a, a' = 0, exact over dt
    for (int o = D.nxr; o < D.nx; o += 9)
        for (int c = 0; c < 3; ++c) {
            A(o + c, o + c) = A(o + 3 + c, o + 3 + c) = A(o + 6 + c, o + 6 + c) = 1.0;
            A(o + c, o + 3 + c) = dt;
            A(o + c, o + 6 + c) = 0.5 * dt * dt;
            A(o + 3 + c, o + 6 + c) = dt;
        }
}

// ---------------------------------------------------------------------------
// Linearisation of one knot (values + Jacobians via Dual)
struct KnotLin {
    std::vector<double> g;      // object-dynamics equality, neq
    Mat C;                      // d g / d x  (neq x nx)
    Mat Df;                     // d g / d f  (neq x nfc) — constant in x
    double r[3];                // EE position
    Mat Jp;                     // d r / d q (3 x nq)
    std::vector<double> hfric;  // friction rows value, nfric
    Mat Ffric;                  // d hfric / d f (nfric x nfc)
    std::vector<double> hobs;   // obstacle rows value, nobs
    Mat Jobs;                   // d hobs / d q (nobs x nq)
    double eo[3] = {0, 0, 0};   // end-effector orientation error (quaternion distance to the target)
    Mat Jo;                     // d eo / d q (3 x nq)
    double ea[2] = {0, 0};      // inertial-alignment residual
    Mat Jea;                    // d ea / d x (2 x nx)
    double hia[5] = {0, 0, 0, 0, 0};  // inertial-alignment constraint rows
    Mat Jia;                    // d hia / d x (5 x nx)
    std::vector<double> hproj;  // projectile-path rows, nproj
    Mat Jproj;                  // d hproj / d x (nproj x nx): robot q block and the last obstacle's nine states
};

static inline bool orientation_weighted(const ub_problem_desc_t& P) {
    return P.ee_weight[3] != 0.0 || P.ee_weight[4] != 0.0 || P.ee_weight[5] != 0.0;
}
// columns of a target row: desired position, and the desired quaternion [x y z w] behind it when the orientation
// part of the end-effector weight is non-zero (wrappers.py:36-42: target states are [r, quat, s])
static inline int target_stride(const ub_problem_desc_t& P) { return orientation_weighted(P) ? 7 : 3; }

static void linearize_knot(const ub_problem_desc_t& P, const Dims& D, const double* body_params, const double* x,
                           const double* u, KnotLin& L, const double* qref = nullptr) {
    const int nq = D.nq, nx = D.nx, nxr = D.nxr;
    std::vector<Dual> xd(nx);
    for (int i = 0; i < nx; ++i) xd[i] = i < nxr ? Dual::variable(x[i], i) : Dual(x[i]);  // AD over the robot state
    const Kinematics<Dual> K = forward_kinematics<Dual>(P, xd.data());
    L.Jp = Mat(3, nq);
    for (int i = 0; i < 3; ++i) {
        L.r[i] = K.r[i].v;
        for (int j = 0; j < nq; ++j) L.Jp(i, j) = K.r[i].d[j];
    }
    L.Jo = Mat(3, nq);
    if (qref != nullptr) {   // end_effector_cost.h:61-67
        const Vec3<Dual> eo = orientation_error<Dual>(K.C_we, qref);
        for (int i = 0; i < 3; ++i) {
            L.eo[i] = eo[i].v;
            for (int j = 0; j < nq; ++j) L.Jo(i, j) = eo[i].d[j];
        }
    }
    L.g.assign(D.neq, 0.0);
    L.C = Mat(D.neq, nx);
    L.Df = Mat(D.neq, D.nfc);
    if (D.neq > 0) {
        std::vector<Dual> f(D.nfc), g(D.neq);
        for (int i = 0; i < D.nfc; ++i) f[i] = Dual(u[nq + i]);
        object_dynamics_constraints<Dual>(P, body_params, K, f.data(), g.data());
        for (int i = 0; i < D.neq; ++i) {
            L.g[i] = g[i].v;
            for (int j = 0; j < nxr; ++j) L.C(i, j) = g[i].d[j];
        }
        // g is affine in the forces: column j of Df = g(f + e_j) - g(f)
        const Kinematics<double> Kd = forward_kinematics<double>(P, x);
        std::vector<double> f0(D.nfc, 0.0), g0(D.neq), g1(D.neq);
        object_dynamics_constraints<double>(P, body_params, Kd, f0.data(), g0.data());
        for (int j = 0; j < D.nfc; ++j) {
            f0[j] = 1.0;
            object_dynamics_constraints<double>(P, body_params, Kd, f0.data(), g1.data());
            f0[j] = 0.0;
            for (int i = 0; i < D.neq; ++i) L.Df(i, j) = g1[i] - g0[i];
        }
    }
    L.hfric.assign(D.nfric, 0.0);
    L.Ffric = Mat(D.nfric, D.nfc);
    if (D.nfric > 0) {
        std::vector<double> f(D.nfc), h0(D.nfric), h1(D.nfric);
        for (int i = 0; i < D.nfc; ++i) f[i] = u[nq + i];
        contact_force_constraints<double>(P, f.data(), L.hfric.data());
        std::fill(f.begin(), f.end(), 0.0);
        contact_force_constraints<double>(P, f.data(), h0.data());
        for (int j = 0; j < D.nfc; ++j) {
            f[j] = 1.0;
            contact_force_constraints<double>(P, f.data(), h1.data());
            f[j] = 0.0;
            for (int i = 0; i < D.nfric; ++i) L.Ffric(i, j) = h1[i] - h0[i];
        }
    }
    L.Jea = Mat(2, nx);
    if (P.ia_cost_enabled) {
        Dual e[2];
        inertial_alignment_error<Dual>(P, K, e);
        for (int r = 0; r < 2; ++r) {
            L.ea[r] = e[r].v;
            for (int j = 0; j < nxr; ++j) L.Jea(r, j) = e[r].d[j];
        }
    }
    L.Jia = Mat(5, nx);
    if (P.ia_constraint_enabled) {
        Dual h[5];
        inertial_alignment_constraints<Dual>(P, K, h);
        for (int r = 0; r < 5; ++r) {
            L.hia[r] = h[r].v;
            for (int j = 0; j < nxr; ++j) L.Jia(r, j) = h[r].d[j];
        }
    }
    L.hproj.assign(D.nproj, 0.0);
    L.Jproj = Mat(D.nproj, nx);
    if (D.nproj > 0) {   // projectile_path_constraint.h:108-146
        std::vector<Dual> h(D.nproj);
        std::vector<double> tc(D.nproj);
        const double* xo = x + nx - 9;
        projectile_constraints<Dual>(P, K, xo, h.data(), tc.data());
        for (int i = 0; i < D.nproj; ++i) {
            L.hproj[i] = h[i].v;
            for (int j = 0; j < nq; ++j) L.Jproj(i, j) = h[i].d[j];
            // - w s n' [I, t I, t^2/2 I] over the obstacle state, n = delta / |delta|
            const double t = tc[i], w = P.projectile_scale / P.projectile_distances[i] * P.projectile_active;
            double n[3], len = 0;
            for (int c = 0; c < 3; ++c) {
                n[c] = K.sphere[P.projectile_spheres[i]][c].v - (xo[c] + t * xo[3 + c] + 0.5 * t * t * xo[6 + c]);
                len += n[c] * n[c];
            }
            len = std::sqrt(len);
            for (int c = 0; c < 3; ++c) {
                L.Jproj(i, nx - 9 + c) = -w * n[c] / len;
                L.Jproj(i, nx - 6 + c) = -w * t * n[c] / len;
                L.Jproj(i, nx - 3 + c) = -w * 0.5 * t * t * n[c] / len;
            }
        }
    }
    L.hobs.assign(D.npairs, 0.0);
    L.Jobs = Mat(D.npairs, nx);   // dense over q, plus the position block of a dynamic obstacle in the pair
    if (D.npairs > 0) {
        std::vector<Dual> h(D.npairs);
        obstacle_constraints<Dual>(P, K, h.data());
        for (int i = 0; i < D.npairs; ++i) {
            L.hobs[i] = h[i].v;
            for (int j = 0; j < nq; ++j) L.Jobs(i, j) = h[i].d[j];
            // d |c_a - c_b| / d c_a = +n, / d c_b = -n for an obstacle-borne centre (analytic: AD runs over x_robot)
            const int pa = P.pairs[i].a, pb = P.pairs[i].b;
            double n[3], len = 0;
            for (int c = 0; c < 3; ++c) {
                n[c] = K.sphere[pa][c].v - K.sphere[pb][c].v;
                len += n[c] * n[c];
            }
            len = std::sqrt(len);
            if (P.spheres[pa].shape == UB_SHAPE_HALFSPACE || P.spheres[pb].shape == UB_SHAPE_HALFSPACE) {
                // sphere against the half-space: d h / d c = +n (the half-space itself never rides on an obstacle)
                const int hs = P.spheres[pb].shape == UB_SHAPE_HALFSPACE ? pb : pa;
                for (int c = 0; c < 3; ++c) n[c] = (hs == pb ? 1.0 : -1.0) * P.spheres[hs].offset[c];
                len = 1.0;
            }
            for (int side = 0; side < 2; ++side) {
                const int s = side == 0 ? pa : pb;
                if (P.spheres[s].link > -2) continue;
                const int o = nxr + 9 * (-2 - P.spheres[s].link);
                for (int c = 0; c < 3; ++c) L.Jobs(i, o + c) += (side == 0 ? 1.0 : -1.0) * n[c] / len;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// QP description.  Stage variable z_k = [du_k (nu); dx_k (nx)] (terminal: dx only).
struct Row {
    int idx = -1;           // >= 0: unit row on z[idx]; otherwise dense `a`
    std::vector<double> a;  // dense coefficients (length nz of the stage)
    double c = 0.0;         // row value at z = 0
    double lb = -INF, ub = INF;
    double rho = 0.0;
    bool hard = false;
    double lambda = 0.0;
    double rho0 = 0.0;
    double t[2] = {0, 0}, lam[2] = {0, 0};  // IPM: slack / multiplier of the lower / upper side
};

struct Stage {
    int nz = 0, nu = 0;
    Mat H;                  // cost Hessian (nz x nz)
    std::vector<double> g;  // cost gradient
    std::vector<double> b;  // dynamics gap A x_k + B u_k - x_{k+1} (k < N)
    std::vector<Row> rows;
};

struct Perf {
    double cost = 0, dyn_sse = 0, eq_sse = 0, ineq_sse = 0;
    double max_eq = 0, min_margin = INF;
    double violation() const { return std::sqrt(dyn_sse + eq_sse + ineq_sse); }
};

struct Workspace {
    Dims D;
    Mat A, B;
    std::vector<Stage> st;
    std::vector<KnotLin> lin;
};

static double sq(double v) { return v * v; }

// Performance index of a trajectory: intermediate cost scaled by dt (OCS2
// multiple-shooting transcription [EXT]); cost terms at
// cost/quadratic_joint_state_input_cost.h:9-33 (weights controller_interface.cpp:400-420)
// and cost/end_effector_cost.h:31-46; terminal equality
// constraint/stationary_desired_position_constraint.h:43-56.
static Perf performance(const ub_problem_desc_t& P, const Dims& D, const Mat& A, const Mat& B,
                        const double* body_params, const double* target, const double* X, const double* U) {
    Perf pf;
    const int nq = D.nq, nx = D.nx, nu = D.nu, N = D.N;
    const double dt = P.dt;
    std::vector<double> g(std::max(D.neq, 1)), h(std::max(std::max(D.nfric, D.nobs), 1)), f(std::max(D.nfc, 1));
    for (int k = 0; k <= N; ++k) {
        const double* x = X + size_t(k) * nx;
        const Kinematics<double> K = forward_kinematics<double>(P, x);
        const double* rd = target + target_stride(P) * k;
        if (k == N) {
            for (int i = 0; i < 3; ++i) {
                const double e = rd[i] - K.r[i];
                pf.eq_sse += e * e;
                pf.max_eq = std::max(pf.max_eq, std::fabs(e));
            }
            for (int i = nq; i < D.nxr; ++i) {
                pf.eq_sse += sq(x[i]);
                pf.max_eq = std::max(pf.max_eq, std::fabs(x[i]));
            }
        }
        if (k >= 1)
            for (int i = 0; i < D.nxr; ++i) {
                const double lo = x[i] - P.state_lb[i], hi = P.state_ub[i] - x[i];
                pf.ineq_sse += dt * (sq(std::min(0.0, lo)) + sq(std::min(0.0, hi)));
                pf.min_margin = std::min(pf.min_margin, std::min(lo, hi));
            }
        if (k == N) break;
        const double* u = U + size_t(k) * nu;
        double c = 0;
        for (int i = 0; i < D.nxr; ++i) c += 0.5 * P.state_weight[i] * sq(x[i] - P.xd[i]);  // no weight on obstacle states
        for (int i = 0; i < nq; ++i) c += 0.5 * P.input_weight[i] * sq(u[i]);
        for (int i = 0; i < D.nfc; ++i) c += 0.5 * P.force_weight * sq(u[nq + i]);
        for (int i = 0; i < 3; ++i) c += 0.5 * P.ee_weight[i] * sq(K.r[i] - rd[i]);
        if (orientation_weighted(P)) {
            const Vec3<double> eo = orientation_error<double>(K.C_we, rd + 3);
            for (int i = 0; i < 3; ++i) c += 0.5 * P.ee_weight[3 + i] * sq(eo[i]);
        }
        if (P.ia_cost_enabled) {   // inertial_alignment.cpp:118-124
            double e[2];
            inertial_alignment_error<double>(P, K, e);
            c += 0.5 * P.ia_cost_weight * (sq(e[0]) + sq(e[1]));
        }
        pf.cost += dt * c;
        const double* xn = X + size_t(k + 1) * nx;
        for (int i = 0; i < nx; ++i) {
            double gap = -xn[i];
            for (int j = 0; j < nx; ++j) gap += A(i, j) * x[j];
            for (int j = 0; j < nq; ++j) gap += B(i, j) * u[j];
            pf.dyn_sse += dt * gap * gap;
        }
        for (int i = 0; i < nu; ++i) {
            const double lbv = i < nq ? P.input_lb[i] : P.force_lb, ubv = i < nq ? P.input_ub[i] : P.force_ub;
            const double lo = u[i] - lbv, hi = ubv - u[i];
            pf.ineq_sse += dt * (sq(std::min(0.0, lo)) + sq(std::min(0.0, hi)));
            pf.min_margin = std::min(pf.min_margin, std::min(lo, hi));
        }
        for (int i = 0; i < D.nfc; ++i) f[i] = u[nq + i];
        if (D.neq > 0) {
            object_dynamics_constraints<double>(P, body_params, K, f.data(), g.data());
            for (int i = 0; i < D.neq; ++i) {
                pf.eq_sse += dt * sq(g[i]);
                pf.max_eq = std::max(pf.max_eq, std::fabs(g[i]));
            }
        }
        if (D.nfric > 0) {
            contact_force_constraints<double>(P, f.data(), h.data());
            for (int i = 0; i < D.nfric; ++i) {
                pf.ineq_sse += dt * sq(std::min(0.0, h[i]));
                pf.min_margin = std::min(pf.min_margin, h[i]);
            }
        }
        if (D.npairs > 0 && k >= 1) {
            obstacle_constraints<double>(P, K, h.data());
            for (int i = 0; i < D.npairs; ++i) {
                pf.ineq_sse += dt * sq(std::min(0.0, h[i]));
                pf.min_margin = std::min(pf.min_margin, h[i]);
            }
        }
        if (P.ia_constraint_enabled && k >= 1) {   // inertial_alignment.cpp:7-53
            double h5[5];
            inertial_alignment_constraints<double>(P, K, h5);
            for (int r = 0; r < 5; ++r) {
                pf.ineq_sse += dt * sq(std::min(0.0, h5[r]));
                pf.min_margin = std::min(pf.min_margin, h5[r]);
            }
        }
        if (P.ee_box_enabled && k >= 1)   // constraint/end_effector_box_constraint.h:46-58
            for (int c = 0; c < 3; ++c) {
                const double hu = rd[c] + P.ee_box_upper[c] - K.r[c], hl = K.r[c] - rd[c] - P.ee_box_lower[c];
                pf.ineq_sse += dt * (sq(std::min(0.0, hu)) + sq(std::min(0.0, hl)));
                pf.min_margin = std::min(pf.min_margin, std::min(hu, hl));
            }
        if (D.nproj > 0 && k >= 1) {   // projectile_path_constraint.h:77-106
            double hp[UB_MAX_PROJECTILE_LINKS];
            projectile_constraints<double>(P, K, x + nx - 9, hp, nullptr);
            for (int i = 0; i < D.nproj; ++i) {
                pf.ineq_sse += dt * sq(std::min(0.0, hp[i]));
                pf.min_margin = std::min(pf.min_margin, hp[i]);
            }
        }
    }
    return pf;
}

// Build the QP of one SQP iteration around (X, U).
static void build_qp(const ub_problem_desc_t& P, Workspace& W, const double* body_params, const double* target,
                     const double* X, const double* U) {
    const Dims& D = W.D;
    const int nq = D.nq, nx = D.nx, nu = D.nu, N = D.N;
    const double dt = P.dt;
    const ub_slack_settings_t& S = P.slacks;
    const double Z = S.upper_L2_penalty;
    const bool soft_poly = S.enabled && S.poly_ineq, soft_x = S.enabled && S.state_box,
               soft_u = S.enabled && S.input_box;
    auto finish = [&](Row& r, bool soft) {
        r.hard = !soft;
        // hard rows are unit-normalised for the proximal/AL term (scaling a hard
        // row does not change the QP); soft rows carry the L2 slack weight as is
        double n2 = 1.0;
        if (r.idx < 0) {
            n2 = 0.0;
            for (double v : r.a) n2 += v * v;
        }
        r.rho = soft ? Z : (n2 > 0.0 ? P.rho_hard / n2 : 0.0);
        r.rho0 = r.rho;
        r.lambda = 0.0;
    };
    W.st.assign(N + 1, Stage());
    W.lin.resize(N + 1);
    std::vector<double> zero_u(nu, 0.0);
    for (int k = 0; k <= N; ++k) {
        const double* x = X + size_t(k) * nx;
        const double* u = (k < N) ? U + size_t(k) * nu : zero_u.data();
        KnotLin& L = W.lin[k];
        const int ts = target_stride(P);
        linearize_knot(P, D, body_params, x, u, L, ts == 7 ? target + ts * k + 3 : nullptr);
        if (const char* ne = std::getenv("ORACLE_LIN_NOISE")) {
            // precision study only (tools/precision_lab2.py): perturb every block of the linearisation by noise x its
            // largest magnitude — the size of the error an fp32 forward-kinematics pass leaves in it
            const double noise = std::atof(ne);
            static thread_local unsigned long long hs = 88172645463325252ull;
            auto jiggle = [&](double* v, size_t n) {
                double mx = 0;
                for (size_t i = 0; i < n; ++i) mx = std::max(mx, std::fabs(v[i]));
                for (size_t i = 0; i < n; ++i) {
                    if (v[i] == 0.0) continue;
                    hs ^= hs << 13; hs ^= hs >> 7; hs ^= hs << 17;
                    v[i] += 2.0 * noise * mx * (double(hs >> 11) / double(1ull << 53) - 0.5);
                }
            };
            jiggle(L.Jp.d.data(), L.Jp.d.size());
            jiggle(L.r, 3);
            jiggle(L.C.d.data(), L.C.d.size());
            jiggle(L.g.data(), L.g.size());
            jiggle(L.Jobs.d.data(), L.Jobs.d.size());
            jiggle(L.hobs.data(), L.hobs.size());
        }
        Stage& s = W.st[k];
        s.nu = (k < N) ? nu : 0;
        s.nz = s.nu + nx;
        const int xo = s.nu;  // offset of dx in z
        s.H = Mat(s.nz, s.nz);
        s.g.assign(s.nz, 0.0);
        const double* rd = target + ts * k;
        if (k < N) {
            // cost: dt * (1/2 (x-xd)'Q(x-xd) + 1/2 u'R u + 1/2 e'W e), Gauss-Newton in e
            // (end_effector_cost.h:48-84)
            for (int i = 0; i < nq; ++i) {
                s.H(i, i) = dt * P.input_weight[i] + P.reg_input;
                s.g[i] = dt * P.input_weight[i] * u[i];
            }
            for (int i = 0; i < D.nfc; ++i) {
                s.H(nq + i, nq + i) = dt * P.force_weight + P.reg_input;
                s.g[nq + i] = dt * P.force_weight * u[nq + i];
            }
            for (int i = 0; i < D.nxr; ++i) {   // obstacle states carry no weight (controller_interface.cpp:410-414)
                s.H(xo + i, xo + i) = dt * P.state_weight[i];
                s.g[xo + i] = dt * P.state_weight[i] * (x[i] - P.xd[i]);
            }
            for (int a = 0; a < nq; ++a)
                for (int c = 0; c < 3; ++c) {
                    const double w = dt * P.ee_weight[c] * L.Jp(c, a);
                    s.g[xo + a] += w * (L.r[c] - rd[c]);
                    for (int b = 0; b < nq; ++b) s.H(xo + a, xo + b) += w * L.Jp(c, b);
                }
            if (ts == 7)   // orientation part of e (end_effector_cost.h:72-81)
                for (int a = 0; a < nq; ++a)
                    for (int c = 0; c < 3; ++c) {
                        const double w = dt * P.ee_weight[3 + c] * L.Jo(c, a);
                        s.g[xo + a] += w * L.eo[c];
                        for (int b = 0; b < nq; ++b) s.H(xo + a, xo + b) += w * L.Jo(c, b);
                    }
            // inertial-alignment cost, Gauss-Newton (inertial_alignment.cpp:126-149)
            if (P.ia_cost_enabled)
                for (int a = 0; a < nx; ++a)
                    for (int r = 0; r < 2; ++r) {
                        const double w = dt * P.ia_cost_weight * L.Jea(r, a);
                        s.g[xo + a] += w * L.ea[r];
                        for (int b = 0; b < nx; ++b) s.H(xo + a, xo + b) += w * L.Jea(r, b);
                    }
            // dynamics gap
            const double* xn = X + size_t(k + 1) * nx;
            s.b.assign(nx, 0.0);
            for (int i = 0; i < nx; ++i) {
                double gap = -xn[i];
                for (int j = 0; j < nx; ++j) gap += W.A(i, j) * x[j];
                for (int j = 0; j < nq; ++j) gap += W.B(i, j) * u[j];
                s.b[i] = gap;
            }
            // input box (controller_interface.cpp:165-169,330-356)
            const int nbox_u = (D.nfc > 0) ? nu : nq;
            for (int i = 0; i < nbox_u; ++i) {
                Row r;
                r.idx = i;
                r.c = 0.0;
                r.lb = (i < nq ? P.input_lb[i] : P.force_lb) - u[i];
                r.ub = (i < nq ? P.input_ub[i] : P.force_ub) - u[i];
                finish(r, soft_u);
                s.rows.push_back(r);
            }
            // object-dynamics equality rows (balancing_constraints.cpp:114-155)
            for (int i = 0; i < D.neq; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                for (int j = 0; j < D.nfc; ++j) r.a[nq + j] = L.Df(i, j);
                for (int j = 0; j < nx; ++j) r.a[xo + j] = L.C(i, j);
                r.c = L.g[i];
                r.lb = r.ub = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
            // friction pyramid rows (balancing_constraints.cpp:32-71), h >= 0
            for (int i = 0; i < D.nfric; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                for (int j = 0; j < D.nfc; ++j) r.a[nq + j] = L.Ffric(i, j);
                r.c = L.hfric[i];
                r.lb = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
        }
        // state box, nodes 1..N (x_0 is fixed) (controller_interface.cpp:157-163)
        if (k >= 1)
            for (int i = 0; i < D.nxr; ++i) {   // bounds on the robot state only (controller_interface.cpp:157-163)
                Row r;
                r.idx = xo + i;
                r.lb = P.state_lb[i] - x[i];
                r.ub = P.state_ub[i] - x[i];
                finish(r, soft_x);
                s.rows.push_back(r);
            }
        // obstacle rows, nodes 1..N-1 (state-only; constant at node 0)
        if (k >= 1 && k < N) {
            for (int i = 0; i < D.npairs; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                for (int j = 0; j < nx; ++j) r.a[xo + j] = L.Jobs(i, j);
                r.c = L.hobs[i];
                r.lb = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
            // end-effector box (constraint/end_effector_box_constraint.h:46-76): three "upper" rows
            // r_d + upper - r >= 0 (Jacobian -J_p), then three "lower" rows r - r_d - lower >= 0 (+J_p)
            if (P.ee_box_enabled)
                for (int side = 0; side < 2; ++side)
                    for (int c = 0; c < 3; ++c) {
                        Row r;
                        r.a.assign(s.nz, 0.0);
                        const double sg = side == 0 ? -1.0 : 1.0;
                        for (int j = 0; j < nq; ++j) r.a[xo + j] = sg * L.Jp(c, j);
                        r.c = side == 0 ? rd[c] + P.ee_box_upper[c] - L.r[c]
                                        : L.r[c] - rd[c] - P.ee_box_lower[c];
                        r.lb = 0.0;
                        finish(r, soft_poly);
                        s.rows.push_back(r);
                    }
            // inertial-alignment constraint rows, dense over x (inertial_alignment.cpp:7-53)
            if (P.ia_constraint_enabled)
                for (int i = 0; i < 5; ++i) {
                    Row r;
                    r.a.assign(s.nz, 0.0);
                    for (int j = 0; j < nx; ++j) r.a[xo + j] = L.Jia(i, j);
                    r.c = L.hia[i];
                    r.lb = 0.0;
                    finish(r, soft_poly);
                    s.rows.push_back(r);
                }
            // projectile-path rows (projectile_path_constraint.h:108-146)
            for (int i = 0; i < D.nproj; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                for (int j = 0; j < nx; ++j) r.a[xo + j] = L.Jproj(i, j);
                r.c = L.hproj[i];
                r.lb = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
        }
        if (k == N) {
            // terminal equality [r_d - r; v; a] = 0 (stationary_desired_position_constraint.h:43-74)
            for (int i = 0; i < 3; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                for (int j = 0; j < nq; ++j) r.a[xo + j] = -L.Jp(i, j);
                r.c = rd[i] - L.r[i];
                r.lb = r.ub = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
            for (int i = nq; i < D.nxr; ++i) {
                Row r;
                r.a.assign(s.nz, 0.0);
                r.a[xo + i] = 1.0;
                r.c = x[i];
                r.lb = r.ub = 0.0;
                finish(r, soft_poly);
                s.rows.push_back(r);
            }
        }
    }
}

static inline double row_value(const Row& r, const double* z) {
    if (r.idx >= 0) return r.c + z[r.idx];
    double v = r.c;
    for (size_t j = 0; j < r.a.size(); ++j) v += r.a[j] * z[j];
    return v;
}
static inline double row_resid(const Row& r, double val) {
    const double s = val + r.lambda / r.rho;
    return s - std::min(std::max(s, r.lb), r.ub);
}

struct QpResult {
    int iters = 0;
    bool converged = false;
    double decrement = 0, hard_infeas = 0;
};

// Cholesky factor (lower) in place; returns false if not positive definite.
static bool cholesky(Mat& M, int n) {
    for (int j = 0; j < n; ++j) {
        double d = M(j, j);
        for (int k = 0; k < j; ++k) d -= M(j, k) * M(j, k);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        M(j, j) = d;
        for (int i = j + 1; i < n; ++i) {
            double s = M(i, j);
            for (int k = 0; k < j; ++k) s -= M(i, k) * M(j, k);
            M(i, j) = s / d;
        }
    }
    return true;
}
static void chol_solve(const Mat& L, int n, double* rhs) {
    for (int i = 0; i < n; ++i) {
        double s = rhs[i];
        for (int k = 0; k < i; ++k) s -= L(i, k) * rhs[k];
        rhs[i] = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = rhs[i];
        for (int k = i + 1; k < n; ++k) s -= L(k, i) * rhs[k];
        rhs[i] = s / L(i, i);
    }
}

// Semismooth-Newton / augmented-Lagrangian QP solve with Riccati linear algebra.
// dz[k] holds the stage iterate [du_k; dx_k]; gains (optional) receives K_k (nu x nx).
static QpResult solve_qp_ssn(const ub_problem_desc_t& P, Workspace& W, std::vector<std::vector<double>>& z,
                             std::vector<Mat>* gains) {
    const Dims& D = W.D;
    const int nx = D.nx, nu = D.nu, N = D.N;
    QpResult res;
    // dynamics-feasible start: du = 0, dx_0 = 0, dx_{k+1} = A dx_k + b_k
    z.assign(N + 1, std::vector<double>());
    for (int k = 0; k <= N; ++k) z[k].assign(W.st[k].nz, 0.0);
    for (int k = 0; k < N; ++k) {
        const double* dx = z[k].data() + nu;
        double* dxn = z[k + 1].data() + W.st[k + 1].nu;
        for (int i = 0; i < nx; ++i) {
            double v = W.st[k].b[i];
            for (int j = 0; j < nx; ++j) v += W.A(i, j) * dx[j];
            dxn[i] = v;
        }
    }
    std::vector<std::vector<double>> grad(N + 1), dir(N + 1);
    std::vector<Mat> Kk(N), Pk(N + 1);
    std::vector<std::vector<double>> kk(N), pk(N + 1);
    std::vector<std::vector<char>> active_prev(N + 1), active(N + 1);
    bool have_hard = false;
    for (auto& s : W.st)
        for (auto& r : s.rows) have_hard |= r.hard;
    double pinf_prev = INF;

    for (int it = 0; it < P.qp_iter_max; ++it) {
        res.iters = it + 1;
        // gradient + active set
        for (int k = 0; k <= N; ++k) {
            const Stage& s = W.st[k];
            grad[k].assign(s.nz, 0.0);
            active[k].assign(s.rows.size(), 0);
            for (int i = 0; i < s.nz; ++i) {
                double v = s.g[i];
                for (int j = 0; j < s.nz; ++j) v += s.H(i, j) * z[k][j];
                grad[k][i] = v;
            }
            for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                const Row& r = s.rows[ri];
                const double e = row_resid(r, row_value(r, z[k].data()));
                if (e != 0.0) {
                    active[k][ri] = 1;
                    if (r.idx >= 0) grad[k][r.idx] += r.rho * e;
                    else
                        for (int j = 0; j < s.nz; ++j) grad[k][j] += r.rho * e * r.a[j];
                }
            }
        }
        // backward Riccati sweep
        for (int k = N; k >= 0; --k) {
            const Stage& s = W.st[k];
            Mat M = s.H;
            std::vector<double> m = grad[k];
            for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                if (!active[k][ri]) continue;
                const Row& r = s.rows[ri];
                if (r.idx >= 0) M(r.idx, r.idx) += r.rho;
                else
                    for (int i = 0; i < s.nz; ++i) {
                        if (r.a[i] == 0.0) continue;
                        for (int j = 0; j < s.nz; ++j) M(i, j) += r.rho * r.a[i] * r.a[j];
                    }
            }
            if (k == N) {
                Pk[k] = M;
                pk[k] = m;
                continue;
            }
            // T = [B A] (nx x nz); M += T' P T ; m += T' p
            const Mat& Pn = Pk[k + 1];
            Mat T(nx, s.nz);
            for (int i = 0; i < nx; ++i) {
                for (int j = 0; j < nu; ++j) T(i, j) = W.B(i, j);
                for (int j = 0; j < nx; ++j) T(i, nu + j) = W.A(i, j);
            }
            Mat PT(nx, s.nz);
            for (int i = 0; i < nx; ++i)
                for (int l = 0; l < nx; ++l) {
                    const double pv = Pn(i, l);
                    if (pv == 0.0) continue;
                    for (int j = 0; j < s.nz; ++j) PT(i, j) += pv * T(l, j);
                }
            for (int l = 0; l < nx; ++l)
                for (int i = 0; i < s.nz; ++i) {
                    const double tv = T(l, i);
                    if (tv == 0.0) continue;
                    m[i] += tv * pk[k + 1][l];
                    for (int j = 0; j < s.nz; ++j) M(i, j) += tv * PT(l, j);
                }
            Mat Luu(nu, nu);
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nu; ++j) Luu(i, j) = M(i, j);
            if (!cholesky(Luu, nu)) {
                res.converged = false;
                res.decrement = std::numeric_limits<double>::quiet_NaN();
                return res;
            }
            Kk[k] = Mat(nu, nx);
            kk[k].assign(nu, 0.0);
            std::vector<double> col(nu);
            for (int j = 0; j < nx; ++j) {
                for (int i = 0; i < nu; ++i) col[i] = -M(i, nu + j);
                chol_solve(Luu, nu, col.data());
                for (int i = 0; i < nu; ++i) Kk[k](i, j) = col[i];
            }
            for (int i = 0; i < nu; ++i) col[i] = -m[i];
            chol_solve(Luu, nu, col.data());
            kk[k] = col;
            Pk[k] = Mat(nx, nx);
            pk[k].assign(nx, 0.0);
            for (int i = 0; i < nx; ++i) {
                double pv = m[nu + i];
                for (int l = 0; l < nu; ++l) pv += M(nu + i, l) * kk[k][l];
                pk[k][i] = pv;
                for (int j = 0; j < nx; ++j) {
                    double v = M(nu + i, nu + j);
                    for (int l = 0; l < nu; ++l) v += M(nu + i, l) * Kk[k](l, j);
                    Pk[k](i, j) = v;
                }
            }
        }
        // forward sweep (homogeneous dynamics, d x_0 = 0)
        std::vector<double> dx(nx, 0.0), dxn(nx);
        for (int k = 0; k < N; ++k) {
            dir[k].assign(W.st[k].nz, 0.0);
            for (int i = 0; i < nu; ++i) {
                double v = kk[k][i];
                for (int j = 0; j < nx; ++j) v += Kk[k](i, j) * dx[j];
                dir[k][i] = v;
            }
            for (int i = 0; i < nx; ++i) dir[k][nu + i] = dx[i];
            for (int i = 0; i < nx; ++i) {
                double v = 0;
                for (int j = 0; j < nx; ++j) v += W.A(i, j) * dx[j];
                for (int j = 0; j < nu; ++j) v += W.B(i, j) * dir[k][j];
                dxn[i] = v;
            }
            dx = dxn;
        }
        dir[N] = dx;

        // exact line search on phi(alpha) = f(z + alpha d): phi' is piecewise linear, increasing
        double c1 = 0, c2 = 0;
        struct LsRow { double s, ds, lb, ub, rho; };
        std::vector<LsRow> ls;
        for (int k = 0; k <= N; ++k) {
            const Stage& s = W.st[k];
            for (int i = 0; i < s.nz; ++i) {
                double hz = s.g[i], hd = 0;
                for (int j = 0; j < s.nz; ++j) {
                    hz += s.H(i, j) * z[k][j];
                    hd += s.H(i, j) * dir[k][j];
                }
                c1 += hz * dir[k][i];
                c2 += hd * dir[k][i];
            }
            for (const Row& r : s.rows) {
                const double val = row_value(r, z[k].data()) + r.lambda / r.rho;
                double dv;
                if (r.idx >= 0) dv = dir[k][r.idx];
                else {
                    dv = 0;
                    for (int j = 0; j < s.nz; ++j) dv += r.a[j] * dir[k][j];
                }
                ls.push_back({val, dv, r.lb, r.ub, r.rho});
            }
        }
        auto dphi = [&](double al, double* curv) {
            double d1 = c1 + al * c2, d2 = c2;
            for (const LsRow& r : ls) {
                const double s = r.s + al * r.ds;
                const double e = s - std::min(std::max(s, r.lb), r.ub);
                if (e != 0.0) {
                    d1 += r.rho * e * r.ds;
                    d2 += r.rho * r.ds * r.ds;
                }
            }
            if (curv) *curv = d2;
            return d1;
        };
        auto phi_delta = [&](double al) {
            double v = al * c1 + 0.5 * al * al * c2;
            for (const LsRow& r : ls) {
                const double s0 = r.s, s1 = r.s + al * r.ds;
                const double e0 = s0 - std::min(std::max(s0, r.lb), r.ub);
                const double e1 = s1 - std::min(std::max(s1, r.lb), r.ub);
                v += 0.5 * r.rho * (e1 * e1 - e0 * e0);
            }
            return v;
        };
        double curv0;
        const double slope0 = dphi(0.0, &curv0);
        res.decrement = -slope0;
        double alpha = 1.0;
        static const char* ls_mode = std::getenv("ORACLE_LS");
        if (slope0 < 0.0) {
            double curv;
            double s1 = dphi(1.0, &curv);
            if (ls_mode && s1 > 0.0 && phi_delta(1.0) <= 1e-4 * slope0) s1 = 0.0;  // accept full step
            if (s1 > 0.0) {
                double lo = 0.0, hi = 1.0, a = 1.0, sa = s1;
                for (int ls_it = 0; ls_it < 60; ++ls_it) {
                    double an = a - sa / curv;
                    if (!(an > lo && an < hi)) an = 0.5 * (lo + hi);
                    a = an;
                    sa = dphi(a, &curv);
                    if (sa > 0) hi = a; else lo = a;
                    if (std::fabs(sa) <= 1e-14 * std::fabs(slope0) || hi - lo < 1e-15) break;
                }
                alpha = a;
            }
        } else {
            alpha = 0.0;
        }
        for (int k = 0; k <= N; ++k)
            for (int i = 0; i < W.st[k].nz; ++i) z[k][i] += alpha * dir[k][i];
        if (std::getenv("ORACLE_TRACE")) {
            int nact = 0, nchg = 0;
            for (int k = 0; k <= N; ++k)
                for (size_t ri = 0; ri < active[k].size(); ++ri) {
                    nact += active[k][ri];
                    if (active_prev[k].size() == active[k].size()) nchg += active[k][ri] != active_prev[k][ri];
                }
            std::fprintf(stderr, "  qp it %2d  dec %.3e  alpha %.4f  active %d  changed %d\n", it, res.decrement, alpha,
                         nact, nchg);
        }

        // inner convergence: Newton decrement small relative to the scale of the problem
        double scale = 0;
        for (int k = 0; k <= N; ++k)
            for (int i = 0; i < W.st[k].nz; ++i) scale = std::max(scale, std::fabs(grad[k][i] * dir[k][i]));
        const bool inner_done = (res.decrement <= 1e-14) || (alpha == 1.0 && active == active_prev) ||
                                (alpha == 0.0);
        active_prev = active;
        (void)scale;
        if (!inner_done) continue;
        if (!have_hard) {
            res.converged = true;
            break;
        }
        // augmented-Lagrangian update of the hard rows
        double pinf = 0;
        for (int k = 0; k <= N; ++k)
            for (Row& r : W.st[k].rows) {
                if (!r.hard) continue;
                const double val = row_value(r, z[k].data());
                pinf = std::max(pinf, std::fabs(val - std::min(std::max(val, r.lb), r.ub)));
            }
        res.hard_infeas = pinf;
        if (std::getenv("ORACLE_TRACE")) std::fprintf(stderr, "  -- AL update: pinf %.3e\n", pinf);
        if (pinf <= 1e-9) {
            res.converged = true;
            break;
        }
        const bool grow = pinf > 0.25 * pinf_prev;
        pinf_prev = pinf;
        for (int k = 0; k <= N; ++k)
            for (Row& r : W.st[k].rows) {
                if (!r.hard) continue;
                r.lambda = r.rho * row_resid(r, row_value(r, z[k].data()));
                if (grow) r.rho = std::min(r.rho * 10.0, 1e8 * r.rho0);
            }
        active_prev.assign(N + 1, std::vector<char>());
    }
    if (gains) {
        gains->assign(N, Mat());
        for (int k = 0; k < N; ++k) (*gains)[k] = Kk[k];
    }
    return res;
}


// ---------------------------------------------------------------------------
// Riccati factorisation of the stage-wise Newton system, reusable for several
// right-hand sides (predictor + corrector).
struct Riccati {
    std::vector<Mat> Luu, K, Pm;
    bool factor(const Workspace& W, const std::vector<Mat>& M) {
        const int nx = W.D.nx, nu = W.D.nu, N = W.D.N;
        Luu.assign(N, Mat());
        K.assign(N, Mat());
        Pm.assign(N + 1, Mat());
        Pm[N] = M[N];
        for (int k = N - 1; k >= 0; --k) {
            const int nz = nu + nx;
            Mat Mk = M[k];
            const Mat& Pn = Pm[k + 1];
            Mat T(nx, nz);
            for (int i = 0; i < nx; ++i) {
                for (int j = 0; j < nu; ++j) T(i, j) = W.B(i, j);
                for (int j = 0; j < nx; ++j) T(i, nu + j) = W.A(i, j);
            }
            Mat PT(nx, nz);
            for (int i = 0; i < nx; ++i)
                for (int l = 0; l < nx; ++l) {
                    const double pv = Pn(i, l);
                    if (pv == 0.0) continue;
                    for (int j = 0; j < nz; ++j) PT(i, j) += pv * T(l, j);
                }
            for (int l = 0; l < nx; ++l)
                for (int i = 0; i < nz; ++i) {
                    const double tv = T(l, i);
                    if (tv == 0.0) continue;
                    for (int j = 0; j < nz; ++j) Mk(i, j) += tv * PT(l, j);
                }
            Luu[k] = Mat(nu, nu);
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nu; ++j) Luu[k](i, j) = Mk(i, j);
            if (!cholesky(Luu[k], nu)) return false;
            K[k] = Mat(nu, nx);
            std::vector<double> col(nu);
            for (int j = 0; j < nx; ++j) {
                for (int i = 0; i < nu; ++i) col[i] = -Mk(i, nu + j);
                chol_solve(Luu[k], nu, col.data());
                for (int i = 0; i < nu; ++i) K[k](i, j) = col[i];
            }
            Pm[k] = Mat(nx, nx);
            for (int i = 0; i < nx; ++i)
                for (int j = 0; j < nx; ++j) {
                    double v = Mk(nu + i, nu + j);
                    for (int l = 0; l < nu; ++l) v += Mk(nu + i, l) * K[k](l, j);
                    Pm[k](i, j) = v;
                }
            // keep the cost-to-go exactly symmetric: the rounding asymmetry of Mxx + Mxu K is amplified from stage
            // to stage and breaks the recursion after ~60 stages (the N = 100 robust-planner horizon)
            for (int i = 0; i < nx; ++i)
                for (int j = 0; j < i; ++j) Pm[k](i, j) = Pm[k](j, i) = 0.5 * (Pm[k](i, j) + Pm[k](j, i));
        }
        return true;
    }
    // minimise 1/2 d'Md + grad'd over homogeneous dynamics with d x_0 = 0
    void solve(const Workspace& W, const std::vector<std::vector<double>>& grad,
               std::vector<std::vector<double>>& dir) const {
        const int nx = W.D.nx, nu = W.D.nu, N = W.D.N;
        std::vector<std::vector<double>> kk(N);
        std::vector<double> p = grad[N], pn(nx);
        for (int k = N - 1; k >= 0; --k) {
            // m = grad_k + [B A]' p
            std::vector<double> m = grad[k];
            for (int l = 0; l < nx; ++l) {
                for (int j = 0; j < nu; ++j) m[j] += W.B(l, j) * p[l];
                for (int j = 0; j < nx; ++j) m[nu + j] += W.A(l, j) * p[l];
            }
            kk[k].assign(nu, 0.0);
            for (int i = 0; i < nu; ++i) kk[k][i] = -m[i];
            chol_solve(Luu[k], nu, kk[k].data());
            for (int i = 0; i < nx; ++i) {
                double v = m[nu + i];
                for (int l = 0; l < nu; ++l) v += K[k](l, i) * m[l];
                pn[i] = v;
            }
            p = pn;
        }
        dir.assign(N + 1, std::vector<double>());
        std::vector<double> dx(nx, 0.0), dxn(nx);
        for (int k = 0; k < N; ++k) {
            dir[k].assign(nu + nx, 0.0);
            for (int i = 0; i < nu; ++i) {
                double v = kk[k][i];
                for (int j = 0; j < nx; ++j) v += K[k](i, j) * dx[j];
                dir[k][i] = v;
            }
            for (int i = 0; i < nx; ++i) dir[k][nu + i] = dx[i];
            for (int i = 0; i < nx; ++i) {
                double v = 0;
                for (int j = 0; j < nx; ++j) v += W.A(i, j) * dx[j];
                for (int j = 0; j < nu; ++j) v += W.B(i, j) * dir[k][j];
                dxn[i] = v;
            }
            dx = dxn;
        }
        dir[N] = dx;
    }
};

static inline void add_row_outer(Mat& M, const Row& r, double w, int nz) {
    if (r.idx >= 0) {
        M(r.idx, r.idx) += w;
        return;
    }
    for (int i = 0; i < nz; ++i) {
        if (r.a[i] == 0.0) continue;
        const double wi = w * r.a[i];
        for (int j = 0; j < nz; ++j) M(i, j) += wi * r.a[j];
    }
}
static inline void add_row_vec(std::vector<double>& g, const Row& r, double w) {
    if (r.idx >= 0) {
        g[r.idx] += w;
        return;
    }
    for (size_t j = 0; j < r.a.size(); ++j) g[j] += w * r.a[j];
}
static inline double row_dot(const Row& r, const std::vector<double>& d) {
    if (r.idx >= 0) return d[r.idx];
    double v = 0;
    for (size_t j = 0; j < r.a.size(); ++j) v += r.a[j] * d[j];
    return v;
}

// Primal-dual interior-point solve of the OCP-QP (Mehrotra predictor-corrector,
// Riccati linear algebra) — the method class of HPIPM, which the reference
// calls through ocs2_sqp.  Inequality side j of a row:
//     d_j(z) + eps_j*lambda_j - t_j = 0,  t_j, lambda_j >= 0,  t_j*lambda_j -> mu
// with eps_j = 1/Z for a row softened by an L2 slack (eliminating the slack
// s_j = lambda_j/Z of HPIPM's soft-constraint formulation; zero L1 weight and
// zero slack bound) and eps_j = 0 for a hard row.  Equality rows (lb == ub):
// softened -> exact quadratic penalty Z/2 r^2 in the Hessian; hard -> proximal
// method of multipliers on the unit-normalised row.  Iterates stay
// dynamics-feasible; the centring target is floored at qp_mu_target so that
// the result is the central-path point at that mu.
static QpResult solve_qp_ipm(const ub_problem_desc_t& P, Workspace& W, std::vector<std::vector<double>>& z,
                             std::vector<Mat>* gains) {
    const Dims& D = W.D;
    const int nx = D.nx, nu = D.nu, N = D.N;
    QpResult res;
    z.assign(N + 1, std::vector<double>());
    for (int k = 0; k <= N; ++k) z[k].assign(W.st[k].nz, 0.0);
    for (int k = 0; k < N; ++k) {
        const double* dx = z[k].data() + nu;
        double* dxn = z[k + 1].data() + W.st[k + 1].nu;
        for (int i = 0; i < nx; ++i) {
            double v = W.st[k].b[i];
            for (int j = 0; j < nx; ++j) v += W.A(i, j) * dx[j];
            dxn[i] = v;
        }
    }
    const double sgn[2] = {1.0, -1.0};
    auto side_on = [](const Row& r, int s) { return r.lb < r.ub && std::isfinite(s == 0 ? r.lb : r.ub); };
    auto side_d = [&](const Row& r, int s, double val) { return s == 0 ? val - r.lb : r.ub - val; };
    // hard inequality rows keep a tiny dual regularisation (elastic mode: an infeasible QP
    // degrades to a 1/eps-weighted least-violation problem instead of diverging)
    auto side_eps = [](const Row& r) { return r.hard ? 1.0e-6 : 1.0 / r.rho; };
    // initialisation
    long nsides = 0;
    for (int k = 0; k <= N; ++k)
        for (Row& r : W.st[k].rows) {
            const double val = row_value(r, z[k].data());
            for (int s = 0; s < 2; ++s) {
                if (!side_on(r, s)) continue;
                r.t[s] = std::max(side_d(r, s, val), P.qp_thr0);
                r.lam[s] = P.qp_mu0 / r.t[s];
                ++nsides;
            }
        }
    std::vector<Mat> M(N + 1);
    std::vector<std::vector<double>> rg(N + 1), grad(N + 1), dz_aff, dz;
    Riccati ric;
    struct Dl { double dt[2], dlam[2]; };
    std::vector<std::vector<Dl>> stepv(N + 1);
    const bool trace = std::getenv("ORACLE_TRACE") != nullptr;
    double last_alpha = 0.0, last_step = INF, pinf_prev = INF;
    int stall = 0;

    for (int it = 0; it < P.qp_iter_max; ++it) {
        // residuals, barrier weights, base gradient
        double mu = 0, rd_max = 0, pinf = 0;
        for (int k = 0; k <= N; ++k) {
            Stage& s = W.st[k];
            M[k] = s.H;
            rg[k].assign(s.nz, 0.0);
            for (int i = 0; i < s.nz; ++i) {
                double v = s.g[i];
                for (int j = 0; j < s.nz; ++j) v += s.H(i, j) * z[k][j];
                rg[k][i] = v;
            }
            for (Row& r : s.rows) {
                const double val = row_value(r, z[k].data());
                if (!(r.lb < r.ub)) {  // equality row
                    const double e = val - r.lb;
                    if (r.rho <= 0.0) continue;
                    add_row_outer(M[k], r, r.rho, s.nz);
                    add_row_vec(rg[k], r, r.rho * e + r.lambda);
                    if (r.hard) pinf = std::max(pinf, std::fabs(e));  // raw row units, as HPIPM's res_eq
                    continue;
                }
                const double eps = side_eps(r);
                for (int sd = 0; sd < 2; ++sd) {
                    if (!side_on(r, sd)) continue;
                    const double rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                    rd_max = std::max(rd_max, std::fabs(rd));
                    mu += r.t[sd] * r.lam[sd];
                    add_row_outer(M[k], r, r.lam[sd] / (r.t[sd] + eps * r.lam[sd]), s.nz);
                    add_row_vec(rg[k], r, -sgn[sd] * r.lam[sd]);
                }
            }
        }
        mu = nsides > 0 ? mu / nsides : 0.0;
        const bool central = it > 0 && mu <= 2.0 * P.qp_mu_target && rd_max <= P.qp_tol && last_alpha >= 0.5;
        if (central && (pinf <= P.qp_tol || last_step <= P.qp_tol)) {
            res.converged = true;
            break;
        }
        // Hopeless hard equality rows: from the fifth iteration on, at the linear rate the multiplier method shows
        // (pinf / pinf_prev) the rows cannot reach the tolerance before the iteration cap — three iterations in a row.
        // Such a QP (rows inconsistent at this linearisation point, e.g. a start state from which the object cannot be
        // balanced) would creep towards its least-violation point until the cap; it ends here as NOT converged
        // (DESIGN.md section 4, item 5).
        {
            bool hopeless = false;
            if (it >= 4 && pinf > P.qp_tol && pinf_prev < INF) {
                const double ratio = pinf / pinf_prev;
                hopeless = ratio >= 1.0 || double(it) + std::log(P.qp_tol / pinf) / std::log(ratio) > double(P.qp_iter_max);
            }
            stall = hopeless ? stall + 1 : 0;
        }
        pinf_prev = pinf;
        if (stall >= 3) break;
        res.iters = it + 1;
        if (!ric.factor(W, M)) {
            res.decrement = std::numeric_limits<double>::quiet_NaN();
            return res;
        }
        auto solve_with = [&](bool corrector, double target, std::vector<std::vector<double>>& out) {
            for (int k = 0; k <= N; ++k) {
                const Stage& s = W.st[k];
                grad[k] = rg[k];
                for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                    const Row& r = s.rows[ri];
                    if (!(r.lb < r.ub)) continue;
                    const double val = row_value(r, z[k].data());
                    const double eps = side_eps(r);
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!side_on(r, sd)) continue;
                        const double rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        double rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += stepv[k][ri].dt[sd] * stepv[k][ri].dlam[sd];
                        add_row_vec(grad[k], r, sgn[sd] * (rc + r.lam[sd] * rd) / (r.t[sd] + eps * r.lam[sd]));
                    }
                }
            }
            ric.solve(W, grad, out);
            // recover d lambda, d t
            for (int k = 0; k <= N; ++k) {
                const Stage& s = W.st[k];
                std::vector<Dl> nv(s.rows.size());
                for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                    const Row& r = s.rows[ri];
                    if (!(r.lb < r.ub)) continue;
                    const double val = row_value(r, z[k].data());
                    const double eps = side_eps(r);
                    const double adz = row_dot(r, out[k]);
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!side_on(r, sd)) continue;
                        const double rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        double rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += stepv[k][ri].dt[sd] * stepv[k][ri].dlam[sd];
                        const double den = r.t[sd] + eps * r.lam[sd];
                        const double dl = -(rc + r.lam[sd] * rd) / den - (r.lam[sd] / den) * sgn[sd] * adz;
                        nv[ri].dlam[sd] = dl;
                        nv[ri].dt[sd] = sgn[sd] * adz + eps * dl + rd;
                    }
                }
                stepv[k] = nv;
            }
        };
        auto max_step = [&]() {
            double a = 1.0;
            for (int k = 0; k <= N; ++k) {
                const Stage& s = W.st[k];
                for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                    const Row& r = s.rows[ri];
                    if (!(r.lb < r.ub)) continue;
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!side_on(r, sd)) continue;
                        if (stepv[k][ri].dt[sd] < 0) a = std::min(a, -r.t[sd] / stepv[k][ri].dt[sd]);
                        if (stepv[k][ri].dlam[sd] < 0) a = std::min(a, -r.lam[sd] / stepv[k][ri].dlam[sd]);
                    }
                }
            }
            return a;
        };
        for (int k = 0; k <= N; ++k) stepv[k].assign(W.st[k].rows.size(), Dl{{0, 0}, {0, 0}});
        double target = P.qp_mu_target;
        if (nsides > 0) {
            // predictor
            solve_with(false, 0.0, dz_aff);
            const double a_aff = max_step();
            double mu_aff = 0;
            for (int k = 0; k <= N; ++k) {
                const Stage& s = W.st[k];
                for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                    const Row& r = s.rows[ri];
                    if (!(r.lb < r.ub)) continue;
                    for (int sd = 0; sd < 2; ++sd)
                        if (side_on(r, sd))
                            mu_aff += (r.t[sd] + a_aff * stepv[k][ri].dt[sd]) * (r.lam[sd] + a_aff * stepv[k][ri].dlam[sd]);
                }
            }
            mu_aff /= nsides;
            const double sigma = std::pow(mu_aff / mu, 3.0);
            target = std::max(sigma * mu, P.qp_mu_target);
            solve_with(true, target, dz);
        } else {
            solve_with(false, 0.0, dz);
        }
        double alpha = nsides > 0 ? std::min(1.0, 0.995 * max_step()) : 1.0;
        double dec = 0;
        for (int k = 0; k <= N; ++k) {
            Stage& s = W.st[k];
            for (int i = 0; i < s.nz; ++i) {
                dec -= grad[k][i] * dz[k][i];
                z[k][i] += alpha * dz[k][i];
            }
            for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                Row& r = s.rows[ri];
                if (!(r.lb < r.ub)) {
                    if (r.hard && r.rho > 0.0) r.lambda += r.rho * (row_value(r, z[k].data()) - r.lb);
                    continue;
                }
                for (int sd = 0; sd < 2; ++sd) {
                    if (!side_on(r, sd)) continue;
                    r.t[sd] += alpha * stepv[k][ri].dt[sd];
                    r.lam[sd] += alpha * stepv[k][ri].dlam[sd];
                }
            }
        }
        last_alpha = alpha;
        last_step = 0.0;
        for (int k = 0; k <= N; ++k)
            for (int i = 0; i < W.st[k].nz; ++i) last_step = std::max(last_step, std::fabs(alpha * dz[k][i]));
        res.decrement = dec;
        res.hard_infeas = pinf;
        if (trace)
            std::fprintf(stderr, "  ipm it %2d  mu %.3e  rd %.3e  pinf %.3e  target %.3e  alpha %.4f  dec %.3e\n", it, mu,
                         rd_max, pinf, target, alpha, dec);
    }
    if (gains) *gains = ric.K;
    return res;
}

static QpResult solve_qp(const ub_problem_desc_t& P, Workspace& W, std::vector<std::vector<double>>& z,
                         std::vector<Mat>* gains) {
    return P.qp_method == 1 ? solve_qp_ssn(P, W, z, gains) : solve_qp_ipm(P, W, z, gains);
}

// One MPC solve (advanceMpc): `sqp_iteration` SQP iterations.
static void solve_one(const ub_problem_desc_t& P, const double* x0, const double* target, const double* body_params,
                      bool warm, double* X, double* U, double* Kout, int32_t* status, double* stats) {
    Workspace W;
    W.D = make_dims(P);
    const Dims& D = W.D;
    const int nx = D.nx, nu = D.nu, N = D.N;
    discrete_dynamics(P, D, W.A, W.B);
    // initial guess: DefaultInitializer = zero input, state held (controller_interface.cpp:385-386)
    if (!warm) {
        for (int k = 0; k <= N; ++k) std::memcpy(X + size_t(k) * nx, x0, sizeof(double) * nx);
        std::fill(U, U + size_t(N) * nu, 0.0);
    } else {
        std::memcpy(X, x0, sizeof(double) * nx);  // x_0 is the observation
    }
    int st = UB_STATUS_CONVERGED;
    double alpha = 0, qp_dec = 0;
    int qp_iters = 0, sqp_done = 0;
    Perf base = performance(P, D, W.A, W.B, body_params, target, X, U);
    std::vector<Mat> gains;
    for (int it = 0; it < std::max(1, P.sqp_iteration); ++it) {
        ++sqp_done;
        build_qp(P, W, body_params, target, X, U);
        std::vector<std::vector<double>> z;
        const QpResult qr = solve_qp(P, W, z, Kout ? &gains : nullptr);
        qp_iters += qr.iters;
        qp_dec = qr.decrement;
        if (!std::isfinite(qr.decrement)) {
            st = UB_STATUS_NAN;
            break;
        }
        if (!qr.converged) st = UB_STATUS_QP_MAXITER;
        // Armijo descent metric: cost gradient along the step
        double descent = 0;
        for (int k = 0; k <= N; ++k)
            for (int i = 0; i < W.st[k].nz; ++i) descent += W.st[k].g[i] * z[k][i];
        // filter line search (ocs2 FilterLinesearch [EXT]; settings names as in ocs2_sqp)
        std::vector<double> Xn(size_t(N + 1) * nx), Un(size_t(N) * nu);
        const double vb = base.violation();
        bool accepted = false;
        Perf pn;
        alpha = 1.0;
        while (alpha >= P.alpha_min) {
            for (int k = 0; k <= N; ++k) {
                const int xo = W.st[k].nu;
                for (int i = 0; i < nx; ++i) Xn[size_t(k) * nx + i] = X[size_t(k) * nx + i] + alpha * z[k][xo + i];
                if (k < N)
                    for (int i = 0; i < nu; ++i) Un[size_t(k) * nu + i] = U[size_t(k) * nu + i] + alpha * z[k][i];
            }
            pn = performance(P, D, W.A, W.B, body_params, target, Xn.data(), Un.data());
            const double vn = pn.violation();
            if (vn > P.g_max) {
                accepted = false;
            } else if (vn < P.g_min) {
                if (vb < P.g_min && descent < 0.0) accepted = pn.cost < base.cost + P.armijo_factor * alpha * descent;
                else accepted = true;
            } else {
                accepted = (vn < (1.0 - P.gamma_c) * vb) || (pn.cost < base.cost - P.gamma_c * vb);
            }
            if (accepted) break;
            alpha *= P.alpha_decay;
        }
        if (!accepted) {
            st = UB_STATUS_LS_FAILED;
            alpha = 0.0;
            break;
        }
        double dxn = 0, dun = 0;
        for (int k = 0; k <= N; ++k) {
            const int xo = W.st[k].nu;
            for (int i = 0; i < nx; ++i) dxn += sq(alpha * z[k][xo + i]);
            for (int i = 0; i < W.st[k].nu; ++i) dun += sq(alpha * z[k][i]);
        }
        std::memcpy(X, Xn.data(), sizeof(double) * Xn.size());
        std::memcpy(U, Un.data(), sizeof(double) * Un.size());
        const double dcost = std::fabs(pn.cost - base.cost);
        base = pn;
        if ((std::sqrt(dxn) < P.delta_tol && std::sqrt(dun) < P.delta_tol) ||
            (dcost < P.cost_tol && base.violation() < P.g_min))
            break;
    }
    if (Kout && !gains.empty())
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < nu; ++i)
                for (int j = 0; j < nx; ++j) Kout[(size_t(k) * nu + i) * nx + j] = gains[k](i, j);
    for (size_t i = 0; i < size_t(N + 1) * nx; ++i)
        if (!std::isfinite(X[i])) st = UB_STATUS_NAN;
    *status = st;
    if (stats) {
        stats[0] = qp_iters;
        stats[1] = base.cost;
        stats[2] = base.violation();
        stats[3] = alpha;
        stats[4] = qp_dec;
        stats[5] = base.max_eq;
        stats[6] = base.min_margin;
        stats[7] = sqp_done;
    }
}


// ---------------------------------------------------------------------------
// Precision study (test infrastructure, tools/precision_lab.py): the interior-point iteration of solve_qp_ipm with
// the Newton matrices, their Riccati factorisation and the direction in type F, and the iterate, the slack records
// and every residual in type R.  <double, double> reproduces solve_qp_ipm; <float, float> is the arithmetic of the
// fp32 kernels (same method, dense recursion instead of their blocked one); <float, double> is the
// mixed-precision refinement proposed in DESIGN.md section 9.
template <typename R>
struct RowT {
    int idx = -1;
    std::vector<R> a;
    R c = 0, lb = 0, ub = 0, rho = 0, lambda = 0;
    bool hard = false, eq = false, on[2] = {false, false};
    R t[2] = {0, 0}, lam[2] = {0, 0}, dt[2] = {0, 0}, dl[2] = {0, 0};
};
template <typename R>
struct StageT {
    int nz = 0, nu = 0;
    std::vector<R> H, g, b;
    std::vector<RowT<R>> rows;
};
template <typename R>
static inline R row_value_t(const RowT<R>& r, const R* z) {
    if (r.idx >= 0) return r.c + z[r.idx];
    R v = r.c;
    for (size_t j = 0; j < r.a.size(); ++j) v += r.a[j] * z[j];
    return v;
}
template <typename R, typename V>
static inline R row_dot_t(const RowT<R>& r, const V* d) {
    if (r.idx >= 0) return R(d[r.idx]);
    R v = 0;
    for (size_t j = 0; j < r.a.size(); ++j) v += r.a[j] * R(d[j]);
    return v;
}

template <typename F>
struct RiccatiT {
    // The recursion in the form the kernels use (ub_solver.cuh: stage_factor_blocked / partial_cholesky): partial
    // Cholesky of the first nu columns of the stage matrix M_k + [B A]' P_{k+1} [B A]; the trailing block it leaves
    // is the Schur complement P_k (symmetric by construction), pivots floored at reg_input.
    int nx = 0, nu = 0, N = 0;
    F pivot_floor = 0;
    std::vector<F> A, B;                 // nx x nx, nx x nu
    std::vector<std::vector<F>> L;       // per stage: nz x nu, column j = column j of the factor (rows j..nz-1)
    std::vector<F> P;                    // running cost-to-go Hessian, nx x nx
    bool factor(const std::vector<std::vector<F>>& M) {
        const int nz = nu + nx;
        L.assign(N, {});
        P = M[N];
        std::vector<F> T(size_t(nx) * nz);
        for (int i = 0; i < nx; ++i) {
            for (int j = 0; j < nu; ++j) T[i * nz + j] = B[i * nu + j];
            for (int j = 0; j < nx; ++j) T[i * nz + nu + j] = A[i * nx + j];
        }
        for (int k = N - 1; k >= 0; --k) {
            std::vector<F> Mk = M[k];
            std::vector<F> PT(size_t(nx) * nz, F(0));
            for (int i = 0; i < nx; ++i)
                for (int l = 0; l < nx; ++l) {
                    const F pv = P[i * nx + l];
                    if (pv == F(0)) continue;
                    for (int j = 0; j < nz; ++j) PT[i * nz + j] += pv * T[l * nz + j];
                }
            for (int l = 0; l < nx; ++l)
                for (int i = 0; i < nz; ++i) {
                    const F tv = T[l * nz + i];
                    if (tv == F(0)) continue;
                    for (int j = 0; j <= i; ++j) Mk[i * nz + j] += tv * PT[l * nz + j];   // lower triangle
                }
            L[k].assign(size_t(nz) * nu, F(0));
            for (int j = 0; j < nu; ++j) {
                F d = Mk[j * nz + j];
                if (d != d) return false;
                if (!(d > pivot_floor)) {
                    if (!(pivot_floor > F(0))) return false;
                    d = pivot_floor;
                }
                const F inv = F(1) / std::sqrt(d);
                L[k][j * nu + j] = std::sqrt(d);
                for (int i = j + 1; i < nz; ++i) L[k][i * nu + j] = Mk[i * nz + j] * inv;
                for (int i = j + 1; i < nz; ++i) {
                    const F lij = L[k][i * nu + j];
                    for (int l = j + 1; l <= i; ++l) Mk[i * nz + l] -= lij * L[k][l * nu + j];
                }
            }
            for (int i = 0; i < nx; ++i)
                for (int j = 0; j <= i; ++j) P[i * nx + j] = P[j * nx + i] = Mk[(nu + i) * nz + nu + j];
        }
        return true;
    }
    void solve(const std::vector<std::vector<F>>& grad, std::vector<std::vector<F>>& dir) const {
        std::vector<std::vector<F>> w(N);
        std::vector<F> p = grad[N];
        for (int k = N - 1; k >= 0; --k) {
            std::vector<F> m = grad[k];
            for (int l = 0; l < nx; ++l) {
                for (int j = 0; j < nu; ++j) m[j] += B[l * nu + j] * p[l];
                for (int j = 0; j < nx; ++j) m[nu + j] += A[l * nx + j] * p[l];
            }
            w[k].assign(nu, F(0));
            for (int i = 0; i < nu; ++i) {          // L_uu w = m_u
                F s = m[i];
                for (int j = 0; j < i; ++j) s -= L[k][i * nu + j] * w[k][j];
                w[k][i] = s / L[k][i * nu + i];
            }
            for (int i = 0; i < nx; ++i) {          // p = m_x - L_xu w
                F s = m[nu + i];
                for (int j = 0; j < nu; ++j) s -= L[k][(nu + i) * nu + j] * w[k][j];
                p[i] = s;
            }
        }
        dir.assign(N + 1, {});
        std::vector<F> dx(nx, F(0)), dxn(nx);
        for (int k = 0; k < N; ++k) {
            dir[k].assign(nu + nx, F(0));
            std::vector<F> rhs(nu);
            for (int j = 0; j < nu; ++j) {          // s = w + L_xu' dx
                F s = w[k][j];
                for (int i = 0; i < nx; ++i) s += L[k][(nu + i) * nu + j] * dx[i];
                rhs[j] = s;
            }
            for (int j = nu - 1; j >= 0; --j) {     // du = - L_uu^-T s
                F s = -rhs[j];
                for (int i = j + 1; i < nu; ++i) s -= L[k][i * nu + j] * dir[k][i];
                dir[k][j] = s / L[k][j * nu + j];
            }
            for (int i = 0; i < nx; ++i) dir[k][nu + i] = dx[i];
            for (int i = 0; i < nx; ++i) {
                F v = 0;
                for (int j = 0; j < nx; ++j) v += A[i * nx + j] * dx[j];
                for (int j = 0; j < nu; ++j) v += B[i * nu + j] * dir[k][j];
                dxn[i] = v;
            }
            dx = dxn;
        }
        dir[N] = dx;
    }
};

// info: {iterations, converged, factorisation failed (iteration index + 1, else 0), last mu, last max slack residual}
template <typename F, typename R>
static void solve_qp_ipm_precision(const ub_problem_desc_t& P, const Workspace& W, std::vector<std::vector<double>>& zout,
                                   double* info, int extra = 0, double mu_factor = 2.0) {
    int extra_left = extra;
    const Dims& D = W.D;
    const int nx = D.nx, nu = D.nu, N = D.N;
    std::vector<StageT<R>> st(N + 1);
    for (int k = 0; k <= N; ++k) {
        const Stage& s = W.st[k];
        StageT<R>& q = st[k];
        q.nz = s.nz;
        q.nu = s.nu;
        q.H.assign(s.H.d.begin(), s.H.d.end());
        q.g.assign(s.g.begin(), s.g.end());
        q.b.assign(s.b.begin(), s.b.end());
        for (const Row& r : s.rows) {
            RowT<R> t;
            t.idx = r.idx;
            t.a.assign(r.a.begin(), r.a.end());
            t.c = R(r.c);
            t.eq = !(r.lb < r.ub);
            t.lb = R(std::isfinite(r.lb) ? r.lb : 0.0);
            t.ub = R(std::isfinite(r.ub) ? r.ub : 0.0);
            t.on[0] = !t.eq && std::isfinite(r.lb);
            t.on[1] = !t.eq && std::isfinite(r.ub);
            t.rho = R(r.rho);
            t.lambda = R(r.lambda);
            t.hard = r.hard;
            q.rows.push_back(t);
        }
    }
    RiccatiT<F> ric;
    ric.nx = nx; ric.nu = nu; ric.N = N;
    ric.pivot_floor = F(P.reg_input);
    ric.A.assign(W.A.d.begin(), W.A.d.end());
    ric.B.assign(W.B.d.begin(), W.B.d.end());
    std::vector<std::vector<R>> z(N + 1);
    for (int k = 0; k <= N; ++k) z[k].assign(st[k].nz, R(0));
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nx; ++i) {
            R v = st[k].b[i];
            for (int j = 0; j < nx; ++j) v += R(W.A(i, j)) * z[k][nu + j];
            z[k + 1][st[k + 1].nu + i] = v;
        }
    const R sgn[2] = {R(1), R(-1)};
    auto side_d = [](const RowT<R>& r, int s, R val) { return s == 0 ? val - r.lb : r.ub - val; };
    auto side_eps = [](const RowT<R>& r) { return r.hard ? R(1.0e-6) : R(1) / r.rho; };
    long nsides = 0;
    for (int k = 0; k <= N; ++k)
        for (RowT<R>& r : st[k].rows) {
            const R val = row_value_t(r, z[k].data());
            for (int s = 0; s < 2; ++s) {
                if (!r.on[s]) continue;
                r.t[s] = std::max(side_d(r, s, val), R(P.qp_thr0));
                r.lam[s] = R(P.qp_mu0) / r.t[s];
                ++nsides;
            }
        }
    std::vector<std::vector<F>> M(N + 1), grad(N + 1), dz;
    std::vector<std::vector<R>> rg(N + 1);
    R last_alpha = 0, last_step = std::numeric_limits<R>::infinity(), mu = 0, rd_max = 0;
    int iters = 0, failed = 0;
    bool converged = false;
    for (int it = 0; it < P.qp_iter_max; ++it) {
        mu = 0;
        rd_max = 0;
        R pinf = 0;
        for (int k = 0; k <= N; ++k) {
            StageT<R>& s = st[k];
            const int nz = s.nz;
            M[k].assign(s.H.begin(), s.H.end());
            rg[k].assign(nz, R(0));
            for (int i = 0; i < nz; ++i) {
                R v = s.g[i];
                for (int j = 0; j < nz; ++j) v += s.H[i * nz + j] * z[k][j];
                rg[k][i] = v;
            }
            auto add_outer = [&](const RowT<R>& r, R w) {
                if (r.idx >= 0) {
                    M[k][r.idx * nz + r.idx] += F(w);
                    return;
                }
                for (int i = 0; i < nz; ++i) {
                    if (r.a[i] == R(0)) continue;
                    const F wi = F(w) * F(r.a[i]);
                    for (int j = 0; j < nz; ++j) M[k][i * nz + j] += wi * F(r.a[j]);
                }
            };
            auto add_vec = [&](const RowT<R>& r, R w) {
                if (r.idx >= 0) {
                    rg[k][r.idx] += w;
                    return;
                }
                for (int j = 0; j < nz; ++j) rg[k][j] += w * r.a[j];
            };
            for (RowT<R>& r : s.rows) {
                const R val = row_value_t(r, z[k].data());
                if (r.eq) {
                    const R e = val - r.lb;
                    if (r.rho <= R(0)) continue;
                    add_outer(r, r.rho);
                    add_vec(r, r.rho * e + r.lambda);
                    if (r.hard) pinf = std::max(pinf, R(std::fabs(e)));
                    continue;
                }
                const R eps = side_eps(r);
                for (int sd = 0; sd < 2; ++sd) {
                    if (!r.on[sd]) continue;
                    const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                    rd_max = std::max(rd_max, R(std::fabs(rd)));
                    mu += r.t[sd] * r.lam[sd];
                    add_outer(r, r.lam[sd] / (r.t[sd] + eps * r.lam[sd]));
                    add_vec(r, -sgn[sd] * r.lam[sd]);
                }
            }
        }
        mu = nsides > 0 ? mu / R(nsides) : R(0);
        if (it > 0 && mu <= R(mu_factor * P.qp_mu_target) && rd_max <= R(P.qp_tol) && last_alpha >= R(0.5) &&
            (pinf <= R(P.qp_tol) || last_step <= R(P.qp_tol))) {
            converged = true;
            // `extra` further Newton iterations on the exact residual: what an inexact (fp32) direction leaves of the
            // stationarity residual contracts by the accuracy of the factorisation per iteration
            if (extra_left-- <= 0) break;
        }
        iters = it + 1;
        if (!ric.factor(M)) {
            failed = it + 1;
            break;
        }
        auto solve_with = [&](bool corrector, R target) {
            for (int k = 0; k <= N; ++k) {
                StageT<R>& s = st[k];
                std::vector<R> gk = rg[k];
                for (RowT<R>& r : s.rows) {
                    if (r.eq) continue;
                    const R val = row_value_t(r, z[k].data());
                    const R eps = side_eps(r);
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        R rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += r.dt[sd] * r.dl[sd];
                        const R w = sgn[sd] * (rc + r.lam[sd] * rd) / (r.t[sd] + eps * r.lam[sd]);
                        if (r.idx >= 0) gk[r.idx] += w;
                        else for (int j = 0; j < s.nz; ++j) gk[j] += w * r.a[j];
                    }
                }
                grad[k].assign(gk.begin(), gk.end());
            }
            ric.solve(grad, dz);
            for (int k = 0; k <= N; ++k)
                for (RowT<R>& r : st[k].rows) {
                    if (r.eq) continue;
                    const R val = row_value_t(r, z[k].data());
                    const R eps = side_eps(r);
                    const R adz = row_dot_t(r, dz[k].data());
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        R rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += r.dt[sd] * r.dl[sd];
                        const R den = r.t[sd] + eps * r.lam[sd];
                        const R dl = -(rc + r.lam[sd] * rd) / den - (r.lam[sd] / den) * sgn[sd] * adz;
                        r.dl[sd] = dl;
                        r.dt[sd] = sgn[sd] * adz + eps * dl + rd;
                    }
                }
        };
        auto max_step = [&]() {
            R a = 1;
            for (int k = 0; k <= N; ++k)
                for (const RowT<R>& r : st[k].rows) {
                    if (r.eq) continue;
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        if (r.dt[sd] < R(0)) a = std::min(a, -r.t[sd] / r.dt[sd]);
                        if (r.dl[sd] < R(0)) a = std::min(a, -r.lam[sd] / r.dl[sd]);
                    }
                }
            return a;
        };
        for (int k = 0; k <= N; ++k)
            for (RowT<R>& r : st[k].rows) r.dt[0] = r.dt[1] = r.dl[0] = r.dl[1] = R(0);
        if (nsides > 0) {
            solve_with(false, R(0));
            const R a_aff = max_step();
            R mu_aff = 0;
            for (int k = 0; k <= N; ++k)
                for (const RowT<R>& r : st[k].rows) {
                    if (r.eq) continue;
                    for (int sd = 0; sd < 2; ++sd)
                        if (r.on[sd]) mu_aff += (r.t[sd] + a_aff * r.dt[sd]) * (r.lam[sd] + a_aff * r.dl[sd]);
                }
            mu_aff /= R(nsides);
            const R ratio = mu_aff / mu;
            solve_with(true, std::max(ratio * ratio * ratio * mu, R(P.qp_mu_target)));
        } else {
            solve_with(false, R(0));
        }
        const R alpha = nsides > 0 ? std::min(R(1), R(0.995) * max_step()) : R(1);
        last_step = 0;
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < st[k].nz; ++i) {
                const R step = alpha * R(dz[k][i]);
                z[k][i] += step;
                last_step = std::max(last_step, R(std::fabs(step)));
            }
            for (RowT<R>& r : st[k].rows) {
                if (r.eq) {
                    if (r.hard && r.rho > R(0)) r.lambda += r.rho * (row_value_t(r, z[k].data()) - r.lb);
                    continue;
                }
                for (int sd = 0; sd < 2; ++sd) {
                    if (!r.on[sd]) continue;
                    r.t[sd] += alpha * r.dt[sd];
                    r.lam[sd] += alpha * r.dl[sd];
                }
            }
        }
        last_alpha = alpha;
        if (!(last_step < std::numeric_limits<R>::infinity())) {   // non-finite step
            failed = it + 1;
            break;
        }
    }
    zout.assign(N + 1, {});
    for (int k = 0; k <= N; ++k) zout[k].assign(z[k].begin(), z[k].end());
    info[0] = iters; info[1] = converged; info[2] = failed; info[3] = double(mu); info[4] = double(rd_max);
}

#include "reduced_lab.h"

}  // namespace orc

// ---------------------------------------------------------------------------
extern "C" {

int oracle_dims(const ub_problem_desc_t* P, int32_t out[8]) {
    const orc::Dims D = orc::make_dims(*P);
    out[0] = D.nx; out[1] = D.nu; out[2] = D.neq; out[3] = D.nfric + D.nobs;
    out[4] = D.nterm; out[5] = D.N; out[6] = P->nb; out[7] = P->nc;
    return 0;
}

// Batched solve, `nthreads` host threads, one instance per task.
int oracle_solve_batch(const ub_problem_desc_t* P, int32_t B, const double* x0, const double* target,
                       const double* body_params, double* X, double* U, double* K, int32_t* status, double* stats,
                       uint32_t flags, int32_t nthreads) {
    const orc::Dims D = orc::make_dims(*P);
    const size_t nxs = size_t(D.N + 1) * D.nx, nus = size_t(D.N) * D.nu;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= B) return;
            const double* bp = body_params ? body_params + size_t(b) * P->nb * UB_BODY_PARAMS : &P->body_params[0][0];
            orc::solve_one(*P, x0 + size_t(b) * D.nx, target + size_t(b) * (D.N + 1) * orc::target_stride(*P), bp,
                           (flags & UB_WARM_START) != 0, X + b * nxs, U + b * nus,
                           K ? K + size_t(b) * D.N * D.nu * D.nx : nullptr, status + b,
                           stats ? stats + size_t(b) * UB_STATS : nullptr);
        }
    };
    nthreads = std::max(1, nthreads);
    if (nthreads == 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    return 0;
}

// FK probe: out = r(3), C_we(9 row-major), v(3), w(3), a(3), alpha(3), spheres(3*n_spheres)
int oracle_fk(const ub_problem_desc_t* P, const double* x, double* out) {
    const orc::Kinematics<double> K = orc::forward_kinematics<double>(*P, x);
    int o = 0;
    for (int i = 0; i < 3; ++i) out[o++] = K.r[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[o++] = K.C_we.m[i][j];
    for (int i = 0; i < 3; ++i) out[o++] = K.v[i];
    for (int i = 0; i < 3; ++i) out[o++] = K.w[i];
    for (int i = 0; i < 3; ++i) out[o++] = K.a[i];
    for (int i = 0; i < 3; ++i) out[o++] = K.al[i];
    for (int s = 0; s < P->n_spheres; ++s)
        for (int i = 0; i < 3; ++i) out[o++] = K.sphere[s][i];
    return o;
}

// Linearisation probe of one knot.  Any output pointer may be NULL.
//   g[neq], C[neq*nx], Df[neq*nfc], r[3], Jp[3*nq], hfric[nfric], Ffric[nfric*nfc], hobs[nobs], Jobs[nobs*nq]
int oracle_linearize(const ub_problem_desc_t* P, const double* x, const double* u, const double* body_params,
                     double* g, double* C, double* Df, double* r, double* Jp, double* hfric, double* Ffric,
                     double* hobs, double* Jobs) {
    const orc::Dims D = orc::make_dims(*P);
    orc::KnotLin L;
    orc::linearize_knot(*P, D, body_params ? body_params : &P->body_params[0][0], x, u, L);
    auto cp = [](double* dst, const std::vector<double>& src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), sizeof(double) * src.size());
    };
    cp(g, L.g); cp(C, L.C.d); cp(Df, L.Df.d); cp(Jp, L.Jp.d);
    cp(hfric, L.hfric); cp(Ffric, L.Ffric.d); cp(hobs, L.hobs); cp(Jobs, L.Jobs.d);
    if (r) std::memcpy(r, L.r, sizeof(L.r));
    return 0;
}

// Projectile-path rows of one knot: h[nproj], J[nproj*nx], tclose[nproj] (any may be NULL); returns nproj
// End-effector orientation error against the target quaternion qref [x y z w] and its Jacobian over q (3 x nq).
int oracle_orientation_error(const ub_problem_desc_t* P, const double* x, const double* qref, double* e, double* J) {
    const orc::Dims D = orc::make_dims(*P);
    orc::KnotLin L;
    std::vector<double> u(D.nu, 0.0);
    orc::linearize_knot(*P, D, &P->body_params[0][0], x, u.data(), L, qref);
    for (int i = 0; i < 3; ++i) {
        e[i] = L.eo[i];
        for (int j = 0; j < D.nq; ++j) J[i * D.nq + j] = L.Jo(i, j);
    }
    return 0;
}

int oracle_projectile(const ub_problem_desc_t* P, const double* x, double* h, double* J, double* tclose) {
    const orc::Dims D = orc::make_dims(*P);
    if (D.nproj == 0) return 0;
    std::vector<double> u(D.nu, 0.0);
    orc::KnotLin L;
    orc::linearize_knot(*P, D, &P->body_params[0][0], x, u.data(), L);
    if (h) std::memcpy(h, L.hproj.data(), sizeof(double) * D.nproj);
    if (J) std::memcpy(J, L.Jproj.d.data(), sizeof(double) * L.Jproj.d.size());
    if (tclose) {
        const orc::Kinematics<double> K = orc::forward_kinematics<double>(*P, x);
        std::vector<double> hv(D.nproj);
        orc::projectile_constraints<double>(*P, K, x + D.nx - 9, hv.data(), tclose);
    }
    return D.nproj;
}

// Performance index of a trajectory: out = {cost, dyn_sse, eq_sse, ineq_sse, violation, max_eq, min_margin}
int oracle_performance(const ub_problem_desc_t* P, const double* target, const double* body_params, const double* X,
                       const double* U, double* out) {
    const orc::Dims D = orc::make_dims(*P);
    orc::Mat A, B;
    orc::discrete_dynamics(*P, D, A, B);
    const orc::Perf pf =
        orc::performance(*P, D, A, B, body_params ? body_params : &P->body_params[0][0], target, X, U);
    out[0] = pf.cost; out[1] = pf.dyn_sse; out[2] = pf.eq_sse; out[3] = pf.ineq_sse;
    out[4] = pf.violation(); out[5] = pf.max_eq; out[6] = pf.min_margin;
    return 0;
}

// Dump the QP of the first SQP iteration around (X,U) for an independent check.
// Per stage k: H (nz*nz), g (nz), b (nx), then rows as (a[nz], c, lb, ub, rho, hard).
// Returns the number of doubles written (or needed, if out == NULL).
int64_t oracle_qp_dump(const ub_problem_desc_t* P, const double* target, const double* body_params, const double* X,
                       const double* U, double* out, int32_t* rows_per_stage) {
    orc::Workspace W;
    W.D = orc::make_dims(*P);
    orc::discrete_dynamics(*P, W.D, W.A, W.B);
    orc::build_qp(*P, W, body_params ? body_params : &P->body_params[0][0], target, X, U);
    int64_t n = 0;
    auto put = [&](double v) {
        if (out) out[n] = v;
        ++n;
    };
    for (int k = 0; k <= W.D.N; ++k) {
        const orc::Stage& s = W.st[k];
        if (rows_per_stage) rows_per_stage[k] = int32_t(s.rows.size());
        for (double v : s.H.d) put(v);
        for (double v : s.g) put(v);
        for (int i = 0; i < W.D.nx; ++i) put(k < W.D.N ? s.b[i] : 0.0);
        for (const orc::Row& r : s.rows) {
            for (int j = 0; j < s.nz; ++j) put(r.idx >= 0 ? (j == r.idx ? 1.0 : 0.0) : r.a[j]);
            put(r.c); put(r.lb); put(r.ub); put(r.rho); put(r.hard ? 1.0 : 0.0);
        }
    }
    return n;
}

// QP step only (first SQP iteration around X,U): dX [N+1,nx], dU [N,nu]; returns Newton iterations.
int oracle_qp_step(const ub_problem_desc_t* P, const double* target, const double* body_params, const double* X,
                   const double* U, double* dX, double* dU, double* info) {
    orc::Workspace W;
    W.D = orc::make_dims(*P);
    orc::discrete_dynamics(*P, W.D, W.A, W.B);
    orc::build_qp(*P, W, body_params ? body_params : &P->body_params[0][0], target, X, U);
    std::vector<std::vector<double>> z;
    const orc::QpResult qr = orc::solve_qp(*P, W, z, nullptr);
    for (int k = 0; k <= W.D.N; ++k) {
        const int xo = W.st[k].nu;
        for (int i = 0; i < W.D.nx; ++i) dX[size_t(k) * W.D.nx + i] = z[k][xo + i];
        if (k < W.D.N)
            for (int i = 0; i < W.D.nu; ++i) dU[size_t(k) * W.D.nu + i] = z[k][i];
    }
    if (info) {
        info[0] = qr.iters; info[1] = qr.converged; info[2] = qr.decrement; info[3] = qr.hard_infeas;
    }
    return qr.iters;
}


// Precision study entry (tools/precision_lab.py): QP step of the first SQP iteration around (X, U) with
// mode 0 = <double, double>, 1 = <float, float> (the fp32 kernels' arithmetic), 2 = <float factors, double iterate>.
// dX [N+1, nx], dU [N, nu]; info[5] as solve_qp_ipm_precision.
int oracle_qp_step_precision(const ub_problem_desc_t* P, const double* target, const double* body_params,
                             const double* X, const double* U, int32_t mode, double* dX, double* dU, double* info) {
    orc::Workspace W;
    W.D = orc::make_dims(*P);
    orc::discrete_dynamics(*P, W.D, W.A, W.B);
    orc::build_qp(*P, W, body_params ? body_params : &P->body_params[0][0], target, X, U);
    std::vector<std::vector<double>> z;
    // hundreds digit: acceptance threshold on the complementarity, mu <= f * mu_target with f = 2 (the specification),
    // 1.2 or 1.05; tens digit: Newton iterations run after the convergence test has passed
    const double factors[3] = {2.0, 1.2, 1.05};
    const double f = factors[std::min(2, mode / 100)];
    const int extra = (mode / 10) % 10;
    mode %= 10;
    if (mode == 0) orc::solve_qp_ipm_precision<double, double>(*P, W, z, info, extra, f);
    else if (mode == 1) orc::solve_qp_ipm_precision<float, float>(*P, W, z, info, extra, f);
    else if (mode == 2) orc::solve_qp_ipm_precision<float, double>(*P, W, z, info, extra, f);
    // reduced stage (forces eliminated in range-space form, Riccati on [jerk; state]): oracle/reduced_lab.h
    else if (mode == 3) orc::solve_qp_ipm_reduced<double, double, double, double>(*P, W, z, info, extra, f);
    else if (mode == 4) orc::solve_qp_ipm_reduced<float, double, double, float>(*P, W, z, info, extra, f);
    else if (mode == 5) orc::solve_qp_ipm_reduced<float, double, double, double>(*P, W, z, info, extra, f);
    else if (mode == 6) orc::solve_qp_ipm_reduced<float, double, float, float>(*P, W, z, info, extra, f);
    else orc::solve_qp_ipm_reduced<float, float, float, float>(*P, W, z, info, extra, f);
    for (int k = 0; k <= W.D.N; ++k) {
        const int xo = W.st[k].nu;
        for (int i = 0; i < W.D.nx; ++i) dX[size_t(k) * W.D.nx + i] = z[k][xo + i];
        if (k < W.D.N)
            for (int i = 0; i < W.D.nu; ++i) dU[size_t(k) * W.D.nu + i] = z[k][i];
    }
    return 0;
}

}  // extern "C"
