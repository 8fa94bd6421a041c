// upright_b200 — device-side problem description, small vector math and the
// forward kinematics (values and lane-parallel forward-mode tangents).
//
// Mapping: ONE WARP PER MPC INSTANCE.  In the linearisation, lane j of the
// warp carries d/dx_j of every kinematic quantity (nx = 3*nq <= 27 <= 32
// lanes), so a whole Jacobian column set is produced by one pass of the chain
// — the role CppAD's taped Jacobians play in the reference
// (upright_control/src/constraint/balancing_constraints.cpp:54-56,105-107).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/upright_b200.h"

namespace ub {

constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

template <typename T>
struct DevProblem {
    int nq, nx, nu, nfc, neq, nfric, nobs, nterm, nb, nc, nf, N, nsph, npairs;
    int nz, nbox_u, nrow, sqp_iters, qp_iter_max;
    int soft_u, soft_x, soft_poly, balancing;
    T dt, Z, invZ, rho_hard, mu0, thr0, mu_target, qp_tol, reg_input, eps_hard;
    T alpha_decay, alpha_min, g_max, g_min, gamma_c, armijo, delta_tol, cost_tol;
    int jtype[UB_MAX_JOINTS];
    T jR[UB_MAX_JOINTS][9], jp[UB_MAX_JOINTS][3], jaxis[UB_MAX_JOINTS][3];
    T toolR[9], toolp[3], grav[3];
    T Qd[UB_MAX_NX], Rd[UB_MAX_JOINTS], Wd[3], fw, xd[UB_MAX_NX];
    // orientation part of the end-effector weight (end_effector_cost.h:61-81); ori != 0: targets carry the desired
    // quaternion [x y z w] behind the position (7 columns)
    int ori;
    T Wo[3];
    T xlb[UB_MAX_NX], xub[UB_MAX_NX], ulb[UB_MAX_JOINTS], uub[UB_MAX_JOINTS], flb, fub;
    T body[UB_MAX_BODIES][UB_BODY_PARAMS];
    int cb1[UB_MAX_CONTACTS], cb2[UB_MAX_CONTACTS];
    T cmu[UB_MAX_CONTACTS], cr1[UB_MAX_CONTACTS][3], cr2[UB_MAX_CONTACTS][3], cn[UB_MAX_CONTACTS][3],
        cspan[UB_MAX_CONTACTS][6];
    int slink[UB_MAX_SPHERES], sshape[UB_MAX_SPHERES];
    T srad[UB_MAX_SPHERES], soff[UB_MAX_SPHERES][3];
    int pa[UB_MAX_PAIRS], pb[UB_MAX_PAIRS];
    T dmin;
    // end-effector box rows (end_effector_box_constraint.h): appended to the obstacle rows, nobs = npairs + 6
    int eebox;
    T eb_lo[3], eb_hi[3];
    // inertial-alignment Gauss-Newton cost (inertial_alignment.cpp:151-163): e = S C' (a - g) / |g|
    int iacost;
    T ia_w, ia_S[6], ia_inv_g;
    // inertial-alignment constraint (inertial_alignment.cpp:7-53): five rows behind the obstacle / box rows
    int iacon, ia_use_ang, ia_fixed, obsw;
    T ia_alpha, ia_n[3], ia_com[3];
    // dynamic obstacles: ndyn x [p, v, a] appended to the state (nxo = 9 ndyn); spheres with slink == -2 - j ride on j
    int ndyn, nxo;
    // projectile-path rows (projectile_path_constraint.h:46-156): one per listed collision sphere, behind the
    // inertial-alignment rows; proj_s = the target-state flag s
    int nproj, proj_sph[UB_MAX_PROJECTILE_LINKS];
    T proj_d[UB_MAX_PROJECTILE_LINKS], proj_scale, proj_s;
    // force block of the reduced stage (ub_solver.cuh): equality rows fall into ngrp groups of ng rows — one group
    // of 6 per body when every contact is with the tray (body1 == -1), ONE group of 6 nb when bodies share contacts;
    // bc_list[bc_start[b] .. bc_start[b+1]) = the contacts of body b as 2 * contact + side (0: it is body2, 1: body1)
    int ngrp, ng;
    int bc_start[UB_MAX_BODIES + 1], bc_list[2 * UB_MAX_CONTACTS];
};

// Division in the interior-point row updates (eight per inequality row and pass): the fp32 kernels use the
// approximate reciprocal path (2 ulp, two instructions instead of ~nine); the Newton residuals themselves are
// formed with exact operations, so this only perturbs the step at rounding level.
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return a / b; }
// Reciprocal of a double in the product kernels' residual arithmetic: float seed + one Newton step (relative error
// ~4e-15, six instructions instead of the ~thirty of an IEEE division); exact outside the float range.
static __device__ __noinline__ double slow_rcp(double d) { return 1.0 / d; }   // out of line: keeps the hot loops small
__device__ __forceinline__ double fast_rcp(double d) {
    if (!(fabs(d) > 1e-30) || !(fabs(d) < 1e30)) return slow_rcp(d);
    const double r = double(__frcp_rn(float(d)));
    return r * (2.0 - d * r);
}

// ------------------------------------------------------------------ vectors
template <typename T>
struct V3 {
    T x, y, z;
    __device__ __forceinline__ V3() : x(0), y(0), z(0) {}
    __device__ __forceinline__ V3(T a, T b, T c) : x(a), y(b), z(c) {}
    __device__ __forceinline__ T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T>
__device__ __forceinline__ V3<T> operator+(const V3<T>& a, const V3<T>& b) { return V3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T>
__device__ __forceinline__ V3<T> operator-(const V3<T>& a, const V3<T>& b) { return V3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T>
__device__ __forceinline__ V3<T> operator*(T s, const V3<T>& a) { return V3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T>
__device__ __forceinline__ T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T>
__device__ __forceinline__ V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return V3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T>
__device__ __forceinline__ V3<T> ld3(const T* p) { return V3<T>(p[0], p[1], p[2]); }

// Separation of collision pair (a, b) given the centres `sph` (3 per object), without the minimum distance:
// sphere-sphere |c_a - c_b| - r_a - r_b, or sphere against the half-space {p : n.p <= d} (the reference's `ground`,
// controller_interface.cpp:93-101): n.c_a - d - r_a.  *dir: d h / d c_a (= -d h / d c_b between spheres).
template <typename T, typename PT>
__device__ __forceinline__ T pair_separation(const DevProblem<PT>& P, int a, int b, const T* sph, V3<T>* dir) {
    if (P.sshape[b] == UB_SHAPE_HALFSPACE || P.sshape[a] == UB_SHAPE_HALFSPACE) {
        const int h = P.sshape[b] == UB_SHAPE_HALFSPACE ? b : a, s = h == b ? a : b;
        const V3<T> n(T(P.soff[h][0]), T(P.soff[h][1]), T(P.soff[h][2]));
        *dir = h == b ? n : T(-1) * n;
        return n.x * sph[3 * s] + n.y * sph[3 * s + 1] + n.z * sph[3 * s + 2] - T(P.srad[h]) - T(P.srad[s]);
    }
    const V3<T> d(sph[3 * a] - sph[3 * b], sph[3 * a + 1] - sph[3 * b + 1], sph[3 * a + 2] - sph[3 * b + 2]);
    const T dist = sqrt(dot(d, d));
    *dir = (T(1) / dist) * d;
    return dist - T(P.srad[a]) - T(P.srad[b]);
}

// cubic_newtons + projectile_closest_time (constraint/projectile_path_constraint.h:11-44): time at which the
// ballistic path r0 + t v0 + t^2 g / 2 is nearest to r; Newton from t = 0, at most 10 steps, step tolerance 1e-4
template <typename T>
__device__ inline T projectile_closest_time(const V3<T>& r, const V3<T>& r0, const V3<T>& v0, const V3<T>& g) {
    const V3<T> dr = r - r0;
    const T a = dot(g, g), b = T(3) * dot(v0, g), c = T(2) * (dot(v0, v0) - dot(dr, g)), d = T(-2) * dot(dr, v0);
    T t = T(0);
    for (int it = 0; it < 10; ++it) {
        const T f = ((a * t + b) * t + c) * t + d;
        const T df = (T(3) * a * t + T(2) * b) * t + c;
        const T update = f / df;
        t -= update;
        if (fabs(update) < T(1e-4)) break;
    }
    return t;
}
// Row i of ProjectilePathConstraint (projectile_path_constraint.h:77-146) for the collision-sphere centre c and the
// projectile state xo = [p, v, a]: value (scale / d_i) s (|c - r_closest| - d_i); *n = unit vector from the closest
// point of the path to c, *tc = time of closest approach (clamped at 0; 0 while s <= 0.5)
template <typename T>
__device__ inline T projectile_row(const DevProblem<T>& P, int i, const V3<T>& c, const T* xo, V3<T>* n, T* tc) {
    const V3<T> p0 = ld3(xo), v0 = ld3(xo + 3), a0 = ld3(xo + 6);
    T t = T(0);
    if (P.proj_s > T(0.5)) t = max(T(0), projectile_closest_time(c, p0, v0, a0));
    const V3<T> delta = c - (p0 + t * v0 + (T(0.5) * t * t) * a0);
    const T dist = sqrt(dot(delta, delta));
    *n = (T(1) / dist) * delta;
    *tc = t;
    return P.proj_scale / P.proj_d[i] * P.proj_s * (dist - P.proj_d[i]);
}

template <typename T>
struct M3 {
    T m[9];
    __device__ __forceinline__ V3<T> mul(const V3<T>& v) const {
        return V3<T>(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
                     m[6] * v.x + m[7] * v.y + m[8] * v.z);
    }
    __device__ __forceinline__ V3<T> tmul(const V3<T>& v) const {  // R^T v
        return V3<T>(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
                     m[2] * v.x + m[5] * v.y + m[8] * v.z);
    }
};
template <typename T>
__device__ __forceinline__ M3<T> matmul(const M3<T>& A, const T* B) {
    M3<T> C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B[j] + A.m[3 * i + 1] * B[3 + j] + A.m[3 * i + 2] * B[6 + j];
    return C;
}
// Quaternion [x y z w] of a rotation matrix by Eigen's conversion (the branch on the trace decides its sign), and
// the end-effector orientation error of ocs2::quaternionDistance against the desired quaternion r:
//     e = q_w r_v - r_w q_v + q_v x r_v      (end_effector_cost.h:61-67 [EXT: ocs2_robotic_tools])
// de: its derivative along a world-frame angular tangent dth of the rotation (q' = 1/2 [dth, 0] * q).
template <typename T>
__device__ inline void matrix_to_quaternion(const M3<T>& R, T q[4]) {
    const T tr = R.m[0] + R.m[4] + R.m[8];
    if (tr > T(0)) {
        T t = sqrt(tr + T(1));
        q[3] = T(0.5) * t;
        t = T(0.5) / t;
        q[0] = (R.m[7] - R.m[5]) * t;
        q[1] = (R.m[2] - R.m[6]) * t;
        q[2] = (R.m[3] - R.m[1]) * t;
    } else {
        int i = 0;
        if (R.m[4] > R.m[0]) i = 1;
        if (R.m[8] > R.m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        T t = sqrt(R.m[4 * i] - R.m[4 * j] - R.m[4 * k] + T(1));
        T qq[3];
        qq[i] = T(0.5) * t;
        t = T(0.5) / t;
        q[3] = (R.m[3 * k + j] - R.m[3 * j + k]) * t;
        qq[j] = (R.m[3 * j + i] + R.m[3 * i + j]) * t;
        qq[k] = (R.m[3 * k + i] + R.m[3 * i + k]) * t;
        q[0] = qq[0];
        q[1] = qq[1];
        q[2] = qq[2];
    }
}
template <typename T, bool TANGENT>
__device__ inline V3<T> orientation_error(const M3<T>& Cwe, const V3<T>& dth, const T* qref, V3<T>* de) {
    T q[4];
    matrix_to_quaternion(Cwe, q);
    const V3<T> qv(q[0], q[1], q[2]), rv(qref[0], qref[1], qref[2]);
    const T qw = q[3], rw = qref[3];
    if (TANGENT) {
        const T dqw = T(-0.5) * dot(dth, qv);
        const V3<T> dqv = T(0.5) * (qw * dth + cross(dth, qv));
        *de = dqw * rv - rw * dqv + cross(dqv, rv);
    }
    return qw * rv - rw * qv + cross(qv, rv);
}

template <typename T>
__device__ __forceinline__ void sincos_t(T a, T* s, T* c);
template <>
__device__ __forceinline__ void sincos_t<float>(float a, float* s, float* c) { sincosf(a, s, c); }
template <>
__device__ __forceinline__ void sincos_t<double>(double a, double* s, double* c) { sincos(a, s, c); }

template <typename T>
__device__ __forceinline__ void axis_angle_sc(const V3<T>& u, T s, T c, T* R);
template <typename T>
__device__ __forceinline__ void axis_angle(const V3<T>& u, T th, T* R) {
    T s, c;
    sincos_t<T>(th, &s, &c);
    axis_angle_sc(u, s, c, R);
}
template <typename T>
__device__ __forceinline__ void axis_angle_sc(const V3<T>& u, T s, T c, T* R) {
    const T t = T(1) - c;
    R[0] = c + t * u.x * u.x;
    R[1] = t * u.x * u.y - s * u.z;
    R[2] = t * u.x * u.z + s * u.y;
    R[3] = t * u.x * u.y + s * u.z;
    R[4] = c + t * u.y * u.y;
    R[5] = t * u.y * u.z - s * u.x;
    R[6] = t * u.x * u.z - s * u.y;
    R[7] = t * u.y * u.z + s * u.x;
    R[8] = c + t * u.z * u.z;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// --------------------------------------------------------------- kinematics
// Tool-frame state: upright::RigidBodyState (upright_core/include/upright_core/types.h:72-85).
template <typename T>
struct Kin {
    V3<T> r, v, w, a, al;
    M3<T> C;
};
// Tangent of the above along one direction; rotation tangents are world-frame
// angular vectors dth with dC = skew(dth) C.
template <typename T>
struct KinTan {
    V3<T> r, v, w, a, al, th;
};

// e (2) and, if TANGENT, its derivative along this lane's direction
template <typename T, bool TANGENT>
__device__ __forceinline__ void inertial_alignment_error(const DevProblem<T>& P, const Kin<T>& K, const KinTan<T>& D, T* e, T* de) {
    const V3<T> t = K.a - ld3(P.grav);
    const V3<T> ce = K.C.tmul(t);
    e[0] = dot(ld3(P.ia_S), ce) * P.ia_inv_g;
    e[1] = dot(ld3(P.ia_S + 3), ce) * P.ia_inv_g;
    if (TANGENT) {
        const V3<T> dce = K.C.tmul(D.a - cross(D.th, t));
        de[0] = dot(ld3(P.ia_S), dce) * P.ia_inv_g;
        de[1] = dot(ld3(P.ia_S + 3), dce) * P.ia_inv_g;
    }
}

// InertialAlignmentConstraint::constraintFunction (inertial_alignment.cpp:7-53), transcribed literally:
//   a = C' (acc - g);  use_angular_acceleration: a += ddC_we com;  align_with_fixed_vector: a = C' n
//   h = [a_n, alpha a_n -+ a_t0 -+ a_t1]   with a_n = n . a, a_t = S a
template <typename T, bool TANGENT>
__device__ __forceinline__ void inertial_alignment_rows(const DevProblem<T>& P, const Kin<T>& K, const KinTan<T>& D, T* h, T* dh) {
    const V3<T> n = ld3(P.ia_n), s0 = ld3(P.ia_S), s1 = ld3(P.ia_S + 3);
    const V3<T> t = K.a - ld3(P.grav);
    V3<T> a = K.C.tmul(t), da;
    if (TANGENT) da = K.C.tmul(D.a - cross(D.th, t));
    if (P.ia_use_ang) {
        const V3<T> cw = K.C.mul(ld3(P.ia_com));
        const V3<T> wc = cross(K.w, cw);
        a = a + cross(K.al, cw) + cross(K.w, wc);   // (S(alpha) + S(w) S(w)) C com
        if (TANGENT) {
            const V3<T> dcw = cross(D.th, cw);
            da = da + cross(D.al, cw) + cross(K.al, dcw) + cross(D.w, wc) + cross(K.w, cross(D.w, cw) + cross(K.w, dcw));
        }
    } else if (P.ia_fixed) {
        a = K.C.tmul(n);
        if (TANGENT) da = K.C.tmul(T(-1) * cross(D.th, n));
    }
    const T an = dot(n, a), t0 = dot(s0, a), t1 = dot(s1, a), al = P.ia_alpha;
    h[0] = an;
    h[1] = al * an - t0 - t1;
    h[2] = al * an - t0 + t1;
    h[3] = al * an + t0 - t1;
    h[4] = al * an + t0 + t1;
    if (TANGENT) {
        const T dn = dot(n, da), d0 = dot(s0, da), d1 = dot(s1, da);
        dh[0] = dn;
        dh[1] = al * dn - d0 - d1;
        dh[2] = al * dn - d0 + d1;
        dh[3] = al * dn + d0 - d1;
        dh[4] = al * dn + d0 + d1;
    }
}

// Forward kinematics with classical accelerations (what the reference obtains
// from ocs2::PinocchioEndEffectorKinematicsCppAd,
// upright_control/src/constraint/balancing_constraints.cpp:15-30) on the
// chain of upright_control/include/upright_control/util.h:15-65.
// If TANGENT, also propagates d/dx_dir (dir = this lane's direction; dir >= nx
// propagates zeros).  Sphere centres (and tangents) are written to sph / dsph
// (3 values per sphere, registers of this lane) when non-null.
// `sc`: optional [sin q_i, cos q_i] pairs computed beforehand (the linearisation shares them across the lanes).
template <typename T, bool TANGENT, typename XT = T>
__device__ void forward_kinematics(const DevProblem<T>& P, const XT* __restrict__ x, int dir, Kin<T>& K, KinTan<T>& D,
                                   T* sph, T* dsph, const T* sc = nullptr) {
    const int nq = P.nq;
    M3<T> R;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0) ? T(1) : T(0);
    V3<T> p, v, w, a, al;
    V3<T> dp, dv, dw, da, dal, dth;

    auto offset_frame = [&](const T* Rt, const T* pt) {
        const V3<T> ro = R.mul(ld3(pt));
        const V3<T> wro = cross(w, ro);
        if (TANGENT) {
            const V3<T> dro = cross(dth, ro);
            dp = dp + dro;
            dv = dv + cross(dw, ro) + cross(w, dro);
            da = da + cross(dal, ro) + cross(al, dro) + cross(dw, wro) + cross(w, cross(dw, ro) + cross(w, dro));
        }
        p = p + ro;
        v = v + wro;
        a = a + cross(al, ro) + cross(w, wro);
        R = matmul(R, Rt);
    };
    auto attach = [&](int link) {
        if (sph == nullptr) return;
        for (int s = 0; s < P.nsph; ++s) {
            if (P.slink[s] != link) continue;
            const V3<T> o = R.mul(ld3(P.soff[s]));
            sph[3 * s] = p.x + o.x;
            sph[3 * s + 1] = p.y + o.y;
            sph[3 * s + 2] = p.z + o.z;
            if (TANGENT) {
                const V3<T> d = dp + cross(dth, o);
                dsph[3 * s] = d.x;
                dsph[3 * s + 1] = d.y;
                dsph[3 * s + 2] = d.z;
            }
        }
    };
    if (sph != nullptr)
        for (int s = 0; s < P.nsph; ++s)
            if (P.slink[s] < 0) {
                sph[3 * s] = P.soff[s][0];
                sph[3 * s + 1] = P.soff[s][1];
                sph[3 * s + 2] = P.soff[s][2];
                if (TANGENT) dsph[3 * s] = dsph[3 * s + 1] = dsph[3 * s + 2] = T(0);
            }

    for (int i = 0; i < nq; ++i) {
        offset_frame(P.jR[i], P.jp[i]);
        const V3<T> ul = ld3(P.jaxis[i]);
        const V3<T> z = R.mul(ul);
        const T qi = T(x[i]), qd = T(x[nq + i]), qdd = T(x[2 * nq + i]);   // the iterate may be stored narrower than T
        const T dq = (TANGENT && dir == i) ? T(1) : T(0);
        const T dqd = (TANGENT && dir == nq + i) ? T(1) : T(0);
        const T dqdd = (TANGENT && dir == 2 * nq + i) ? T(1) : T(0);
        if (P.jtype[i] == UB_JOINT_REVOLUTE) {
            const V3<T> wz = cross(w, z);
            if (TANGENT) {
                const V3<T> dz = cross(dth, z);
                dal = dal + dqdd * z + qdd * dz + dqd * wz + qd * (cross(dw, z) + cross(w, dz));
                dw = dw + dqd * z + qd * dz;
                dth = dth + dq * z;
            }
            al = al + qdd * z + qd * wz;
            w = w + qd * z;
            T Rq[9];
            if (sc != nullptr) axis_angle_sc(ul, sc[2 * i], sc[2 * i + 1], Rq);
            else axis_angle(ul, qi, Rq);
            R = matmul(R, Rq);
        } else {
            const V3<T> d = qi * z;
            const V3<T> wz = cross(w, z);
            const V3<T> wd = cross(w, d);
            if (TANGENT) {
                const V3<T> dz = cross(dth, z);
                const V3<T> dd = dq * z + qi * dz;
                da = da + dqdd * z + qdd * dz + T(2) * (dqd * wz + qd * (cross(dw, z) + cross(w, dz))) + cross(dal, d) +
                     cross(al, dd) + cross(dw, wd) + cross(w, cross(dw, d) + cross(w, dd));
                dv = dv + dqd * z + qd * dz + cross(dw, d) + cross(w, dd);
                dp = dp + dd;
            }
            a = a + qdd * z + T(2) * qd * wz + cross(al, d) + cross(w, wd);
            v = v + qd * z + wd;
            p = p + d;
        }
        attach(i);
    }
    offset_frame(P.toolR, P.toolp);
    attach(nq);
    K.r = p;
    K.v = v;
    K.w = w;
    K.a = a;
    K.al = al;
    K.C = R;
    if (TANGENT) {
        D.r = dp;
        D.v = dv;
        D.w = dw;
        D.a = da;
        D.al = dal;
        D.th = dth;
    }
}

// Rigid body inertial parameters unpacked from [m, m*com, vech(I)]
// (upright_core/include/upright_core/rigid_body.h:36-46).
template <typename T>
struct BodyP {
    T m;
    V3<T> com;
    T I[6];  // xx xy xz yy yz zz
    __device__ __forceinline__ V3<T> Imul(const V3<T>& v) const {
        return V3<T>(I[0] * v.x + I[1] * v.y + I[2] * v.z, I[1] * v.x + I[3] * v.y + I[4] * v.z,
                     I[2] * v.x + I[4] * v.y + I[5] * v.z);
    }
};
template <typename T, typename S>
__device__ __forceinline__ BodyP<T> load_body(const S* p) {   // S: storage type of the parameters
    BodyP<T> b;
    b.m = T(p[0]);
    const T inv = T(1) / T(p[0]);
    b.com = V3<T>(T(p[1]) * inv, T(p[2]) * inv, T(p[3]) * inv);
#pragma unroll
    for (int i = 0; i < 6; ++i) b.I[i] = T(p[4 + i]);
    return b;
}

// State-dependent part of the object-dynamics rows of one body and (if
// TANGENT) its derivative along this lane's direction:
//   force  rows = scale * C_ew (a + ddC com - g)          (f_sum/m added via Df)
//   torque rows = scale * (w_e x I w_e + I al_e) / m
// upright_core/include/upright_core/contact_constraints.h:79-102 with
// dC_dtt of util.h:37-50 and the 1/sqrt(6 nb) of balancing_constraints.cpp:140-151.
template <typename T, bool TANGENT>
__device__ __forceinline__ void object_dynamics_state_part(const DevProblem<T>& P, const BodyP<T>& B, const Kin<T>& K,
                                                           const KinTan<T>& D, T scale, T* g6, T* dg6) {
    const V3<T> grav = ld3(P.grav);
    const V3<T> c = K.C.mul(B.com);
    const V3<T> wc = cross(K.w, c);
    const V3<T> t1 = K.a + cross(K.al, c) + cross(K.w, wc) - grav;
    const V3<T> gi = K.C.tmul(t1);
    const V3<T> we = K.C.tmul(K.w), ale = K.C.tmul(K.al);
    const V3<T> Iwe = B.Imul(we);
    const V3<T> tau = cross(we, Iwe) + B.Imul(ale);
    const T sm = scale / B.m;
    g6[0] = scale * gi.x;
    g6[1] = scale * gi.y;
    g6[2] = scale * gi.z;
    g6[3] = sm * tau.x;
    g6[4] = sm * tau.y;
    g6[5] = sm * tau.z;
    if (TANGENT) {
        const V3<T> dc = cross(D.th, c);
        const V3<T> dt1 = D.a + cross(D.al, c) + cross(K.al, dc) + cross(D.w, wc) + cross(K.w, cross(D.w, c) + cross(K.w, dc));
        const V3<T> dgi = K.C.tmul(dt1 - cross(D.th, t1));
        const V3<T> dwe = K.C.tmul(D.w - cross(D.th, K.w));
        const V3<T> dale = K.C.tmul(D.al - cross(D.th, K.al));
        const V3<T> dtau = cross(dwe, Iwe) + cross(we, B.Imul(dwe)) + B.Imul(dale);
        dg6[0] = scale * dgi.x;
        dg6[1] = scale * dgi.y;
        dg6[2] = scale * dgi.z;
        dg6[3] = sm * dtau.x;
        dg6[4] = sm * dtau.y;
        dg6[5] = sm * dtau.z;
    }
}

}  // namespace ub
