// TEST INFRASTRUCTURE — CPU oracle (fp64) for the upright MPC hot path.
//
// Scalar-templated restatement of the problem functions the reference feeds
// to OCS2.  Each function cites the reference lines it follows.  Instantiated
// with `double` for values and with `orc::Dual` for exact first derivatives
// (the role CppAD plays in the reference).
#pragma once
#include <vector>

#include "../include/upright_b200.h"
#include "smallmath.h"

namespace orc {

// upright::RigidBodyState (upright_core/include/upright_core/types.h:72-85) +
// the collision-sphere centres the obstacle constraint needs.
template <typename S>
struct Kinematics {
    Vec3<S> r;       // EE (tray) origin, world
    Mat3<S> C_we;    // EE orientation
    Vec3<S> v, w;    // linear / angular velocity, world-aligned
    Vec3<S> a, al;   // classical linear / angular acceleration, world-aligned
    Vec3<S> sphere[UB_MAX_SPHERES];
};

// Forward kinematics with velocities and classical accelerations of the tool
// frame.  Replaces ocs2::PinocchioEndEffectorKinematicsCppAd as used at
// upright_control/src/constraint/balancing_constraints.cpp:15-30
// (getPosition/Orientation/Velocity/AngularVelocity/Acceleration/
// AngularAcceleration, LOCAL_WORLD_ALIGNED) on the chain built at
// upright_control/include/upright_control/util.h:15-65; state slicing
// x -> (q, v, a) follows dynamics/system_pinocchio_mapping.h:11-57.
template <typename S>
Kinematics<S> forward_kinematics(const ub_problem_desc_t& P, const S* x) {
    const int nq = P.nq;
    Mat3<S> R = Mat3<S>::identity();
    Vec3<S> p, v, w, a, al;
    Kinematics<S> K;
    auto attach = [&](int link) {
        for (int s = 0; s < P.n_spheres; ++s)
            if (P.spheres[s].link == link) K.sphere[s] = p + R * vec3_from<S>(P.spheres[s].offset);
    };
    for (int s = 0; s < P.n_spheres; ++s) {
        if (P.spheres[s].link == -1) K.sphere[s] = vec3_from<S>(P.spheres[s].offset);
        // sphere riding on dynamic obstacle j (link == -2 - j): centre = position block of that obstacle's state,
        // x = [x_robot (3 nq), x_obs_0 (9), ...] (dimensions.h:36-41, obstacle_constraint.h:8-27)
        if (P.spheres[s].link <= -2) {
            const S* po = x + 3 * P.nq + 9 * (-2 - P.spheres[s].link);
            K.sphere[s] = Vec3<S>(po[0], po[1], po[2]);
        }
    }

    auto offset_frame = [&](const double* Rt, const double* pt) {
        // move the frame origin by a body-fixed offset and re-orient it
        const Vec3<S> ro = R * vec3_from<S>(pt);
        p = p + ro;
        v = v + cross(w, ro);
        a = a + cross(al, ro) + cross(w, cross(w, ro));
        R = R * Mat3<S>::from_rowmajor(Rt);
    };
    for (int i = 0; i < nq; ++i) {
        const ub_joint_t& J = P.joints[i];
        offset_frame(J.R, J.p);
        const Vec3<S> ul = vec3_from<S>(J.axis);
        const Vec3<S> z = R * ul;
        const S qi = x[i], qd = x[nq + i], qdd = x[2 * nq + i];
        if (J.type == UB_JOINT_REVOLUTE) {
            al = al + qdd * z + qd * cross(w, z);
            w = w + qd * z;
            R = R * axis_angle(ul, qi);
        } else {
            const Vec3<S> d = qi * z;
            a = a + qdd * z + S(2.0) * qd * cross(w, z) + cross(al, d) + cross(w, cross(w, d));
            v = v + qd * z + cross(w, d);
            p = p + d;
        }
        attach(i);
    }
    offset_frame(P.tool_R, P.tool_p);
    attach(nq);
    K.r = p;
    K.C_we = R;
    K.v = v;
    K.w = w;
    K.a = a;
    K.al = al;
    return K;
}

// upright::RigidBody::from_parameters (upright_core/include/upright_core/rigid_body.h:36-46)
template <typename S>
struct Body {
    S mass;
    Vec3<S> com;
    Mat3<S> inertia;
    static Body from_parameters(const double* p) {
        Body b;
        b.mass = S(p[0]);
        b.com = Vec3<S>(S(p[1] / p[0]), S(p[2] / p[0]), S(p[3] / p[0]));
        const double I[9] = {p[4], p[5], p[6], p[5], p[7], p[8], p[6], p[8], p[9]};
        b.inertia = Mat3<S>::from_rowmajor(I);
        return b;
    }
};

template <typename S>
struct Wrench {
    Vec3<S> force, torque;
};

// compute_object_wrenches (upright_core/include/upright_core/contact_constraints.h:106-157):
// contact force f acts on object1 with lever r_co_o1 - com1 and as -f on
// object2 with lever r_co_o2 - com2; frictionless => f = f_i * normal (:111-120).
template <typename S>
void object_wrenches(const ub_problem_desc_t& P, const Body<S>* bodies, const S* forces, Wrench<S>* out) {
    for (int b = 0; b < P.nb; ++b) out[b] = Wrench<S>();
    for (int i = 0; i < P.nc; ++i) {
        const ub_contact_t& c = P.contacts[i];
        Vec3<S> f;
        if (P.nf == 1) {
            f = forces[i] * vec3_from<S>(c.normal);
        } else {
            f = Vec3<S>(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);
        }
        if (c.body1 >= 0) {
            const Vec3<S> lever = vec3_from<S>(c.r_co_o1) - bodies[c.body1].com;
            out[c.body1].force = out[c.body1].force + f;
            out[c.body1].torque = out[c.body1].torque + cross(lever, f);
        }
        {
            const Vec3<S> lever = vec3_from<S>(c.r_co_o2) - bodies[c.body2].com;
            out[c.body2].force = out[c.body2].force - f;
            out[c.body2].torque = out[c.body2].torque + cross(lever, -f);
        }
    }
}

// compute_object_dynamics_constraints (contact_constraints.h:79-102,161-194)
// scaled by 1/sqrt(6 nb) as ObjectDynamicsConstraints::constraintFunction does
// (upright_control/src/constraint/balancing_constraints.cpp:140-151).
// dC_dtt = (S(alpha) + S(omega) S(omega)) C_we  (upright_core/include/upright_core/util.h:37-50).
template <typename S>
void object_dynamics_constraints(const ub_problem_desc_t& P, const double* body_params, const Kinematics<S>& X,
                                 const S* forces, S* g) {
    Body<S> bodies[UB_MAX_BODIES];
    for (int b = 0; b < P.nb; ++b) bodies[b] = Body<S>::from_parameters(body_params + UB_BODY_PARAMS * b);
    Wrench<S> wr[UB_MAX_BODIES];
    object_wrenches(P, bodies, forces, wr);

    const Vec3<S> gravity = vec3_from<S>(P.gravity);
    const Mat3<S> C_ew = X.C_we.transpose();
    const Mat3<S> Sw = skew3(X.w);
    const Mat3<S> ddC = (skew3(X.al) + Sw * Sw) * X.C_we;
    const S scale = S(1.0 / std::sqrt(6.0 * P.nb));
    for (int b = 0; b < P.nb; ++b) {
        const Body<S>& body = bodies[b];
        const Vec3<S> gi_force = body.mass * (C_ew * (X.a + ddC * body.com - gravity));
        const Vec3<S> w_e = C_ew * X.w;
        const Vec3<S> al_e = C_ew * X.al;
        const Vec3<S> inertial_torque = cross(w_e, body.inertia * w_e) + body.inertia * al_e;
        const Vec3<S> cf = (gi_force - wr[b].force);
        const Vec3<S> ct = (inertial_torque - wr[b].torque);
        for (int i = 0; i < 3; ++i) {
            g[6 * b + i] = scale * cf[i] / body.mass;
            g[6 * b + 3 + i] = scale * ct[i] / body.mass;
        }
    }
}

// compute_contact_force_constraints_linearized (contact_constraints.h:49-77): h >= 0
template <typename S>
void contact_force_constraints(const ub_problem_desc_t& P, const S* forces, S* h) {
    for (int i = 0; i < P.nc; ++i) {
        const ub_contact_t& c = P.contacts[i];
        const Vec3<S> f(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);
        const S fn = dot(vec3_from<S>(c.normal), f);
        const S ft0 = dot(vec3_from<S>(c.span), f);
        const S ft1 = dot(vec3_from<S>(c.span + 3), f);
        const S mu(c.mu);
        h[5 * i + 0] = fn;
        h[5 * i + 1] = mu * fn - ft0 - ft1;
        h[5 * i + 2] = mu * fn - ft0 + ft1;
        h[5 * i + 3] = mu * fn + ft0 - ft1;
        h[5 * i + 4] = mu * fn + ft0 + ft1;
    }
}

// InertialAlignmentCostGaussNewton::function (upright_control/src/inertial_alignment.cpp:151-163):
// e = S C_we' (a - g) / |g|, S = contact_plane_span (2 x 3)
template <typename S>
void inertial_alignment_error(const ub_problem_desc_t& P, const Kinematics<S>& X, S* e) {
    const Vec3<S> gravity = vec3_from<S>(P.gravity);
    const double gn = std::sqrt(P.gravity[0] * P.gravity[0] + P.gravity[1] * P.gravity[1] + P.gravity[2] * P.gravity[2]);
    const Vec3<S> a_e = X.C_we.transpose() * (X.a - gravity);
    e[0] = dot(vec3_from<S>(P.ia_span), a_e) / S(gn);
    e[1] = dot(vec3_from<S>(P.ia_span + 3), a_e) / S(gn);
}

// InertialAlignmentConstraint::constraintFunction (upright_control/src/inertial_alignment.cpp:7-53), literally
template <typename S>
void inertial_alignment_constraints(const ub_problem_desc_t& P, const Kinematics<S>& X, S* h) {
    const Vec3<S> gravity = vec3_from<S>(P.gravity);
    const Vec3<S> n = vec3_from<S>(P.ia_normal);
    Vec3<S> a = X.C_we.transpose() * (X.a - gravity);
    if (P.ia_use_angular_acceleration) {
        const Mat3<S> Sw = skew3(X.w);
        const Mat3<S> ddC = (skew3(X.al) + Sw * Sw) * X.C_we;
        a = a + ddC * vec3_from<S>(P.ia_com);
    } else if (P.ia_align_with_fixed_vector) {
        a = X.C_we.transpose() * n;
    }
    const S an = dot(n, a), t0 = dot(vec3_from<S>(P.ia_span), a), t1 = dot(vec3_from<S>(P.ia_span + 3), a);
    const S al(P.ia_alpha);
    h[0] = an;
    h[1] = al * an - t0 - t1;
    h[2] = al * an - t0 + t1;
    h[3] = al * an + t0 - t1;
    h[4] = al * an + t0 + t1;
}

// Sphere-sphere distances minus the minimum distance, h >= 0.  Closed form of
// Orientation error of the end effector (end_effector_cost.h:61-67 -> ocs2 PinocchioEndEffectorKinematics::
// getOrientationError -> ocs2::quaternionDistance [EXT: restated from upstream ocs2_robotic_tools
// RotationTransforms.h]): with q the measured and r the desired quaternion,
//     e = q_w r_v - r_w q_v + q_v x r_v        (the vector part of r * q^-1),
// q from the rotation matrix by Eigen's matrix -> quaternion conversion (the branch on the trace decides the sign of
// q, and with it of e, exactly as Eigen::Quaternion(R) does); target quaternions are stored [x, y, z, w]
// (reference_trajectory.h:14-16: Quatd(target.segment<4>(3)) reads Eigen's coefficient order).
template <typename S>
void matrix_to_quaternion(const Mat3<S>& R, S q[4]) {   // q = [x, y, z, w]
    const S tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
    if (value(tr) > 0.0) {
        S t = sqrt(tr + S(1.0));
        q[3] = S(0.5) * t;
        t = S(0.5) / t;
        q[0] = (R.m[2][1] - R.m[1][2]) * t;
        q[1] = (R.m[0][2] - R.m[2][0]) * t;
        q[2] = (R.m[1][0] - R.m[0][1]) * t;
    } else {
        int i = 0;
        if (value(R.m[1][1]) > value(R.m[0][0])) i = 1;
        if (value(R.m[2][2]) > value(R.m[i][i])) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        S t = sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + S(1.0));
        q[i] = S(0.5) * t;
        t = S(0.5) / t;
        q[3] = (R.m[k][j] - R.m[j][k]) * t;
        q[j] = (R.m[j][i] + R.m[i][j]) * t;
        q[k] = (R.m[k][i] + R.m[i][k]) * t;
    }
}
template <typename S>
Vec3<S> orientation_error(const Mat3<S>& C_we, const double* qref /* x y z w */) {
    S q[4];
    matrix_to_quaternion(C_we, q);
    const Vec3<S> qv{q[0], q[1], q[2]};
    const Vec3<S> rv{S(qref[0]), S(qref[1]), S(qref[2])};
    return q[3] * rv - S(qref[3]) * qv + cross(qv, rv);
}

// ocs2::SelfCollisionConstraintCppAd + hpp-fcl for sphere pairs
// (upright_control/src/controller_interface.cpp:450-481).
// The `ground` object of the reference is a half-space (add_ground_plane, controller_interface.cpp:93-101:
// hpp::fcl::Halfspace(UnitZ, 0), solid {p : n.p <= d}): sphere against it -> n.c - d - r.
template <typename S>
void obstacle_constraints(const ub_problem_desc_t& P, const Kinematics<S>& X, S* h) {
    for (int i = 0; i < P.n_pairs; ++i) {
        const int a = P.pairs[i].a, b = P.pairs[i].b;
        if (P.spheres[a].shape == UB_SHAPE_HALFSPACE || P.spheres[b].shape == UB_SHAPE_HALFSPACE) {
            const int hs = P.spheres[b].shape == UB_SHAPE_HALFSPACE ? b : a, s = hs == b ? a : b;
            const Vec3<S> n = vec3_from<S>(P.spheres[hs].offset);
            h[i] = dot(n, X.sphere[s]) - S(P.spheres[hs].radius + P.spheres[s].radius + P.minimum_distance);
            continue;
        }
        const Vec3<S> d = X.sphere[a] - X.sphere[b];
        h[i] = sqrt(dot(d, d)) - S(P.spheres[a].radius + P.spheres[b].radius + P.minimum_distance);
    }
}

// cubic_newtons + projectile_closest_time (constraint/projectile_path_constraint.h:11-44), literally: Newton on
// a t^3 + b t^2 + c t + d from t = 0, at most 10 steps, stop once a step is below 1e-4.
inline double projectile_closest_time(const double* r, const double* r0, const double* v0, const double* g) {
    double gg = 0, vg = 0, vv = 0, dg = 0, dv = 0;
    for (int i = 0; i < 3; ++i) {
        const double dr = r[i] - r0[i];
        gg += g[i] * g[i];
        vg += v0[i] * g[i];
        vv += v0[i] * v0[i];
        dg += dr * g[i];
        dv += dr * v0[i];
    }
    const double a = gg, b = 3 * vg, c = 2 * (vv - dg), d = -2 * dv;
    double t = 0.0;
    for (int it = 0; it < 10; ++it) {
        const double f = a * t * t * t + b * t * t + c * t + d;
        const double df = 3 * a * t * t + 2 * b * t + c;
        const double update = f / df;
        t -= update;
        if (std::fabs(update) < 1e-4) break;
    }
    return t;
}

// ProjectilePathConstraint::getValue / getLinearApproximation (projectile_path_constraint.h:77-146).  x_obs = the
// LAST nine states (state.tail(9)); the time of closest approach is evaluated on values and held fixed in the
// derivative, as the reference does; `tclose[i]` returns it for the obstacle-state Jacobian [I, t I, t^2/2 I].
template <typename S>
void projectile_constraints(const ub_problem_desc_t& P, const Kinematics<S>& X, const double* x_obs, S* h,
                            double* tclose) {
    const double s = P.projectile_active;
    for (int i = 0; i < P.n_projectile_links; ++i) {
        const Vec3<S>& c = X.sphere[P.projectile_spheres[i]];
        const double cv[3] = {value(c[0]), value(c[1]), value(c[2])};
        double t = 0.0;
        if (s > 0.5) t = std::max(0.0, projectile_closest_time(cv, x_obs, x_obs + 3, x_obs + 6));
        const Vec3<S> closest(S(x_obs[0] + t * x_obs[3] + 0.5 * t * t * x_obs[6]),
                              S(x_obs[1] + t * x_obs[4] + 0.5 * t * t * x_obs[7]),
                              S(x_obs[2] + t * x_obs[5] + 0.5 * t * t * x_obs[8]));
        const Vec3<S> delta = c - closest;
        const double w = P.projectile_scale / P.projectile_distances[i];
        h[i] = S(w * s) * (sqrt(dot(delta, delta)) - S(P.projectile_distances[i]));
        if (tclose) tclose[i] = t;
    }
}

}  // namespace orc
