// Quadrotor body-rate model, tracking reward and termination, written once over a scalar
// type S in {float, Dual, HDual}.  The float instantiation is what every rollout thread runs
// (state in registers); Dual / HDual (first / second order forward AD) are what the CoVO
// Hessian kernels run on ONE transition at a time -- the same source of truth for both.
//
// Follows (reference file:line, relative to /root/reference):
//   envs/quadrotor.py:215-263   Quad3D.step_env / raw_step (clip, thrust / body-rate mapping)
//   dynamics/free.py:74-112     quad_dynamics_bodyrate (explicit Euler, old-velocity position update)
//   dynamics/free.py:114-155    free_dynamics_3d_bodyrate (re-normalise q, time+1, clamped target gather)
//   dynamics/geom.py:41-77      L(q), H, qtoQ -- only Q(q) e3 and 0.5 L(q) H w are needed
//   dynamics/utils.py:266-294   log_pos_fn, tracking_penyaw_reward_fn
//   envs/quadrotor.py:479-503   is_terminal with disable_rollover_terminate=True
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define COVO_HD __host__ __device__ __forceinline__
#else
#define COVO_HD inline
#endif

namespace covo {

// Physical constants of EnvParams3D (dynamics/dataclass.py:40-100) used on the hot path.
struct EnvConsts {
    float m, g, max_thrust, dt;
    float alpha_bodyrate, action_scale, pos_limit;
    float max_omega[3];
    int max_steps;  // max_steps_in_episode
};

// ---------------------------------------------------------------------------------------------
// scalar algebra
// ---------------------------------------------------------------------------------------------
struct Dual {
    float v, d;
};
struct HDual {
    float v, a, b, ab;  // value, d/da, d/db, d2/(da db)
};

COVO_HD float cst(float, float x) { return x; }
COVO_HD Dual cst(Dual, float x) { return Dual{x, 0.f}; }
COVO_HD HDual cst(HDual, float x) { return HDual{x, 0.f, 0.f, 0.f}; }
COVO_HD float val(float x) { return x; }
COVO_HD float val(Dual x) { return x.v; }
COVO_HD float val(HDual x) { return x.v; }

// ---- Dual
COVO_HD Dual operator+(Dual x, Dual y) { return Dual{x.v + y.v, x.d + y.d}; }
COVO_HD Dual operator-(Dual x, Dual y) { return Dual{x.v - y.v, x.d - y.d}; }
COVO_HD Dual operator-(Dual x) { return Dual{-x.v, -x.d}; }
COVO_HD Dual operator*(Dual x, Dual y) { return Dual{x.v * y.v, x.d * y.v + x.v * y.d}; }
COVO_HD Dual operator+(Dual x, float y) { return Dual{x.v + y, x.d}; }
COVO_HD Dual operator+(float y, Dual x) { return Dual{x.v + y, x.d}; }
COVO_HD Dual operator-(Dual x, float y) { return Dual{x.v - y, x.d}; }
COVO_HD Dual operator-(float y, Dual x) { return Dual{y - x.v, -x.d}; }
COVO_HD Dual operator*(Dual x, float y) { return Dual{x.v * y, x.d * y}; }
COVO_HD Dual operator*(float y, Dual x) { return Dual{x.v * y, x.d * y}; }
COVO_HD Dual unary(Dual x, float f0, float f1, float) { return Dual{f0, f1 * x.d}; }
// ---- HDual
COVO_HD HDual operator+(HDual x, HDual y) { return HDual{x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
COVO_HD HDual operator-(HDual x, HDual y) { return HDual{x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
COVO_HD HDual operator-(HDual x) { return HDual{-x.v, -x.a, -x.b, -x.ab}; }
COVO_HD HDual operator*(HDual x, HDual y) {
    return HDual{x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b,
                 x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab};
}
COVO_HD HDual operator+(HDual x, float y) { return HDual{x.v + y, x.a, x.b, x.ab}; }
COVO_HD HDual operator+(float y, HDual x) { return HDual{x.v + y, x.a, x.b, x.ab}; }
COVO_HD HDual operator-(HDual x, float y) { return HDual{x.v - y, x.a, x.b, x.ab}; }
COVO_HD HDual operator-(float y, HDual x) { return HDual{y - x.v, -x.a, -x.b, -x.ab}; }
COVO_HD HDual operator*(HDual x, float y) { return HDual{x.v * y, x.a * y, x.b * y, x.ab * y}; }
COVO_HD HDual operator*(float y, HDual x) { return HDual{x.v * y, x.a * y, x.b * y, x.ab * y}; }
COVO_HD HDual unary(HDual x, float f0, float f1, float f2) {
    return HDual{f0, f1 * x.a, f1 * x.b, f1 * x.ab + f2 * x.a * x.b};
}

// ---- reciprocal / division
COVO_HD float recip_(float x) { return 1.0f / x; }
COVO_HD Dual recip_(Dual x) {
    float r = 1.0f / x.v;
    return unary(x, r, -r * r, 0.f);
}
COVO_HD HDual recip_(HDual x) {
    float r = 1.0f / x.v;
    return unary(x, r, -r * r, 2.0f * r * r * r);
}
// ---- sqrt.  d/dx sqrt at 0 is singular; the reference (JAX) would produce NaN there only when
// the tangent is instantiated.  EXTENSION shared with the oracle: locally constant at exactly 0.
COVO_HD float sqrt_(float x) { return sqrtf(x); }
COVO_HD Dual sqrt_(Dual x) {
    float s = sqrtf(x.v);
    if (s == 0.f) return Dual{0.f, 0.f};
    return unary(x, s, 0.5f / s, 0.f);
}
COVO_HD HDual sqrt_(HDual x) {
    float s = sqrtf(x.v);
    if (s == 0.f) return HDual{0.f, 0.f, 0.f, 0.f};
    return unary(x, s, 0.5f / s, -0.25f / (s * x.v));
}
// 1/sqrt(x) for the two quaternion normalisations: MUFU.RSQ + one Newton step in the float path
// (<= 1 ulp from the reference's q / ||q||), exact chain rule in the AD paths.
COVO_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
    float y = rsqrtf(x);
    return y * (1.5f - 0.5f * x * y * y);
#else
    return 1.0f / sqrtf(x);
#endif
}
COVO_HD Dual rsqrt_(Dual x) {
    float y = 1.0f / sqrtf(x.v);
    return unary(x, y, -0.5f * y / x.v, 0.f);
}
COVO_HD HDual rsqrt_(HDual x) {
    float y = 1.0f / sqrtf(x.v);
    return unary(x, y, -0.5f * y / x.v, 0.75f * y / (x.v * x.v));
}

COVO_HD float log_(float x) { return logf(x); }
COVO_HD Dual log_(Dual x) { return unary(x, logf(x.v), 1.0f / x.v, 0.f); }
COVO_HD HDual log_(HDual x) {
    float r = 1.0f / x.v;
    return unary(x, logf(x.v), r, -r * r);
}

// jnp.abs JVP = sign(x) * tangent, sign(0) = 0
COVO_HD float sgn_(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
COVO_HD float abs_(float x) { return fabsf(x); }
COVO_HD Dual abs_(Dual x) { return unary(x, fabsf(x.v), sgn_(x.v), 0.f); }
COVO_HD HDual abs_(HDual x) { return unary(x, fabsf(x.v), sgn_(x.v), 0.f); }

// jnp.clip == minimum(maximum(x, lo), hi): derivative 1 inside, 0 outside, 1/2 exactly on a bound
// (lax.max / lax.min JVPs split ties evenly).
COVO_HD float clip_w(float v, float lo, float hi) {
    if (v < lo || v > hi) return 0.f;
    return (v == lo || v == hi) ? 0.5f : 1.0f;
}
COVO_HD float clip_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
COVO_HD Dual clip_(Dual x, float lo, float hi) {
    float w = clip_w(x.v, lo, hi);
    return Dual{fminf(fmaxf(x.v, lo), hi), w * x.d};
}
COVO_HD HDual clip_(HDual x, float lo, float hi) {
    float w = clip_w(x.v, lo, hi);
    return HDual{fminf(fmaxf(x.v, lo), hi), w * x.a, w * x.b, w * x.ab};
}

// atan2(y, x):  d = (x dy - y dx) / r,  r = x^2 + y^2
COVO_HD float atan2_(float y, float x) { return atan2f(y, x); }
COVO_HD Dual atan2_(Dual y, Dual x) {
    float r = x.v * x.v + y.v * y.v;
    return Dual{atan2f(y.v, x.v), (x.v * y.d - y.v * x.d) / r};
}
COVO_HD HDual atan2_(HDual y, HDual x) {
    float r = x.v * x.v + y.v * y.v;
    float ir = 1.0f / r;
    float wa = x.v * y.a - y.v * x.a;
    float wb = x.v * y.b - y.v * x.b;
    float drb = 2.0f * (x.v * x.b + y.v * y.b);
    float ab = (x.v * y.ab - y.v * x.ab + x.b * y.a - y.b * x.a) * ir - wa * drb * ir * ir;
    return HDual{atan2f(y.v, x.v), wa * ir, wb * ir, ab};
}

// ---------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------
template <class S>
struct QState {
    S p[3];  // position (world)
    S q[4];  // quaternion (x, y, z, w)   dynamics/dataclass.py:14
    S v[3];  // velocity (world)
    S w[3];  // body rates
};

// log_pos_fn, dynamics/utils.py:266-274
template <class S>
COVO_HD S log_pos(S e) {
    S l = log_(e + 1.0f);
    return e * 0.4f + clip_(l * 4.0f, 0.f, 1.f) * 0.4f + clip_(l * 8.0f, 0.f, 1.f) * 0.2f +
           clip_(l * 16.0f, 0.f, 1.f) * 0.1f + clip_(l * 32.0f, 0.f, 1.f) * 0.1f;
}

// tracking_penyaw_reward_fn, dynamics/utils.py:285-294.  Evaluated on the STORED quaternion
// (un-normalised at h = 0, where it is the noisy quaternion).
template <class S>
COVO_HD S quad_reward(const QState<S>& s, const float ptar[3], const float vtar[3]) {
    S ex = ptar[0] - s.p[0], ey = ptar[1] - s.p[1], ez = ptar[2] - s.p[2];
    S err_pos = sqrt_(ex * ex + ey * ey + ez * ez);
    S vx = vtar[0] - s.v[0], vy = vtar[1] - s.v[1], vz = vtar[2] - s.v[2];
    S err_vel = sqrt_(vx * vx + vy * vy + vz * vz);
    S yaw = atan2_((s.q[3] * s.q[2] + s.q[0] * s.q[1]) * 2.0f, 1.0f - (s.q[1] * s.q[1] + s.q[2] * s.q[2]) * 2.0f);
    return 1.3f - err_vel * 0.05f - log_pos(err_pos) - abs_(yaw) * 0.2f;
}

// is_terminal, envs/quadrotor.py:479-503 (rollover terms disabled by main, :779)
template <class S>
COVO_HD bool quad_terminal(const QState<S>& s, int time, const EnvConsts& c) {
    return (time >= c.max_steps) || (fabsf(val(s.p[0])) > c.pos_limit) || (fabsf(val(s.p[1])) > c.pos_limit) ||
           (fabsf(val(s.p[2])) > c.pos_limit);
}

// One step_env transition of the dynamic state.  `fd` is the disturbance force acting DURING this
// step (state.f_disturb, dynamics/free.py:92-99); the caller owns what it becomes afterwards.
template <class S>
COVO_HD void quad_step(QState<S>& s, const S u_in[4], const float fd[3], const EnvConsts& c) {
    // envs/quadrotor.py:223 and :257 -- clipped twice
    S a0 = clip_(clip_(u_in[0], -1.f, 1.f), -1.f, 1.f);
    S a1 = clip_(clip_(u_in[1], -1.f, 1.f), -1.f, 1.f);
    S a2 = clip_(clip_(u_in[2], -1.f, 1.f), -1.f, 1.f);
    S a3 = clip_(clip_(u_in[3], -1.f, 1.f), -1.f, 1.f);
    S thrust = (a0 + 1.0f) * (0.5f * c.max_thrust * c.action_scale);  // quadrotor.py:258, free.py:82
    // torque / max_torque * max_omega == a * max_omega (free.py:122) up to 1 ulp
    S wt0 = a1 * (c.max_omega[0] * c.action_scale);
    S wt1 = a2 * (c.max_omega[1] * c.action_scale);
    S wt2 = a3 * (c.max_omega[2] * c.action_scale);

    // q <- q / ||q||   (free.py:88)
    S qn = rsqrt_(s.q[0] * s.q[0] + s.q[1] * s.q[1] + s.q[2] * s.q[2] + s.q[3] * s.q[3]);
    S x = s.q[0] * qn, y = s.q[1] * qn, z = s.q[2] * qn, w = s.q[3] * qn;
    // Q(q) e3 (third column of geom.qtoQ, geom.py:68-77)
    S r0 = (x * z + y * w) * 2.0f;
    S r1 = (y * z - x * w) * 2.0f;
    S r2 = 1.0f - (x * x + y * y) * 2.0f;
    // 0.5 L(q) H omega  (geom.py:41-55, free.py:96)
    S o0 = s.w[0], o1 = s.w[1], o2 = s.w[2];
    S qd0 = (w * o0 + (y * o2 - z * o1)) * 0.5f;
    S qd1 = (w * o1 + (z * o0 - x * o2)) * 0.5f;
    S qd2 = (w * o2 + (x * o1 - y * o0)) * 0.5f;
    S qd3 = (x * o0 + y * o1 + z * o2) * (-0.5f);
    const float inv_m = 1.0f / c.m;
    const float dt = c.dt;
    // explicit Euler; the position update uses the OLD velocity (free.py:102-103)
    s.p[0] = s.p[0] + s.v[0] * dt;
    s.p[1] = s.p[1] + s.v[1] * dt;
    s.p[2] = s.p[2] + s.v[2] * dt;
    s.v[0] = s.v[0] + (r0 * thrust + fd[0]) * (inv_m * dt);
    s.v[1] = s.v[1] + (r1 * thrust + fd[1]) * (inv_m * dt);
    s.v[2] = s.v[2] + ((r2 * thrust + fd[2]) * inv_m - c.g) * dt;
    S n0 = x + qd0 * dt, n1 = y + qd1 * dt, n2 = z + qd2 * dt, n3 = w + qd3 * dt;
    S nn = rsqrt_(n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3);  // free.py:139
    s.q[0] = n0 * nn;
    s.q[1] = n1 * nn;
    s.q[2] = n2 * nn;
    s.q[3] = n3 * nn;
    const float al = c.alpha_bodyrate;
    s.w[0] = o0 * al + wt0 * (1.0f - al);  // free.py:105-107
    s.w[1] = o1 * al + wt1 * (1.0f - al);
    s.w[2] = o2 * al + wt2 * (1.0f - al);
}

// The C-ABI state record (include/covo_b200.h, covo_state24): 24 floats
//   [0:3] pos  [3:7] quat(xyzw)  [7:10] vel  [10:13] omega  [13:16] f_disturb
//   [16:19] pos_tar  [19:22] vel_tar  [22:24] pad
COVO_HD void load_state24(const float* s24, QState<float>& s, float fd[3], float ptar[3], float vtar[3]) {
    for (int k = 0; k < 3; ++k) s.p[k] = s24[k];
    for (int k = 0; k < 4; ++k) s.q[k] = s24[3 + k];
    for (int k = 0; k < 3; ++k) s.v[k] = s24[7 + k];
    for (int k = 0; k < 3; ++k) s.w[k] = s24[10 + k];
    for (int k = 0; k < 3; ++k) fd[k] = s24[13 + k];
    for (int k = 0; k < 3; ++k) ptar[k] = s24[16 + k];
    for (int k = 0; k < 3; ++k) vtar[k] = s24[19 + k];
}

}  // namespace covo
