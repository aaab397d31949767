/* covo_oracle.c -- C restatement of the two heavy loops of the CoVO-MPC hot path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Never linked into or called by the product
 * (covo_mpc_b200); used by oracle/oracle_c.py for `bench.py`'s cpu_baseline / `--impl reference`
 * legs and cross-checked against oracle/oracle_np.py in tests/test_oracle_c.py.
 * PARITY UNPINNED for the same reasons as oracle_np.py (the reference is JAX and cannot run here).
 *
 * It follows the REFERENCE's algorithm, not the product's:
 *   oracle_rollout_costs  <-> vmap over N of the lax.scan over H of step_env with reward freeze
 *                             (quadjax/controllers/covo.py:227-263, envs/quadrotor.py:215-263,
 *                              dynamics/free.py:74-155, dynamics/utils.py:266-294, quadrotor.py:479-503)
 *   oracle_hessian_fof    <-> jacfwd(jacfwd(get_cumulated_cost)) (controllers/covo.py:134-185):
 *                             forward-over-forward, one second-order tangent lane (i <= j) at a time
 *                             through the whole unrolled H-step rollout -- all (4H)(4H+1)/2 lanes.
 * OpenMP over samples / tangent lanes (the reference's XLA:CPU runs on all host cores too).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float m, g, max_thrust, dt, alpha, action_scale, pos_limit;
    float max_omega[3];
    int max_steps;
} env_t;

/* ---- float model ------------------------------------------------------------------------- */
static inline float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

static inline float log_pos(float e) { /* dynamics/utils.py:266-274 */
    float l = logf(e + 1.0f);
    return e * 0.4f + clipf(l * 4.f, 0.f, 1.f) * 0.4f + clipf(l * 8.f, 0.f, 1.f) * 0.2f +
           clipf(l * 16.f, 0.f, 1.f) * 0.1f + clipf(l * 32.f, 0.f, 1.f) * 0.1f;
}

static inline float reward_f(const float* x, const float* pt, const float* vt) { /* utils.py:285-294 */
    float ex = pt[0] - x[0], ey = pt[1] - x[1], ez = pt[2] - x[2];
    float vx = vt[0] - x[7], vy = vt[1] - x[8], vz = vt[2] - x[9];
    float err_pos = sqrtf(ex * ex + ey * ey + ez * ez), err_vel = sqrtf(vx * vx + vy * vy + vz * vz);
    float yaw = atan2f(2.f * (x[6] * x[5] + x[3] * x[4]), 1.f - 2.f * (x[4] * x[4] + x[5] * x[5]));
    return 1.3f - 0.05f * err_vel - log_pos(err_pos) - fabsf(yaw) * 0.2f;
}

static inline void step_f(float* x, const float* u, const float* fd, const env_t* c) { /* free.py:74-139 */
    float a0 = clipf(u[0], -1.f, 1.f), a1 = clipf(u[1], -1.f, 1.f), a2 = clipf(u[2], -1.f, 1.f), a3 = clipf(u[3], -1.f, 1.f);
    float thrust = (a0 + 1.f) / 2.f * c->max_thrust * c->action_scale;
    float w0 = a1 * c->max_omega[0] * c->action_scale, w1 = a2 * c->max_omega[1] * c->action_scale,
          w2 = a3 * c->max_omega[2] * c->action_scale;
    float qn = sqrtf(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
    float qx = x[3] / qn, qy = x[4] / qn, qz = x[5] / qn, qw = x[6] / qn;
    float r0 = 2.f * (qx * qz + qy * qw), r1 = 2.f * (qy * qz - qx * qw), r2 = 1.f - 2.f * (qx * qx + qy * qy);
    float o0 = x[10], o1 = x[11], o2 = x[12];
    float d0 = 0.5f * (qw * o0 + (qy * o2 - qz * o1)), d1 = 0.5f * (qw * o1 + (qz * o0 - qx * o2)),
          d2 = 0.5f * (qw * o2 + (qx * o1 - qy * o0)), d3 = -0.5f * (qx * o0 + qy * o1 + qz * o2);
    float im = 1.f / c->m, dt = c->dt;
    x[0] += x[7] * dt; x[1] += x[8] * dt; x[2] += x[9] * dt;
    x[7] += im * (r0 * thrust + fd[0]) * dt;
    x[8] += im * (r1 * thrust + fd[1]) * dt;
    x[9] += (-c->g + im * (r2 * thrust + fd[2])) * dt;
    float n0 = qx + d0 * dt, n1 = qy + d1 * dt, n2 = qz + d2 * dt, n3 = qw + d3 * dt;
    float nn = sqrtf(n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3);
    x[3] = n0 / nn; x[4] = n1 / nn; x[5] = n2 / nn; x[6] = n3 / nn;
    x[10] = c->alpha * o0 + (1.f - c->alpha) * w0;
    x[11] = c->alpha * o1 + (1.f - c->alpha) * w1;
    x[12] = c->alpha * o2 + (1.f - c->alpha) * w2;
}

/* state24: pos3 quat4 vel3 omega3 fd3 ptar3 vtar3 pad2.  a [N][H][4] (already clipped samples). */
void oracle_rollout_costs(const float* envp, const float* state24, int time, const float* pos_traj, const float* vel_traj,
                          int T, const float* a, int N, int H, float discount, float* cost) {
    env_t c;
    c.m = envp[0]; c.g = envp[1]; c.max_thrust = envp[2]; c.dt = envp[3]; c.alpha = envp[4]; c.action_scale = envp[5];
    c.pos_limit = envp[6]; c.max_omega[0] = envp[7]; c.max_omega[1] = envp[8]; c.max_omega[2] = envp[9];
    c.max_steps = (int)envp[10];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float x[13], fd[3], pt[3], vt[3];
        memcpy(x, state24, 13 * sizeof(float));
        memcpy(fd, state24 + 13, 3 * sizeof(float));
        memcpy(pt, state24 + 16, 3 * sizeof(float));
        memcpy(vt, state24 + 19, 3 * sizeof(float));
        float rb = 0.f, sum = 0.f, disc = 1.f;
        int done_before = 0;
        for (int h = 0; h < H; ++h) {
            float r = reward_f(x, pt, vt);
            int done = (time + h >= c.max_steps) || fabsf(x[0]) > c.pos_limit || fabsf(x[1]) > c.pos_limit || fabsf(x[2]) > c.pos_limit;
            step_f(x, a + ((size_t)i * H + h) * 4, fd, &c);
            fd[0] = fd[1] = fd[2] = 0.f;
            int row = time + h + 1;
            if (row > T - 1) row = T - 1;
            for (int k = 0; k < 3; ++k) { pt[k] = pos_traj[row * 3 + k]; vt[k] = vel_traj[row * 3 + k]; }
            if (done_before) r = rb;
            rb = r;
            done_before |= done;
            sum += r * disc;
            disc *= discount;
        }
        cost[i] = -sum;
    }
}

/* ---- hyper-dual model (second-order tangent lane) -------------------------------------------- */
typedef struct { float v, a, b, ab; } hd;
static inline hd H_(float v) { hd r = {v, 0, 0, 0}; return r; }
static inline hd hadd(hd x, hd y) { hd r = {x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; return r; }
static inline hd hsub(hd x, hd y) { hd r = {x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; return r; }
static inline hd hmul(hd x, hd y) {
    hd r = {x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b, x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab};
    return r;
}
static inline hd hscale(hd x, float s) { hd r = {x.v * s, x.a * s, x.b * s, x.ab * s}; return r; }
static inline hd haddf(hd x, float s) { x.v += s; return x; }
static inline hd hun(hd x, float f0, float f1, float f2) { hd r = {f0, f1 * x.a, f1 * x.b, f1 * x.ab + f2 * x.a * x.b}; return r; }
static inline hd hsqrt(hd x) {
    float s = sqrtf(x.v);
    if (s == 0.f) return H_(0.f); /* extension shared with oracle_np.py / quad_model.cuh */
    return hun(x, s, 0.5f / s, -0.25f / (s * x.v));
}
static inline hd hrecip(hd x) { float r = 1.f / x.v; return hun(x, r, -r * r, 2.f * r * r * r); }
static inline hd hlog(hd x) { float r = 1.f / x.v; return hun(x, logf(x.v), r, -r * r); }
static inline hd habs(hd x) { float s = (x.v > 0.f) - (x.v < 0.f); return hun(x, fabsf(x.v), s, 0.f); }
static inline hd hclip(hd x, float lo, float hi) { /* jnp.clip = min(max()), balanced ties */
    float w = (x.v < lo || x.v > hi) ? 0.f : ((x.v == lo || x.v == hi) ? 0.5f : 1.f);
    hd r = {clipf(x.v, lo, hi), w * x.a, w * x.b, w * x.ab};
    return r;
}
static inline hd hatan2(hd y, hd x) {
    float r = x.v * x.v + y.v * y.v, ir = 1.f / r;
    float wa = x.v * y.a - y.v * x.a, wb = x.v * y.b - y.v * x.b, drb = 2.f * (x.v * x.b + y.v * y.b);
    hd o = {atan2f(y.v, x.v), wa * ir, wb * ir, (x.v * y.ab - y.v * x.ab + x.b * y.a - y.b * x.a) * ir - wa * drb * ir * ir};
    return o;
}
static inline hd hlog_pos(hd e) {
    hd l = hlog(haddf(e, 1.f));
    hd r = hscale(e, 0.4f);
    r = hadd(r, hscale(hclip(hscale(l, 4.f), 0.f, 1.f), 0.4f));
    r = hadd(r, hscale(hclip(hscale(l, 8.f), 0.f, 1.f), 0.2f));
    r = hadd(r, hscale(hclip(hscale(l, 16.f), 0.f, 1.f), 0.1f));
    r = hadd(r, hscale(hclip(hscale(l, 32.f), 0.f, 1.f), 0.1f));
    return r;
}
static inline hd hreward(const hd* x, const float* pt, const float* vt) {
    hd ex = hsub(H_(pt[0]), x[0]), ey = hsub(H_(pt[1]), x[1]), ez = hsub(H_(pt[2]), x[2]);
    hd vx = hsub(H_(vt[0]), x[7]), vy = hsub(H_(vt[1]), x[8]), vz = hsub(H_(vt[2]), x[9]);
    hd ep = hsqrt(hadd(hadd(hmul(ex, ex), hmul(ey, ey)), hmul(ez, ez)));
    hd ev = hsqrt(hadd(hadd(hmul(vx, vx), hmul(vy, vy)), hmul(vz, vz)));
    hd yn = hscale(hadd(hmul(x[6], x[5]), hmul(x[3], x[4])), 2.f);
    hd yd = hsub(H_(1.f), hscale(hadd(hmul(x[4], x[4]), hmul(x[5], x[5])), 2.f));
    hd yaw = hatan2(yn, yd);
    hd r = hsub(H_(1.3f), hscale(ev, 0.05f));
    r = hsub(r, hlog_pos(ep));
    r = hsub(r, hscale(habs(yaw), 0.2f));
    return r;
}
static inline void hstep(hd* x, const hd* u, const float* fd, const env_t* c) {
    hd a0 = hclip(hclip(u[0], -1.f, 1.f), -1.f, 1.f), a1 = hclip(hclip(u[1], -1.f, 1.f), -1.f, 1.f);
    hd a2 = hclip(hclip(u[2], -1.f, 1.f), -1.f, 1.f), a3 = hclip(hclip(u[3], -1.f, 1.f), -1.f, 1.f);
    hd thrust = hscale(haddf(a0, 1.f), 0.5f * c->max_thrust * c->action_scale);
    hd w0 = hscale(a1, c->max_omega[0] * c->action_scale), w1 = hscale(a2, c->max_omega[1] * c->action_scale),
       w2 = hscale(a3, c->max_omega[2] * c->action_scale);
    hd qn = hrecip(hsqrt(hadd(hadd(hmul(x[3], x[3]), hmul(x[4], x[4])), hadd(hmul(x[5], x[5]), hmul(x[6], x[6])))));
    hd qx = hmul(x[3], qn), qy = hmul(x[4], qn), qz = hmul(x[5], qn), qw = hmul(x[6], qn);
    hd r0 = hscale(hadd(hmul(qx, qz), hmul(qy, qw)), 2.f), r1 = hscale(hsub(hmul(qy, qz), hmul(qx, qw)), 2.f);
    hd r2 = hsub(H_(1.f), hscale(hadd(hmul(qx, qx), hmul(qy, qy)), 2.f));
    hd o0 = x[10], o1 = x[11], o2 = x[12];
    hd d0 = hscale(hadd(hmul(qw, o0), hsub(hmul(qy, o2), hmul(qz, o1))), 0.5f);
    hd d1 = hscale(hadd(hmul(qw, o1), hsub(hmul(qz, o0), hmul(qx, o2))), 0.5f);
    hd d2 = hscale(hadd(hmul(qw, o2), hsub(hmul(qx, o1), hmul(qy, o0))), 0.5f);
    hd d3 = hscale(hadd(hadd(hmul(qx, o0), hmul(qy, o1)), hmul(qz, o2)), -0.5f);
    float im = 1.f / c->m, dt = c->dt;
    x[0] = hadd(x[0], hscale(x[7], dt)); x[1] = hadd(x[1], hscale(x[8], dt)); x[2] = hadd(x[2], hscale(x[9], dt));
    x[7] = hadd(x[7], hscale(haddf(hmul(r0, thrust), fd[0]), im * dt));
    x[8] = hadd(x[8], hscale(haddf(hmul(r1, thrust), fd[1]), im * dt));
    x[9] = hadd(x[9], hscale(haddf(hscale(haddf(hmul(r2, thrust), fd[2]), im), -c->g), dt));
    hd n0 = hadd(qx, hscale(d0, dt)), n1 = hadd(qy, hscale(d1, dt)), n2 = hadd(qz, hscale(d2, dt)), n3 = hadd(qw, hscale(d3, dt));
    hd nn = hrecip(hsqrt(hadd(hadd(hmul(n0, n0), hmul(n1, n1)), hadd(hmul(n2, n2), hmul(n3, n3)))));
    x[3] = hmul(n0, nn); x[4] = hmul(n1, nn); x[5] = hmul(n2, nn); x[6] = hmul(n3, nn);
    x[10] = hadd(hscale(o0, c->alpha), hscale(w0, 1.f - c->alpha));
    x[11] = hadd(hscale(o1, c->alpha), hscale(w1, 1.f - c->alpha));
    x[12] = hadd(hscale(o2, c->alpha), hscale(w2, 1.f - c->alpha));
}

/* R [n][n], n = 4H.  a_mean [H][4] (already shifted). */
void oracle_hessian_fof(const float* envp, const float* state24, int time, const float* pos_traj, const float* vel_traj,
                        int T, const float* a_mean, int H, float* R) {
    env_t c;
    c.m = envp[0]; c.g = envp[1]; c.max_thrust = envp[2]; c.dt = envp[3]; c.alpha = envp[4]; c.action_scale = envp[5];
    c.pos_limit = envp[6]; c.max_omega[0] = envp[7]; c.max_omega[1] = envp[8]; c.max_omega[2] = envp[9];
    c.max_steps = (int)envp[10];
    const int n = 4 * H;
    const long npairs = (long)n * (n + 1) / 2;
#pragma omp parallel for schedule(dynamic, 64)
    for (long pi = 0; pi < npairs; ++pi) {
        /* pair (i <= j) from the linear index */
        int i = 0; long rem = pi; int cnt = n;
        while (rem >= cnt) { rem -= cnt; --cnt; ++i; }
        int j = i + (int)rem;
        hd x[13];
        for (int k = 0; k < 13; ++k) x[k] = H_(state24[k]);
        float fd[3] = {state24[13], state24[14], state24[15]};
        float pt[3] = {state24[16], state24[17], state24[18]}, vt[3] = {state24[19], state24[20], state24[21]};
        hd total = H_(0.f);
        for (int h = 0; h < H; ++h) {
            total = hadd(total, hreward(x, pt, vt));
            hd u[4];
            for (int k = 0; k < 4; ++k) {
                int idx = 4 * h + k;
                u[k] = H_(a_mean[idx]);
                if (idx == i) u[k].a = 1.f;
                if (idx == j) u[k].b = 1.f;
            }
            hstep(x, u, fd, &c);
            fd[0] = fd[1] = fd[2] = 0.f;
            int row = time + h + 1;
            if (row > T - 1) row = T - 1;
            for (int k = 0; k < 3; ++k) { pt[k] = pos_traj[row * 3 + k]; vt[k] = vel_traj[row * 3 + k]; }
        }
        float val = -total.ab;
        R[(size_t)i * n + j] = val;
        R[(size_t)j * n + i] = val;
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
