/* covo_oracle_impl.h -- body of covo_oracle.c, written once over the scalar type REAL (float: the reference's arithmetic;
 * double: the same algorithm in float64, the "exact" answer the float32 results are compared with).  Included twice by
 * covo_oracle.c with REAL / N_(name) / K(literal) / M_(libm name) defined.  TEST INFRASTRUCTURE ONLY (see covo_oracle.c). */
typedef struct {
    REAL m, g, max_thrust, dt, alpha, action_scale, pos_limit;
    REAL max_omega[3];
    int max_steps;
} N_(env_t);

/* ---- REAL model ------------------------------------------------------------------------- */
static inline REAL N_(clipf)(REAL x, REAL lo, REAL hi) { return M_(fmin)(M_(fmax)(x, lo), hi); }

static inline REAL N_(log_pos)(REAL e) { /* dynamics/utils.py:266-274 */
    REAL l = M_(log)(e + K(1.0));
    return e * K(0.4) + N_(clipf)(l * K(4.), K(0.), K(1.)) * K(0.4) + N_(clipf)(l * K(8.), K(0.), K(1.)) * K(0.2) +
           N_(clipf)(l * K(16.), K(0.), K(1.)) * K(0.1) + N_(clipf)(l * K(32.), K(0.), K(1.)) * K(0.1);
}

static inline REAL N_(reward_f)(const REAL* x, const REAL* pt, const REAL* vt) { /* utils.py:285-294 */
    REAL ex = pt[0] - x[0], ey = pt[1] - x[1], ez = pt[2] - x[2];
    REAL vx = vt[0] - x[7], vy = vt[1] - x[8], vz = vt[2] - x[9];
    REAL err_pos = M_(sqrt)(ex * ex + ey * ey + ez * ez), err_vel = M_(sqrt)(vx * vx + vy * vy + vz * vz);
    REAL yaw = M_(atan2)(K(2.) * (x[6] * x[5] + x[3] * x[4]), K(1.) - K(2.) * (x[4] * x[4] + x[5] * x[5]));
    return K(1.3) - K(0.05) * err_vel - N_(log_pos)(err_pos) - M_(fabs)(yaw) * K(0.2);
}

static inline void N_(step_f)(REAL* x, const REAL* u, const REAL* fd, const N_(env_t)* c) { /* free.py:74-139 */
    REAL a0 = N_(clipf)(u[0], -K(1.), K(1.)), a1 = N_(clipf)(u[1], -K(1.), K(1.)), a2 = N_(clipf)(u[2], -K(1.), K(1.)), a3 = N_(clipf)(u[3], -K(1.), K(1.));
    REAL thrust = (a0 + K(1.)) / K(2.) * c->max_thrust * c->action_scale;
    REAL w0 = a1 * c->max_omega[0] * c->action_scale, w1 = a2 * c->max_omega[1] * c->action_scale,
          w2 = a3 * c->max_omega[2] * c->action_scale;
    REAL qn = M_(sqrt)(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
    REAL qx = x[3] / qn, qy = x[4] / qn, qz = x[5] / qn, qw = x[6] / qn;
    REAL r0 = K(2.) * (qx * qz + qy * qw), r1 = K(2.) * (qy * qz - qx * qw), r2 = K(1.) - K(2.) * (qx * qx + qy * qy);
    REAL o0 = x[10], o1 = x[11], o2 = x[12];
    REAL d0 = K(0.5) * (qw * o0 + (qy * o2 - qz * o1)), d1 = K(0.5) * (qw * o1 + (qz * o0 - qx * o2)),
          d2 = K(0.5) * (qw * o2 + (qx * o1 - qy * o0)), d3 = -K(0.5) * (qx * o0 + qy * o1 + qz * o2);
    REAL im = K(1.) / c->m, dt = c->dt;
    x[0] += x[7] * dt; x[1] += x[8] * dt; x[2] += x[9] * dt;
    x[7] += im * (r0 * thrust + fd[0]) * dt;
    x[8] += im * (r1 * thrust + fd[1]) * dt;
    x[9] += (-c->g + im * (r2 * thrust + fd[2])) * dt;
    REAL n0 = qx + d0 * dt, n1 = qy + d1 * dt, n2 = qz + d2 * dt, n3 = qw + d3 * dt;
    REAL nn = M_(sqrt)(n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3);
    x[3] = n0 / nn; x[4] = n1 / nn; x[5] = n2 / nn; x[6] = n3 / nn;
    x[10] = c->alpha * o0 + (K(1.) - c->alpha) * w0;
    x[11] = c->alpha * o1 + (K(1.) - c->alpha) * w1;
    x[12] = c->alpha * o2 + (K(1.) - c->alpha) * w2;
}

/* state24: pos3 quat4 vel3 omega3 fd3 ptar3 vtar3 pad2.  a [N][H][4] (already clipped samples). */
void N_(oracle_rollout_costs)(const float* envp, const float* state24, int time, const float* pos_traj, const float* vel_traj,
                          int T, const float* a, int N, int H, float discount, REAL* cost) {
    N_(env_t) c;
    c.m = envp[0]; c.g = envp[1]; c.max_thrust = envp[2]; c.dt = envp[3]; c.alpha = envp[4]; c.action_scale = envp[5];
    c.pos_limit = envp[6]; c.max_omega[0] = envp[7]; c.max_omega[1] = envp[8]; c.max_omega[2] = envp[9];
    c.max_steps = (int)envp[10];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        REAL x[13], fd[3], pt[3], vt[3];
        for (int k = 0; k < 13; ++k) x[k] = state24[k];
        for (int k = 0; k < 3; ++k) { fd[k] = state24[13 + k]; pt[k] = state24[16 + k]; vt[k] = state24[19 + k]; }
        REAL rb = K(0.), sum = K(0.), disc = K(1.);
        int done_before = 0;
        for (int h = 0; h < H; ++h) {
            REAL r = N_(reward_f)(x, pt, vt);
            int done = (time + h >= c.max_steps) || M_(fabs)(x[0]) > c.pos_limit || M_(fabs)(x[1]) > c.pos_limit || M_(fabs)(x[2]) > c.pos_limit;
            const float* ai = a + ((size_t)i * H + h) * 4;
            REAL u4[4] = {ai[0], ai[1], ai[2], ai[3]};
            N_(step_f)(x, u4, fd, &c);
            fd[0] = fd[1] = fd[2] = K(0.);
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
typedef struct { REAL v, a, b, ab; } N_(hd);
static inline N_(hd) N_(H_)(REAL v) { N_(hd) r = {v, 0, 0, 0}; return r; }
static inline N_(hd) N_(hadd)(N_(hd) x, N_(hd) y) { N_(hd) r = {x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; return r; }
static inline N_(hd) N_(hsub)(N_(hd) x, N_(hd) y) { N_(hd) r = {x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; return r; }
static inline N_(hd) N_(hmul)(N_(hd) x, N_(hd) y) {
    N_(hd) r = {x.v * y.v, x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b, x.ab * y.v + x.a * y.b + x.b * y.a + x.v * y.ab};
    return r;
}
static inline N_(hd) N_(hscale)(N_(hd) x, REAL s) { N_(hd) r = {x.v * s, x.a * s, x.b * s, x.ab * s}; return r; }
static inline N_(hd) N_(haddf)(N_(hd) x, REAL s) { x.v += s; return x; }
static inline N_(hd) N_(hun)(N_(hd) x, REAL f0, REAL f1, REAL f2) { N_(hd) r = {f0, f1 * x.a, f1 * x.b, f1 * x.ab + f2 * x.a * x.b}; return r; }
static inline N_(hd) N_(hsqrt)(N_(hd) x) {
    REAL s = M_(sqrt)(x.v);
    if (s == K(0.)) return N_(H_)(K(0.)); /* extension shared with oracle_np.py / quad_model.cuh */
    return N_(hun)(x, s, K(0.5) / s, -K(0.25) / (s * x.v));
}
static inline N_(hd) N_(hrecip)(N_(hd) x) { REAL r = K(1.) / x.v; return N_(hun)(x, r, -r * r, K(2.) * r * r * r); }
static inline N_(hd) N_(hlog)(N_(hd) x) { REAL r = K(1.) / x.v; return N_(hun)(x, M_(log)(x.v), r, -r * r); }
static inline N_(hd) N_(habs)(N_(hd) x) { REAL s = (x.v > K(0.)) - (x.v < K(0.)); return N_(hun)(x, M_(fabs)(x.v), s, K(0.)); }
static inline N_(hd) N_(hclip)(N_(hd) x, REAL lo, REAL hi) { /* jnp.clip = min(max()), balanced ties */
    REAL w = (x.v < lo || x.v > hi) ? K(0.) : ((x.v == lo || x.v == hi) ? K(0.5) : K(1.));
    N_(hd) r = {N_(clipf)(x.v, lo, hi), w * x.a, w * x.b, w * x.ab};
    return r;
}
static inline N_(hd) N_(hatan2)(N_(hd) y, N_(hd) x) {
    REAL r = x.v * x.v + y.v * y.v, ir = K(1.) / r;
    REAL wa = x.v * y.a - y.v * x.a, wb = x.v * y.b - y.v * x.b, drb = K(2.) * (x.v * x.b + y.v * y.b);
    N_(hd) o = {M_(atan2)(y.v, x.v), wa * ir, wb * ir, (x.v * y.ab - y.v * x.ab + x.b * y.a - y.b * x.a) * ir - wa * drb * ir * ir};
    return o;
}
static inline N_(hd) N_(hlog_pos)(N_(hd) e) {
    N_(hd) l = N_(hlog)(N_(haddf)(e, K(1.)));
    N_(hd) r = N_(hscale)(e, K(0.4));
    r = N_(hadd)(r, N_(hscale)(N_(hclip)(N_(hscale)(l, K(4.)), K(0.), K(1.)), K(0.4)));
    r = N_(hadd)(r, N_(hscale)(N_(hclip)(N_(hscale)(l, K(8.)), K(0.), K(1.)), K(0.2)));
    r = N_(hadd)(r, N_(hscale)(N_(hclip)(N_(hscale)(l, K(16.)), K(0.), K(1.)), K(0.1)));
    r = N_(hadd)(r, N_(hscale)(N_(hclip)(N_(hscale)(l, K(32.)), K(0.), K(1.)), K(0.1)));
    return r;
}
static inline N_(hd) N_(hreward)(const N_(hd)* x, const REAL* pt, const REAL* vt) {
    N_(hd) ex = N_(hsub)(N_(H_)(pt[0]), x[0]), ey = N_(hsub)(N_(H_)(pt[1]), x[1]), ez = N_(hsub)(N_(H_)(pt[2]), x[2]);
    N_(hd) vx = N_(hsub)(N_(H_)(vt[0]), x[7]), vy = N_(hsub)(N_(H_)(vt[1]), x[8]), vz = N_(hsub)(N_(H_)(vt[2]), x[9]);
    N_(hd) ep = N_(hsqrt)(N_(hadd)(N_(hadd)(N_(hmul)(ex, ex), N_(hmul)(ey, ey)), N_(hmul)(ez, ez)));
    N_(hd) ev = N_(hsqrt)(N_(hadd)(N_(hadd)(N_(hmul)(vx, vx), N_(hmul)(vy, vy)), N_(hmul)(vz, vz)));
    N_(hd) yn = N_(hscale)(N_(hadd)(N_(hmul)(x[6], x[5]), N_(hmul)(x[3], x[4])), K(2.));
    N_(hd) yd = N_(hsub)(N_(H_)(K(1.)), N_(hscale)(N_(hadd)(N_(hmul)(x[4], x[4]), N_(hmul)(x[5], x[5])), K(2.)));
    N_(hd) yaw = N_(hatan2)(yn, yd);
    N_(hd) r = N_(hsub)(N_(H_)(K(1.3)), N_(hscale)(ev, K(0.05)));
    r = N_(hsub)(r, N_(hlog_pos)(ep));
    r = N_(hsub)(r, N_(hscale)(N_(habs)(yaw), K(0.2)));
    return r;
}
static inline void N_(hstep)(N_(hd)* x, const N_(hd)* u, const REAL* fd, const N_(env_t)* c) {
    N_(hd) a0 = N_(hclip)(N_(hclip)(u[0], -K(1.), K(1.)), -K(1.), K(1.)), a1 = N_(hclip)(N_(hclip)(u[1], -K(1.), K(1.)), -K(1.), K(1.));
    N_(hd) a2 = N_(hclip)(N_(hclip)(u[2], -K(1.), K(1.)), -K(1.), K(1.)), a3 = N_(hclip)(N_(hclip)(u[3], -K(1.), K(1.)), -K(1.), K(1.));
    N_(hd) thrust = N_(hscale)(N_(haddf)(a0, K(1.)), K(0.5) * c->max_thrust * c->action_scale);
    N_(hd) w0 = N_(hscale)(a1, c->max_omega[0] * c->action_scale), w1 = N_(hscale)(a2, c->max_omega[1] * c->action_scale),
       w2 = N_(hscale)(a3, c->max_omega[2] * c->action_scale);
    N_(hd) qn = N_(hrecip)(N_(hsqrt)(N_(hadd)(N_(hadd)(N_(hmul)(x[3], x[3]), N_(hmul)(x[4], x[4])), N_(hadd)(N_(hmul)(x[5], x[5]), N_(hmul)(x[6], x[6])))));
    N_(hd) qx = N_(hmul)(x[3], qn), qy = N_(hmul)(x[4], qn), qz = N_(hmul)(x[5], qn), qw = N_(hmul)(x[6], qn);
    N_(hd) r0 = N_(hscale)(N_(hadd)(N_(hmul)(qx, qz), N_(hmul)(qy, qw)), K(2.)), r1 = N_(hscale)(N_(hsub)(N_(hmul)(qy, qz), N_(hmul)(qx, qw)), K(2.));
    N_(hd) r2 = N_(hsub)(N_(H_)(K(1.)), N_(hscale)(N_(hadd)(N_(hmul)(qx, qx), N_(hmul)(qy, qy)), K(2.)));
    N_(hd) o0 = x[10], o1 = x[11], o2 = x[12];
    N_(hd) d0 = N_(hscale)(N_(hadd)(N_(hmul)(qw, o0), N_(hsub)(N_(hmul)(qy, o2), N_(hmul)(qz, o1))), K(0.5));
    N_(hd) d1 = N_(hscale)(N_(hadd)(N_(hmul)(qw, o1), N_(hsub)(N_(hmul)(qz, o0), N_(hmul)(qx, o2))), K(0.5));
    N_(hd) d2 = N_(hscale)(N_(hadd)(N_(hmul)(qw, o2), N_(hsub)(N_(hmul)(qx, o1), N_(hmul)(qy, o0))), K(0.5));
    N_(hd) d3 = N_(hscale)(N_(hadd)(N_(hadd)(N_(hmul)(qx, o0), N_(hmul)(qy, o1)), N_(hmul)(qz, o2)), -K(0.5));
    REAL im = K(1.) / c->m, dt = c->dt;
    x[0] = N_(hadd)(x[0], N_(hscale)(x[7], dt)); x[1] = N_(hadd)(x[1], N_(hscale)(x[8], dt)); x[2] = N_(hadd)(x[2], N_(hscale)(x[9], dt));
    x[7] = N_(hadd)(x[7], N_(hscale)(N_(haddf)(N_(hmul)(r0, thrust), fd[0]), im * dt));
    x[8] = N_(hadd)(x[8], N_(hscale)(N_(haddf)(N_(hmul)(r1, thrust), fd[1]), im * dt));
    x[9] = N_(hadd)(x[9], N_(hscale)(N_(haddf)(N_(hscale)(N_(haddf)(N_(hmul)(r2, thrust), fd[2]), im), -c->g), dt));
    N_(hd) n0 = N_(hadd)(qx, N_(hscale)(d0, dt)), n1 = N_(hadd)(qy, N_(hscale)(d1, dt)), n2 = N_(hadd)(qz, N_(hscale)(d2, dt)), n3 = N_(hadd)(qw, N_(hscale)(d3, dt));
    N_(hd) nn = N_(hrecip)(N_(hsqrt)(N_(hadd)(N_(hadd)(N_(hmul)(n0, n0), N_(hmul)(n1, n1)), N_(hadd)(N_(hmul)(n2, n2), N_(hmul)(n3, n3)))));
    x[3] = N_(hmul)(n0, nn); x[4] = N_(hmul)(n1, nn); x[5] = N_(hmul)(n2, nn); x[6] = N_(hmul)(n3, nn);
    x[10] = N_(hadd)(N_(hscale)(o0, c->alpha), N_(hscale)(w0, K(1.) - c->alpha));
    x[11] = N_(hadd)(N_(hscale)(o1, c->alpha), N_(hscale)(w1, K(1.) - c->alpha));
    x[12] = N_(hadd)(N_(hscale)(o2, c->alpha), N_(hscale)(w2, K(1.) - c->alpha));
}

/* R [n][n], n = 4H.  a_mean [H][4] (already shifted). */
void N_(oracle_hessian_fof)(const float* envp, const float* state24, int time, const float* pos_traj, const float* vel_traj,
                        int T, const float* a_mean, int H, REAL* R) {
    N_(env_t) c;
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
        N_(hd) x[13];
        for (int k = 0; k < 13; ++k) x[k] = N_(H_)(state24[k]);
        REAL fd[3] = {state24[13], state24[14], state24[15]};
        REAL pt[3] = {state24[16], state24[17], state24[18]}, vt[3] = {state24[19], state24[20], state24[21]};
        N_(hd) total = N_(H_)(K(0.));
        for (int h = 0; h < H; ++h) {
            total = N_(hadd)(total, N_(hreward)(x, pt, vt));
            N_(hd) u[4];
            for (int k = 0; k < 4; ++k) {
                int idx = 4 * h + k;
                u[k] = N_(H_)(a_mean[idx]);
                if (idx == i) u[k].a = K(1.);
                if (idx == j) u[k].b = K(1.);
            }
            N_(hstep)(x, u, fd, &c);
            fd[0] = fd[1] = fd[2] = K(0.);
            int row = time + h + 1;
            if (row > T - 1) row = T - 1;
            for (int k = 0; k < 3; ++k) { pt[k] = pos_traj[row * 3 + k]; vt[k] = vel_traj[row * 3 + k]; }
        }
        REAL val = -total.ab;
        R[(size_t)i * n + j] = val;
        R[(size_t)j * n + i] = val;
    }
}

