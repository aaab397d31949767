// TEST INFRASTRUCTURE ONLY: compiles the product's __host__ __device__ model headers with g++ so the
// dynamics / reward / forward-AD algebra / PID policy can be checked against oracle/oracle_np.py on a
// machine without a GPU.  Nothing in covo_mpc_b200 loads this library; the product has no CPU path.
#include <cstring>

#include "../../covo_mpc_b200/csrc/hessian_local.cuh"
#include "../../covo_mpc_b200/csrc/pid.cuh"

using namespace covo;

static EnvConsts make_env(const float* e) {
    EnvConsts c;
    c.m = e[0]; c.g = e[1]; c.max_thrust = e[2]; c.dt = e[3]; c.alpha_bodyrate = e[4]; c.action_scale = e[5];
    c.pos_limit = e[6]; c.max_omega[0] = e[7]; c.max_omega[1] = e[8]; c.max_omega[2] = e[9]; c.max_steps = (int)e[10];
    return c;
}

extern "C" {

// one float transition: x[13] -> x_next[13]; reward and done of the PRE-step state
void hc_step(const float* envp, const float* x, const float* u, const float* fd, const float* pt, const float* vt, int time,
             float* x_next, float* reward, int* done) {
    EnvConsts c = make_env(envp);
    QState<float> s;
    for (int k = 0; k < 3; ++k) s.p[k] = x[k];
    for (int k = 0; k < 4; ++k) s.q[k] = x[3 + k];
    for (int k = 0; k < 3; ++k) s.v[k] = x[7 + k];
    for (int k = 0; k < 3; ++k) s.w[k] = x[10 + k];
    *reward = quad_reward(s, pt, vt);
    *done = quad_terminal(s, time, c) ? 1 : 0;
    quad_step(s, u, fd, c);
    for (int k = 0; k < 3; ++k) x_next[k] = s.p[k];
    for (int k = 0; k < 4; ++k) x_next[3 + k] = s.q[k];
    for (int k = 0; k < 3; ++k) x_next[7 + k] = s.v[k];
    for (int k = 0; k < 3; ++k) x_next[10 + k] = s.w[k];
}

// all local derivatives at (x, u): G[14][17] first derivatives (row 13 = cost), T[14][153] second
void hc_local(const float* envp, const float* x, const float* u, const float* fd, const float* pt, const float* vt,
              float* G, float* T) {
    EnvConsts c = make_env(envp);
    for (int pi = 0; pi < NPAIR; ++pi) {
        int a, b;
        pair_from_index(pi, a, b);
        float Fab[14], Fa[14];
        hess_local_task(x, u, fd, pt, vt, c, a, b, Fab, Fa);
        for (int k = 0; k < 14; ++k) T[k * NPAIR + pi] = Fab[k];
        if (a == b)
            for (int k = 0; k < 14; ++k) G[k * NZ + a] = Fa[k];
    }
}

void hc_pid(const float* envp, const float* x, const float* pt, const float* vt, const float* at, float* act) {
    EnvConsts c = make_env(envp);
    QState<float> s;
    for (int k = 0; k < 3; ++k) s.p[k] = x[k];
    for (int k = 0; k < 4; ++k) s.q[k] = x[3 + k];
    for (int k = 0; k < 3; ++k) s.v[k] = x[7 + k];
    for (int k = 0; k < 3; ++k) s.w[k] = x[10 + k];
    pid_action(s, pt, vt, at, c, c.max_thrust, 10.f, 5.f, 10.f, act);
}

int hc_pair_index(int a, int b) { return pair_index(a, b); }
}
