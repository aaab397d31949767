// Per-transition derivative task of the CoVO Hessian (shared by hessian.cu and the host-side model
// check in tests/host_check): one hyper-dual evaluation of step + running cost for one (a <= b) pair
// of the 17 local inputs z = (x[13], u[4]).
#pragma once
#include "quad_model.cuh"

namespace covo {

constexpr int NZ = 17;                    // local inputs: 13 state + 4 control
constexpr int NPAIR = NZ * (NZ + 1) / 2;  // 153
constexpr int NX = 13;

COVO_HD void pair_from_index(int pi, int& a, int& b) {
    a = 0;
    int cnt = NZ;
    while (pi >= cnt) {
        pi -= cnt;
        --cnt;
        ++a;
    }
    b = a + pi;
}
COVO_HD int pair_index(int a, int b) {  // a <= b
    return a * NZ - (a * (a - 1)) / 2 + (b - a);
}

//   Fab[k] = d2 F_k / dz_a dz_b (k < 13),  Fab[13] = d2 c / dz_a dz_b     (c = -reward)
//   Fa[k]  = d F_k / dz_a,                 Fa[13]  = d c / dz_a
COVO_HD void hess_local_task(const float x[13], const float u[4], const float fd[3], const float pt[3],
                             const float vt[3], const EnvConsts& c, int a, int b, float Fab[14], float Fa[14]) {
    QState<HDual> s;
    HDual z[NZ];
    for (int i = 0; i < NZ; ++i) {
        float v = (i < NX) ? x[i] : u[i - NX];
        z[i] = HDual{v, (i == a) ? 1.f : 0.f, (i == b) ? 1.f : 0.f, 0.f};
    }
    for (int k = 0; k < 3; ++k) s.p[k] = z[k];
    for (int k = 0; k < 4; ++k) s.q[k] = z[3 + k];
    for (int k = 0; k < 3; ++k) s.v[k] = z[7 + k];
    for (int k = 0; k < 3; ++k) s.w[k] = z[10 + k];
    HDual r = quad_reward(s, pt, vt);
    Fab[13] = -r.ab;
    Fa[13] = -r.a;
    HDual uu[4] = {z[13], z[14], z[15], z[16]};
    quad_step(s, uu, fd, c);
    for (int k = 0; k < 3; ++k) { Fab[k] = s.p[k].ab; Fa[k] = s.p[k].a; }
    for (int k = 0; k < 4; ++k) { Fab[3 + k] = s.q[k].ab; Fa[3 + k] = s.q[k].a; }
    for (int k = 0; k < 3; ++k) { Fab[7 + k] = s.v[k].ab; Fa[7 + k] = s.v[k].a; }
    for (int k = 0; k < 3; ++k) { Fab[10 + k] = s.w[k].ab; Fa[10 + k] = s.w[k].a; }
}

}  // namespace covo
