// PID expansion policy of CoVO-offline (shared by offline.cu and tests/host_check).
#pragma once
#include "quad_model.cuh"

namespace covo {

// qtoQ for an un-normalised quaternion = |q|^2 R(q/|q|)  (geom.py:68-77)
COVO_HD void qtoQ(const float q[4], float Q[3][3]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    Q[0][0] = w * w + x * x - y * y - z * z; Q[0][1] = 2.f * (x * y - w * z); Q[0][2] = 2.f * (x * z + w * y);
    Q[1][0] = 2.f * (x * y + w * z); Q[1][1] = w * w - x * x + y * y - z * z; Q[1][2] = 2.f * (y * z - w * x);
    Q[2][0] = 2.f * (x * z - w * y); Q[2][1] = 2.f * (y * z + w * x); Q[2][2] = w * w - x * x - y * y + z * z;
}

// controllers/pid.py:38-83 (Ki * integral term optional: the CoVO-offline expansion policy and get_controller("pid") use Ki = 0)
COVO_HD void pid_action(const QState<float>& s, const float ptar[3], const float vtar[3],
                                    const float atar[3], const EnvConsts& c, float max_thrust, float Kp, float Kd,
                                    float Kp_att, float act[4], float Ki = 0.f, const float* integ = nullptr) {
    float Q[3][3];
    qtoQ(s.q, Q);
    float fd[3];
    for (int k = 0; k < 3; ++k)
        fd[k] = c.m * (((k == 2) ? c.g : 0.f) - Kp * (s.p[k] - ptar[k]) - Kd * (s.v[k] - vtar[k]) - (integ ? Ki * integ[k] : 0.f) + atar[k]);
    float thrust = Q[0][2] * fd[0] + Q[1][2] * fd[1] + Q[2][2] * fd[2];
    thrust = fminf(fmaxf(thrust, 0.f), max_thrust);
    float nrm = sqrtf(fd[0] * fd[0] + fd[1] * fd[1] + fd[2] * fd[2]);
    if (nrm < 1e-3f) nrm = 1e-3f;
    float zd[3] = {fd[0] / nrm, fd[1] / nrm, fd[2] / nrm};
    float aa[3] = {-zd[1], zd[0], 0.f};  // e3 x z_d
    float angle = sqrtf(aa[0] * aa[0] + aa[1] * aa[1]);
    if (angle < 1e-3f) angle = 5e-4f;  // pid.py:59 (the test on :60 then never fires)
    float ax[3] = {aa[0] / angle, aa[1] / angle, 0.f};
    float an = sqrtf(ax[0] * ax[0] + ax[1] * ax[1]);
    float Rd[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    if (an > 0.f) {  // EXTENSION: the reference divides 0/0 here when f_d is exactly vertical
        float u[3] = {ax[0] / an, ax[1] / an, 0.f};
        float Hx[3][3] = {{0.f, -u[2], u[1]}, {u[2], 0.f, -u[0]}, {-u[1], u[0], 0.f}};
        float sa = sinf(angle), ca = 1.f - cosf(angle);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                float h2 = 0.f;
                for (int k = 0; k < 3; ++k) h2 += Hx[i][k] * Hx[k][j];
                Rd[i][j] += sa * Hx[i][j] + ca * h2;
            }
    }
    float Re[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc += Rd[k][i] * Q[k][j];
            Re[i][j] = acc;
        }
    float err[3] = {Re[2][1] - Re[1][2], Re[0][2] - Re[2][0], Re[1][0] - Re[0][1]};
    act[0] = thrust / max_thrust * 2.f - 1.f;
    for (int k = 0; k < 3; ++k) act[1 + k] = -Kp_att * err[k] / c.max_omega[k];
}

}  // namespace covo
