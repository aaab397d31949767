// TEST INFRASTRUCTURE: runs the kernels of covo_mpc_b200/csrc/sigma_dense.cu on the CPU stand-in of tests/emu/cuda_runtime.h
// (g++ -DCOVO_CPU_EMU -Itests/emu) -- one environment, the same launch geometry and shared-memory sizes as launch_sigma_dense.
#include "../../covo_mpc_b200/csrc/sigma_dense.cu"

extern "C" int emu_sigma_dense(int n, float sample_sigma, const float* R, const double* zolo, float* cov, double* scal4, int* status, int variant) {
    using namespace covo;
    if (n > kSigmaMaxN || (n & 3)) return 1;
    std::vector<float> xbuf((size_t)kZoloPoles * n * n, std::nanf(""));  // NaN: an entry read before it was written poisons the result
    DenseArgs a;
    a.n = n;
    a.n_pad = round_up8(n);
    a.sample_sigma = sample_sigma;
    a.R = R;
    a.scal = scal4;
    a.Xbuf = xbuf.data();
    a.cov = cov;
    a.zolo = zolo;
    a.status = status;
    a.prof = nullptr;
    a.Asym = nullptr;
    const size_t smem1 = (size_t)(2 * n + 8 + 2 * kLanczosMax) * sizeof(double) + (size_t)n * (n + 1) * sizeof(float);
    if (variant & 32) emu_launch_cluster(lanczos_cluster_kernel, dim3(LC_CL, 1), LC_CL, LC_T, sizeof(LcSmem), a);  // 8-CTA cluster
    else if ((variant & 16) || !lanczos2_layout(n).fits) emu_launch(lanczos_kernel, dim3(1), TL, smem1, a);  // the first Lanczos kernel
    else emu_launch(lanczos2_kernel, dim3(1), TL2, lanczos2_layout(n).bytes, a);
    variant &= 15;
    if (variant == 3) {
        emu_launch_cluster(gjb_inverse_kernel, dim3(GB_CL * (kZoloPoles + 1), 1), GB_CL, GB_T, sizeof(GjbSmem), a);
    } else if (variant == 2) {
        emu_launch(gj_inverse_kernel<14>, dim3(kZoloPoles + 1, 1), TG, 1024 * sizeof(float), a);
    } else {
        const size_t smem2 = ((size_t)n * n + 2 * 8 * a.n_pad + 64 + a.n_pad) * sizeof(float) + 16;
        emu_launch(shifted_inverse_kernel, dim3(kZoloPoles + 1, 1), TD, smem2, a);
    }
    const int npairs = n * (n + 1) / 2;
    emu_launch(combine_kernel, dim3((npairs + 255) / 256, 1), 256, 0, a);
    return 0;
}
