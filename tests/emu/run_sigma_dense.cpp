// TEST INFRASTRUCTURE: runs the kernels of covo_mpc_b200/csrc/sigma_dense.cu on the CPU stand-in of tests/emu/cuda_runtime.h
// (g++ -DCOVO_CPU_EMU -Itests/emu) -- one environment, the same launch geometry and shared-memory sizes as launch_sigma_dense.
#include "../../covo_mpc_b200/csrc/sigma_dense.cu"

extern "C" int emu_sigma_dense(int n, float sample_sigma, const float* R, const double* zolo, float* cov, double* scal4, int* status, int variant) {
    using namespace covo;
    if (n > kSigmaMaxN || (n & 3)) return 1;
    std::vector<float> xbuf((size_t)kDensePoles * n * n, std::nanf(""));  // NaN: an entry read before it was written poisons the result
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
    emu_launch_cluster(lanczos_cluster_kernel, dim3(LC_CL, 1), LC_CL, LC_TT, sizeof(LcSmem), a);  // 8-CTA cluster, 4 + 1 warps
    if (variant & 64) return 0;  // lambda_min only (studies of the Lanczos stage)
    if (variant & 16) emu_launch_cluster(gjb_inverse_kernel_t<16>, dim3(GB_CL * (kDensePoles + 1), 1), GB_CL, GB_T, sizeof(GjbSmemT<16>), a);  // 16 pivots per step
    else emu_launch_cluster(gjb_inverse_kernel_t<8>, dim3(GB_CL * (kDensePoles + 1), 1), GB_CL, GB_T, sizeof(GjbSmemT<8>), a);
    if (const char* dump = getenv("COVO_EMU_DUMP_X")) {  // development: the per-pole inverses
        FILE* f = fopen(dump, "wb");
        if (f) {
            fwrite(xbuf.data(), sizeof(float), xbuf.size(), f);
            fclose(f);
        }
    }
    const int npairs = n * (n + 1) / 2;
    emu_launch(combine_kernel, dim3((npairs + 255) / 256, 1), 256, 0, a);
    return 0;
}
