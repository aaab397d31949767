"""Per-kernel CUDA-event times of one CoVO-online step for each optimize_sigma variant (COVO_SIGMA unset / dense / dense-gj).
Development tool: `python tools/time_sigma_variants.py [N] [H]` on a GPU box (each variant runs in a fresh process)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np

    import bench
    from covo_mpc_b200 import _lib

    N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    bench.N_SAMPLES, bench.HORIZON = N, H
    env, states, times, traj = bench.record_states(24, 100, "covo-online", device=0)
    cfg = _lib.default_config()
    cfg.mode, cfg.n_samples, cfg.horizon, cfg.traj_len, cfg.device = _lib.MODE_COVO_ONLINE, N, H, int(traj[0].shape[0]), 0
    h = _lib.Handle(cfg)
    h.set_reference(traj[0][None], traj[1][None])
    for i in range(4):
        h.step(states[i], times[i:i + 1])
    h.set_profiling(True)
    acc = np.zeros(6)
    for i in range(4, 20):
        h.step(states[i], times[i:i + 1])
        acc += h.kernel_ms()
    names = ["hessian", "E1|dense", "E2", "E3", "cholesky", "rollout"]
    print(os.environ.get("COVO_SIGMA", "tridiag"), {k: round(float(v) / 16 * 1e3, 1) for k, v in zip(names, acc)}, "us; status", h.status())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for v in ("", "dense-gj", "dense-gjb"):
            env = dict(os.environ)
            if v:
                env["COVO_SIGMA"] = v
            else:
                env.pop("COVO_SIGMA", None)
            r = subprocess.run([sys.executable, __file__, "child", *sys.argv[1:]], env=env, capture_output=True, text=True, timeout=300)
            print(r.stdout.strip() or r.stderr[-2000:])
