"""The reference's evaluation protocol (quadjax/envs/quadrotor.py:506-591; scripts/covo_quadrotor.sh, covo_quadrotor_N.sh)
through the drop-in controllers: PRNGKey(1), 4 reference trajectories x 10 episodes x 300 steps, metric = mean +- std over
episodes of the per-episode mean ||pos_tar - pos||, keys threaded exactly as eval_env threads them (harness.eval_env(keyed=True)).
Prints one JSON line per run and a summary line with CoVO's improvement over MPPI (the reference README:24 quotes 43-54 %).

    python tools/eval_table.py [--task tracking_zigzag] [--H 32] [--N 8192] [--episodes 40] [--disturb none] [--ablation]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import covo_mpc_b200 as cm  # noqa: E402


def run(task, name, N, H, lam, episodes, disturb):
    env = cm.Quad3D(task, disturb_type=disturb)
    ctl, _ = cm.get_controller(env, name, f"N{N}_H{H}_lam{lam}")
    t0 = time.time()
    mean, std, per_ep = cm.eval_env(env, ctl, total_steps=300 * episodes, num_trajs=4, seed=1, keyed=True)
    rec = {"task": task, "controller": name, "N": N, "H": H, "lam": lam, "disturb_type": disturb, "episodes": len(per_ep),
           "err_pos_mean_cm": round(100 * mean, 3), "err_pos_std_cm": round(100 * std, 3), "seconds": round(time.time() - t0, 1)}
    if hasattr(ctl, "close"):
        ctl.close()
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="tracking_zigzag")
    ap.add_argument("--H", type=int, default=32)
    ap.add_argument("--N", type=int, default=8192)
    ap.add_argument("--lam", type=float, default=0.01)
    ap.add_argument("--episodes", type=int, default=40)
    ap.add_argument("--disturb", default="none")
    ap.add_argument("--ablation", action="store_true", help="also the N ablation of scripts/covo_quadrotor_N.sh (subset)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    names = ("mppi", "covo-online", "covo-offline")
    recs = [run(a.task, c, a.N, a.H, a.lam, a.episodes, a.disturb) for c in names]
    base = recs[0]["err_pos_mean_cm"]
    summary = {"summary": "improvement over mppi = 1 - err/err_mppi",
               **{r["controller"]: round(1.0 - r["err_pos_mean_cm"] / base, 3) for r in recs[1:]},
               "reference_claim": "43 to 54 % (README.md:24; simulation + hardware, configuration of scripts/covo_quadrotor.sh)"}
    print(json.dumps(summary), flush=True)
    if a.ablation:
        for N in (16, 64, 256, 1024):
            for c in ("mppi", "covo-online"):
                recs.append(run(a.task, c, N, a.H, a.lam, a.episodes, a.disturb))
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"runs": recs, "summary": summary}, f, indent=1)


if __name__ == "__main__":
    main()
