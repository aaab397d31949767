"""Summarise an `ncu --set full` report of one MPC step into profiles/: a markdown table and the per-kernel DRAM
traffic (profiles/ncu_traffic.json, read by bench.py for roofline.traffic).  Runs where ncu is installed (no GPU
needed):  python tools/summarise_ncu.py gpurun_out/step_full.ncu-rep r1"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
GROUP = {"hess_local": "hessian", "hess_assemble": "hessian", "hess_forward": "hessian", "tridiag_reg": "tridiag", "qacc": "qacc", "sigma_trifunc": "trifunc",
         "sandwich": "sandwich", "cholesky": "cholesky", "rollout": "rollout",
         "lanczos_cluster": "lanczos", "gjb_inverse": "pole_inverses", "combine": "combine"}


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    merge = len(sys.argv) > 3 and sys.argv[3] == "--merge-traffic"  # add the kernels only this capture has to profiles/ncu_traffic.json (other optimize_sigma path)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_raw.csv"), "w") as f:
        w = csv.writer(f)
        keep = [i for i, h in enumerate(hdr) if h in ("Kernel Name",) or h in KEYS]
        for r in rows:
            if len(r) == len(hdr):
                w.writerow([r[i] for i in keep])
    ci = {h: i for i, h in enumerate(hdr)}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = {}
    if merge and os.path.exists(tpath):
        traffic = json.load(open(tpath))
    known = set(traffic)  # merge mode: this capture only contributes the kernels the first one did not contain
    seen = set()
    lines = ["| kernel | " + " | ".join(k.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", "") for k in KEYS) + " |",
             "|" + "---|" * (len(KEYS) + 1)]
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("covo::", "")
        cells = []
        for k in KEYS:
            if k in ci:
                cells.append(f"{r[ci[k]]} {units[ci[k]]}".strip())
            else:
                cells.append("-")
        lines.append("| " + name + " | " + " | ".join(cells) + " |")
        for pat, g in GROUP.items():
            if pat in name:
                b = to_bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]]) + to_bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]])
                if merge and g in known:
                    continue
                traffic[g] = (traffic.get(g, 0) if g in seen else 0) + int(b)
                seen.add(g)
    if not merge:
        traffic["_source"] = (f"profiles/{tag}_ncu_full_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch (hessian = local + assemble + "
                              "forward); ncu --set full --clock-control none, one CoVO-online step at N=8192, H=50")
    json.dump(traffic, open(tpath, "w"))
    head = (f"# ncu --set full, one MPC step (CoVO-online, N=8192, H=50), {tag}\n\nCaptured with `ncu --set full --clock-control none "
            "--import-source on` around `tools/one_step.py` (fourth step, direct launches: COVO_GRAPH=0). Units as printed by ncu.\n\n")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write(head + "\n".join(lines) + "\n")
    print("\n".join(lines))
    print(traffic)


if __name__ == "__main__":
    main()
