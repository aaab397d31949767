"""SASS mnemonic census of the product library -> profiles/<tag>_sass_mnemonics.md (which hardware features the kernels really use).
Runs where cuobjdump is installed (no GPU needed):  python tools/sass_census.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "FFMA2", "FFMA", "DFMA", "STAS", "UCGABAR", "SYNCS",
        "LDGSTS", "ACQBULK", "MEMBAR", "BAR", "NANOSLEEP", "MUFU", "F2F"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    lib = os.path.join(ROOT, "covo_mpc_b200", "libcovo_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    rows, tot = [], collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter()
        for line in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            base = re.split(r"[._]", op)[0]
            for p in PATS:
                if base == p:
                    c[p] += 1
        rows.append((name, c))
        tot.update(c)

    def dem(n):
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("covo::", "").replace("void ", "")

    out = [f"# SASS mnemonic census of covo_mpc_b200/libcovo_b200.so (sm_100a), {tag}", "",
           "`cuobjdump -sass covo_mpc_b200/libcovo_b200.so`, counted per kernel (tools/sass_census.py).  What the path uses: packed FP32",
           "(`FFMA2`), FP64 (`DFMA`) for the tridiagonal matrix function / Lanczos, 1-D TMA bulk copies (`UBLKCP`, global->shared and",
           "shared->distributed-shared) with mbarriers (`SYNCS`), distributed-shared-memory stores that carry their own completion",
           "(`STAS` = st.async), cluster barriers (`UCGABAR_ARV/_WAIT`), `cp.async` (`LDGSTS`).  What it does NOT use: tcgen05 tensor cores",
           "(`UTC*MMA`, `LDTM`/`STTM`) and tensor-map TMA (`UTMALDG`/`UTMASTG`): the step is a chain of latency-bound small-matrix kernels",
           "(DESIGN.md section 4); section 9 names the contractions that belong on tensor cores in the throughput configurations.", "",
           "| kernel | " + " | ".join(PATS) + " |", "|---|" + "---|" * len(PATS)]
    for n, c in rows:
        if sum(c.values()):
            out.append("| " + dem(n) + " | " + " | ".join(str(c[p]) if c[p] else "" for p in PATS) + " |")
    out.append("| **total** | " + " | ".join(str(tot[p]) for p in PATS) + " |")
    open(os.path.join(ROOT, "profiles", f"{tag}_sass_mnemonics.md"), "w").write("\n".join(out) + "\n")
    print(out[-1])


if __name__ == "__main__":
    main()
