#!/usr/bin/env python
"""Config-1/2 parity on a REAL example case of the reference: runs oracle/_ref/lfm_solve_gpu (the reference's own main()
and Mesh::solve with CFDv0_solver_gpu) on a copy of tmp_cases/<case> (extracted from /root/reference/examples/<case>,
already advanced by the unmodified reference binary on the CPU: the time directory it wrote is the expected result) and
compares every field of the final time directory.

    python scripts/run_example_dropin.py cylinder_vortex 0.06 [n_ranks]
"""
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402


def main():
    case, tname = sys.argv[1], sys.argv[2]
    ranks = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    src = os.path.join(ROOT, "tmp_cases", case)
    if not os.path.isdir(src):                      # shipped compressed (the gpurun snapshot is size-limited)
        subprocess.check_call(["tar", "xzf", src + ".tgz", "-C", "/tmp"])
        src = os.path.join("/tmp", case)
    dst = os.path.join("/tmp", case + "_gpu")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(tname, "output", "log.*"))
    env = dict(os.environ, LFM_WRITE_PRECISION="17")
    args = [os.path.join(ROOT, "oracle", "_ref", "lfm_solve_gpu")]
    if ranks > 1:
        env["LFM_MPI_NP"] = str(ranks)
        args.append("-p")
    t0 = time.time()
    out = subprocess.run(args, cwd=dst, env=env, capture_output=True, text=True, timeout=3000)
    wall = time.time() - t0
    ok = "Simulation finished successfully" in out.stdout
    print(out.stdout[-1500:])
    print("\n".join(l for l in out.stderr.splitlines() if "lfmgpu" in l)[:2000])
    if not ok:
        print(out.stderr[-3000:])
        raise SystemExit("drop-in run failed")
    rk = re.search(r"\[\s*([0-9.eE+-]+)\]: Rk Loop", out.stdout)
    print(f"# {case}: drop-in run wall {wall:.1f} s, Rk Loop {rk.group(1) if rk else '?'} s")
    worst = 0.0
    subdirs = [f"processor{r}" for r in range(ranks)] if ranks > 1 else [""]
    pairs = [(sd, name) for sd in subdirs for name in sorted(os.listdir(os.path.join(src, sd, tname)))]
    for sd, name in pairs:
        a_path, b_path = os.path.join(src, sd, tname, name), os.path.join(dst, sd, tname, name)
        if not os.path.isfile(a_path) or not os.path.exists(b_path):
            continue
        try:
            ncomp = 3 if name == "U" else 1
            a, b = common.read_field(a_path, ncomp), common.read_field(b_path, ncomp)
        except Exception as e:
            print(f"  {name}: skipped ({e})")
            continue
        rel = common.rel_max(a, b)
        worst = max(worst, rel)
        print(f"  {sd:11s} field {name:8s} n={a.shape[0]:8d} identical={bool(np.array_equal(a, b))} rel_max={rel:.3e} (max |ref| {np.abs(a).max():.6g})")
    print(f"# worst relative max-norm difference: {worst:.3e} (bar 1e-12)")
    if worst > 1e-12:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
