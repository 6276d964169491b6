#!/usr/bin/env python
"""Extracts one of the reference's example cases into tmp_cases/<name>, shortens the run to N steps, optionally
decomposes it (strips along x) and runs the UNMODIFIED reference binary on it (CPU); the time directory it writes is the
expected result of scripts/run_example_dropin.py.  Needs /root/reference (this container only).

    python scripts/prepare_example.py tandem_vortex 10 2 [key=value ...]     # key=value: fvSchemes/controlDict overrides
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lfm_public_b200.tools import foamcase  # noqa: E402


def main():
    name, steps, ranks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    overrides = dict(kv.split("=") for kv in sys.argv[4:])
    src = os.path.join("/root/reference/examples", name)
    dst = os.path.join(ROOT, "tmp_cases", name)
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    subprocess.check_call(["chmod", "-R", "u+w", dst])
    for f in glob.glob(os.path.join(dst, "**", "*.tgz"), recursive=True):
        subprocess.check_call(["tar", "xzf", os.path.basename(f)], cwd=os.path.dirname(f))
        os.remove(f)
    for f in glob.glob(os.path.join(dst, "**", "*.tgz"), recursive=True):     # faces.tgz.tgz holds faces.tgz
        subprocess.check_call(["tar", "xzf", os.path.basename(f)], cwd=os.path.dirname(f))
        os.remove(f)
    cd = os.path.join(dst, "system", "controlDict")
    s = open(cd).read()
    dt = float(re.search(r"^deltaT\s+([0-9.eE+-]+)", s, flags=re.M).group(1))
    s = re.sub(r"^endTime\s+[^;]+;", f"endTime         {dt * steps!r};", s, flags=re.M)
    s = re.sub(r"^writeInterval\s+[^;]+;", f"writeInterval   {steps};", s, flags=re.M)
    s = re.sub(r"^writePrecision\s+[^;]+;", "writePrecision  17;", s, flags=re.M)
    for k, v in overrides.items():
        s = re.sub(rf"(\b{k}\s+)[^;]+;", rf"\g<1>{v};", s)
    open(cd, "w").write(s)
    fs = os.path.join(dst, "system", "fvSchemes")
    s = open(fs).read()
    for k, v in overrides.items():
        s = re.sub(rf"(\b{k}\s+)[^;]+;", rf"\g<1>{v};", s)
    open(fs, "w").write(s)
    env = dict(os.environ, LFM_WRITE_PRECISION="17")
    args = [os.path.join(ROOT, "oracle", "_ref", "lfm_solve_ref")]
    if ranks > 1:
        t0 = time.time()
        foamcase.decompose_case(dst, ranks)
        print(f"decomposed into {ranks} ranks in {time.time() - t0:.0f} s", flush=True)
        env["LFM_MPI_NP"] = str(ranks)
        args.append("-p")
    t0 = time.time()
    out = subprocess.run(args, cwd=dst, env=env, capture_output=True, text=True)
    open(os.path.join(dst, "log.ref"), "w").write(out.stdout + out.stderr)
    print(out.stdout[-800:])
    print(f"reference run: {time.time() - t0:.0f} s; end time {dt * steps!r}")
    shutil.rmtree(os.path.join(dst, "output"), ignore_errors=True)


if __name__ == "__main__":
    main()
