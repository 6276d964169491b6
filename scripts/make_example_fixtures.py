#!/usr/bin/env python
"""Fixtures of the reference's own example cases (BASELINE.json configs 1 and 2) for the GPU suite.  Needs /root/reference
(this container only).  For each case:

  1. scripts/prepare_example.py extracts the example, shortens it to N_STEPS time steps and runs the UNMODIFIED reference
     executable (oracle/_ref/lfm_solve_ref) on it;
  2. the fields it wrote (rho, U, E, p: 17 digits, exact doubles) are hashed into tests/golden/examples.json (committed);
  3. the case's INPUTS (polyMesh, 0/, dictionaries; converted to OpenFOAM's binary stream format, which both the reference's
     reader shim and liblfmhost.so read) are packed into tmp_cases/<case>.txz -- git-ignored (tens of MB of mesh data that is
     not ours to commit), NOT gpurun-ignored, so it travels to the GPU box with the snapshot.

tests/test_examples_gpu.py runs the CUDA path on every archive it finds and compares the hashes.

    python scripts/make_example_fixtures.py [case ...]
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tarfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from lfm_public_b200.tools import foamcase  # noqa: E402

N_STEPS = 3
CASES = ["cylinder_vortex", "cylinder_vortex_unstructured"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def main():
    names = sys.argv[1:] or CASES
    path = os.path.join(ROOT, "tests", "golden", "examples.json")
    fix = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "prepare_example.py"), name, str(N_STEPS), "1"])
        case = os.path.join(ROOT, "tmp_cases", name)
        cd = open(os.path.join(case, "system", "controlDict")).read()
        fs = open(os.path.join(case, "system", "fvSchemes")).read()
        dt = float(re.search(r"^deltaT\s+([0-9.eE+-]+)", cd, flags=re.M).group(1))
        D = int(re.search(r"\bdimension\s+(\d)", fs).group(1))
        scheme = int(re.search(r"\bsolver\s+(\d)", fs).group(1))
        tname = [d for d in os.listdir(case) if re.fullmatch(r"[0-9.eE+-]+", d) and d != "0"]
        assert len(tname) == 1, tname
        td = os.path.join(case, tname[0])
        f = dict(rho=common.read_field(os.path.join(td, "rho")), U=common.read_field(os.path.join(td, "U"), 3)[:, :D],
                 E=common.read_field(os.path.join(td, "E")), p=common.read_field(os.path.join(td, "p")))
        fix[name] = {"n_steps": N_STEPS, "deltaT": dt, "dimension": D, "solver": scheme, "n_cells": int(len(f["rho"])),
                     "sha256": {k: sha(v) for k, v in f.items()},
                     "source": f"oracle/_ref/lfm_solve_ref on /root/reference/examples/{name}, {N_STEPS} steps, fields of time {tname[0]}"}
        # inputs only, binary stream format, packed
        stage = os.path.join(ROOT, "tmp_cases", "_pack", name)
        shutil.rmtree(os.path.dirname(stage), ignore_errors=True)
        os.makedirs(stage)
        for sub in ("0", "constant", "system"):
            shutil.copytree(os.path.join(case, sub), os.path.join(stage, sub))
        foamcase.to_binary(stage)
        arc = os.path.join(ROOT, "tmp_cases", name + ".txz")
        with tarfile.open(arc, "w:xz") as t:
            t.add(stage, arcname=name)
        shutil.rmtree(os.path.dirname(stage), ignore_errors=True)
        shutil.rmtree(case, ignore_errors=True)
        print(f"{name}: {fix[name]['n_cells']} cells, archive {os.path.getsize(arc) / 1e6:.1f} MB")
    json.dump(fix, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
