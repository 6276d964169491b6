#!/usr/bin/env python
"""Runs the reference's own hpathRenumber plugin (oracle/_ref/hpath_plugin: hpathRenumber.C compiled unchanged against
oracle/shim/openfoam_stub/) on the meshes of the reference's example cases and records the SHA-256 of the order it returns in
tests/golden/hpath_examples.json.  Needs /root/reference (this container only); tests/test_hpath_vs_plugin.py compares the
restatement (lfm_public_b200/host/hpath.cpp) with these fixtures wherever the example meshes can be read.

    python scripts/make_hpath_fixtures.py
"""
import glob
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXAMPLES = ["cylinder_vortex", "cylinder_vortex_unstructured"]
PLUGIN = os.path.join(ROOT, "oracle", "_ref", "hpath_plugin")


def extract(name, dst_root):
    """A writable copy of /root/reference/examples/<name> with its mesh archives unpacked."""
    dst = os.path.join(dst_root, name)
    shutil.copytree(os.path.join("/root/reference/examples", name), dst)
    subprocess.check_call(["chmod", "-R", "u+w", dst])
    for _ in range(2):   # faces.tgz.tgz holds faces.tgz
        for f in glob.glob(os.path.join(dst, "**", "*.tgz"), recursive=True):
            subprocess.check_call(["tar", "xzf", os.path.basename(f)], cwd=os.path.dirname(f))
            os.remove(f)
    return dst


def plugin_order(poly_dir, out_path):
    r = subprocess.run([PLUGIN, poly_dir, out_path], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("hpath_plugin failed: " + r.stderr[-2000:])
    return np.fromfile(out_path, dtype=np.int32)


def main():
    fix = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in EXAMPLES:
            case = extract(name, tmp)
            order = plugin_order(os.path.join(case, "constant", "polyMesh"), os.path.join(tmp, name + ".bin"))
            fix[name] = {"n_cells": int(len(order)), "sha256": hashlib.sha256(order.tobytes()).hexdigest()}
            print(name, fix[name])
    json.dump(fix, open(os.path.join(ROOT, "tests", "golden", "hpath_examples.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
