#!/usr/bin/env python
"""Writes profiles/ncu_traffic.json from a full ncu capture of the dominant kernels: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), stamped with the SHA-256 of the device code (.nv_fatbin section) of the library
that was profiled so that bench.py reports `roofline.traffic` only for those kernels.

    python scripts/ncu_traffic.py <stage.ncu-rep> <grad.ncu-rep> <n> <numbering> [f64|f32]
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def traffic(rep, kernel=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, units = r[0], r[1]
    tot, dur = [], []
    for row in r[2:]:
        if kernel and kernel not in row[h.index("Kernel Name")]:
            continue   # one capture may hold both kernels
        b = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = float(row[h.index(k)]), units[h.index(k)]
            b += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tot.append(b)
        dur.append(float(row[h.index("gpu__time_duration.sum")]))
    return sum(tot) / len(tot), sum(dur) / len(dur), units[h.index("gpu__time_duration.sum")]


def main():
    stage, grad, n, numbering = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    prec = sys.argv[5] if len(sys.argv) > 5 else "f64"
    sys.path.insert(0, ROOT)
    from lfm_public_b200.tools.libstamp import device_code_sha256
    lib = os.path.join(ROOT, "lfm_public_b200", "liblfmgpu.so")
    so = hashlib.sha256(open(lib, "rb").read()).hexdigest()
    fat = device_code_sha256(lib)   # what bench.py compares: the device code (the file hash changes from build to build)
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name, rep in (("tile_stage", stage), ("tile_grad", grad)):
        b, d, u = traffic(rep, {"tile_stage": "k_stage_pipe", "tile_grad": "k_grad_pipe"}[name] if stage == grad else None)
        out.setdefault(name, {})[prec] = {"n": n, "numbering": numbering, "bytes_per_launch": b, "ncu_duration": d, "ncu_duration_unit": u,
                                          "lib_sha256": so, "fatbin_sha256": fat, "capture": os.path.basename(rep)}
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
