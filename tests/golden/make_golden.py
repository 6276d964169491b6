"""Generates tests/golden/*.npz by running the REFERENCE's own CPU solver (oracle/_ref/lfm_solve_ref[_sp], i.e.
/root/reference/src/*.cpp + info/lfm_solve.cpp compiled unchanged by oracle/Makefile) on the case zoo of
tests/common.py.  Stored per case: the fields the reference writes after N_STEPS steps (rho, U, E, p per rank, in
polyMesh cell order, written with 17 significant digits => exact doubles) and, for decomposed cases, every halo
message it sent (concatenated per (src,dst) pair).  Needs /root/reference (to have built oracle/_ref); the
fixtures travel with the repo so that neither the CPU suite nor the GPU box needs the reference.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import common  # noqa: E402

GOLDEN_CASES = ["quad2d_m1", "quad2d_m1_p4", "quad2d_m2_p2_packed", "tri2d_m2", "hex3d_m2", "hex3d_m2_p4", "hex3d_m1_p8",
                "ogrid3d_m2", "ogrid2d_m1", "hex3d_m2_les_p4", "ogrid3d_m1_les", "hex3d_ausm_p4", "ogrid2d_ausm", "hex3d_m1_pc2", "hex3d_m2_pc8"]
GOLDEN_SP = ["quad2d_m1", "hex3d_m2_p4", "tri2d_m2"]
HALO_DIGEST = {"hex3d_m1_pc2", "hex3d_m2_pc8"}      # many long messages: the fixture keeps a SHA-256 per (src, dst) stream instead of the stream


def make(name, sp=False):
    with tempfile.TemporaryDirectory() as tmp:
        case_dir = os.path.join(tmp, name)
        m, o = common.build_case(name, case_dir, doublePrecision=not sp)
        common.run_reference(case_dir, o, sp=sp)
        D = o["dimension"]
        ref = common.read_reference_q(case_dir, o, o["deltaT"] * common.N_STEPS, D)
        out = dict(n_ranks=o["n_ranks"], n_steps=common.N_STEPS, D=D)
        for r, f in enumerate(ref):
            for k, v in f.items():
                out[f"r{r}_{k}"] = v.astype(np.float32) if sp else v
        if o["parallel"] and not sp:
            for src in range(o["n_ranks"]):
                for dst in range(o["n_ranks"]):
                    p = os.path.join(case_dir, "dump", f"send_r{src}_to{dst}.bin")
                    if os.path.exists(p):
                        msgs = common.read_dump(case_dir, src, dst, np.float64)
                        stream = np.concatenate([a for _, a in msgs])
                        if name in HALO_DIGEST:
                            out[f"halo_{src}_{dst}_sha256"] = np.frombuffer(hashlib.sha256(stream.tobytes()).digest(), dtype=np.uint8)
                        else:
                            out[f"halo_{src}_{dst}"] = stream
                        out[f"halo_{src}_{dst}_n"] = np.array([len(a) for _, a in msgs])
        path = os.path.join(HERE, name + ("_sp" if sp else "") + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    assert common.have_ref() and common.have_ref(sp=True), "build oracle/_ref first (make -C oracle ref)"
    only = sys.argv[1:]
    for n in GOLDEN_CASES:
        if not only or n in only:
            make(n)
    for n in GOLDEN_SP:
        if not only or n in only:
            make(n, sp=True)
