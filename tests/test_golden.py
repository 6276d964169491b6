"""The CPU restatement (oracle/lfm_oracle.c) against the committed golden vectors, which were produced by the
reference's own CPU solver (tests/golden/make_golden.py).  Runs without /root/reference."""
import hashlib
import os

import numpy as np
import pytest

import common
import oracle_lib
from common import N_STEPS

GOLDEN = sorted(f[:-4] for f in os.listdir(common.GOLDEN_DIR) if f.endswith(".npz") and not f.endswith("_sp.npz"))
GOLDEN_SP = sorted(f[:-7] for f in os.listdir(common.GOLDEN_DIR) if f.endswith("_sp.npz"))


def load(name, sp=False):
    return np.load(os.path.join(common.GOLDEN_DIR, name + ("_sp" if sp else "") + ".npz"))


def test_golden_present():
    assert len(GOLDEN) >= 9 and len(GOLDEN_SP) >= 3


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden_fp64(name, tmp_path):
    g = load(name)
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir)
    cases = common.open_ranks(case_dir, o)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    record = {}
    oracle_lib.lockstep_run(oracles, cases, o["solver"], o["deltaT"], N_STEPS, record=record)
    assert int(g["n_ranks"]) == len(cases)
    for r, (c, orc) in enumerate(zip(cases, oracles)):
        mine = common.primitives_from_q(c.to_mesh_order(orc.download(0)), c.desc.c.gamma_m1)
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(mine[k], g[f"r{r}_{k}"]), f"{name} rank {r} {k}"
    for (src, dst), msgs in record.items():
        stream = np.concatenate(msgs).tobytes()
        if f"halo_{src}_{dst}" in g:
            assert stream == g[f"halo_{src}_{dst}"].tobytes(), f"{name} halo {src}->{dst}"
        else:                                      # large cases keep a digest of the stream (make_golden.py HALO_DIGEST)
            assert hashlib.sha256(stream).digest() == g[f"halo_{src}_{dst}_sha256"].tobytes(), f"{name} halo {src}->{dst}"
            assert int(g[f"halo_{src}_{dst}_n"].sum()) * 8 == len(stream)


@pytest.mark.parametrize("name", GOLDEN_SP)
def test_oracle_matches_golden_fp32(name, tmp_path):
    g = load(name, sp=True)
    case_dir = str(tmp_path / name)
    m, o = common.build_case(name, case_dir, doublePrecision=False)
    cases = common.open_ranks(case_dir, o)
    oracles = [oracle_lib.Oracle(c) for c in cases]
    oracle_lib.run(oracles, o["solver"], o["deltaT"], N_STEPS)
    D = o["dimension"]
    for r, (c, orc) in enumerate(zip(cases, oracles)):
        q = c.to_mesh_order(orc.download(0)).astype(np.float64)
        assert common.rel_max(q[:, 0], g[f"r{r}_rho"]) <= 1e-5
        assert common.rel_max(q[:, 1:D + 1] / q[:, :1], g[f"r{r}_U"]) <= 1e-5
        assert common.rel_max(q[:, D + 1] / q[:, 0], g[f"r{r}_E"]) <= 1e-5
