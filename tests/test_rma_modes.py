"""The reference binary in its one-sided halo modes (haloCommType 3 = blocking, 4 = non-blocking: MPI_Win_post / start /
complete / wait + MPI_Get, /root/reference/src/mpi_env.cpp:114-137, 467-511 and cfd_v0.cpp:3284-3545) over the mini-MPI
shim's emulation of general active target synchronisation (oracle/shim/mpi_shim.cpp).  The transport must not change a
result: fields after N_STEPS steps are bit-identical to the two-sided non-blocking run of the same case, for the packed
(commType 1) and the split (commType 2) exchange, in 2D and 3D, laminar and LES (tauMC travels in the packed message)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.skipif(not common.have_ref(), reason="reference binary not built")


@pytest.mark.parametrize("name", ["quad2d_m1_p4", "hex3d_m2_p4", "hex3d_m2_les_p4", "hex3d_m2_pc8", "hex3d_ausm_p4"])
@pytest.mark.parametrize("halo,comm", [(3, 1), (3, 2), (4, 1), (4, 2)])
def test_one_sided_modes_equal_two_sided(name, halo, comm, tmp_path):
    fields = {}
    for tag, (h, c) in (("two_sided", (2, comm)), ("one_sided", (halo, comm))):
        d = str(tmp_path / tag)
        m, o = common.build_case(name, d, haloCommType=h, commType=c)
        common.run_reference(d, o, dump=False)
        D = o["dimension"]
        fields[tag] = common.read_reference_q(d, o, o["deltaT"] * common.N_STEPS, D)
        n_ranks = o["n_ranks"]
    for r in range(n_ranks):
        for k in ("rho", "U", "E", "p"):
            a, b = fields["two_sided"][r][k], fields["one_sided"][r][k]
            assert np.isfinite(a).all() and np.array_equal(a, b), f"{name} halo {halo} comm {comm} rank {r} field {k}: rel max {common.rel_max(a, b):.3e}"


@pytest.mark.skipif(not common.have_ref(sp=True), reason="single-precision reference binary not built")
@pytest.mark.parametrize("halo,comm", [(3, 2), (4, 1)])
def test_one_sided_modes_equal_two_sided_fp32(halo, comm, tmp_path):
    """The single-precision build moves float payloads through the same windows (MPI_CHAR gets of float buffers)."""
    fields = {}
    for tag, h in (("two_sided", 2), ("one_sided", halo)):
        d = str(tmp_path / tag)
        m, o = common.build_case("hex3d_m2_p4", d, haloCommType=h, commType=comm)
        common.run_reference(d, o, sp=True, dump=False)
        fields[tag] = common.read_reference_q(d, o, o["deltaT"] * common.N_STEPS, o["dimension"])
        n_ranks = o["n_ranks"]
    for r in range(n_ranks):
        for k in ("rho", "U", "E", "p"):
            assert np.array_equal(fields["two_sided"][r][k], fields["one_sided"][r][k]), f"fp32 halo {halo} comm {comm} rank {r} field {k}"
