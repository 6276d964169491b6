"""lfm_public_b200 -- B200 (sm_100a) implementation of libFastMesh's per-iteration solve.

Only what the hot path needs lives here:
  csrc/       CUDA kernels + the C ABI (include/lfmgpu.h)            -> liblfmgpu.so
  host/       C++ host: gpu_solver.h = CFDv0_solver_gpu, the reference-side binding;
              OpenFOAM-free case I/O and the reference's pre-loop setup restated (include/lfmhost.h) -> liblfmhost.so
  host_api.py / gpu_api.py   thin ctypes views of the two libraries (tests, bench)
  tools/      synthetic polyMesh generators / decomposer / case writer, OpenFOAM case reader, renumbering tool,
              tuning sweep (test + bench tooling)
"""
from . import _ctypes_defs as defs  # noqa: F401

__all__ = ["defs"]
