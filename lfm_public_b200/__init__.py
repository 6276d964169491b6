"""lfm_public_b200 -- B200 (sm_100a) implementation of libFastMesh's per-iteration solve.

Only what the hot path needs lives here:
  csrc/       CUDA kernels + the C ABI (include/lfmgpu.h)            -> liblfmgpu.so
  host/       C++ host: OpenFOAM-free case I/O, the reference's pre-loop setup restated, ISolver/Mesh mirror
              (include/lfmhost.h)                                     -> liblfmhost.so, lfm_solve_gpu
  host_api.py / gpu_api.py   thin ctypes views of the two libraries (tests, bench)
  tools/      synthetic polyMesh generators / decomposer / case writer (test + bench tooling)
"""
from . import _ctypes_defs as defs  # noqa: F401

__all__ = ["defs"]
