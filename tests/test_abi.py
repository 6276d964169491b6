"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import common

ROOT = common.ROOT


def _declared(header):
    s = open(os.path.join(ROOT, "include", header)).read()
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return sorted(set(re.findall(r"\b(lfm(?:gpu|host)_[a-z0-9_]+)\s*\(", s)))


def test_lfmhost_exports_every_declared_symbol():
    L = C.CDLL(os.path.join(ROOT, "lfm_public_b200", "liblfmhost.so"))
    for name in _declared("lfmhost.h"):
        assert hasattr(L, name), name


def test_lfmgpu_exports_every_declared_symbol():
    path = os.path.join(ROOT, "lfm_public_b200", "liblfmgpu.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", ROOT, "gpu"])
    L = C.CDLL(path)          # links libcudart only; NCCL is dlopen'ed on first use
    names = _declared("lfmgpu.h")
    assert len(names) >= 35
    for name in names:
        assert hasattr(L, name), name
    from lfm_public_b200 import gpu_api
    assert sorted(gpu_api.SYMBOLS) == names


def test_gpu_path_fails_loudly_without_device():
    from lfm_public_b200 import gpu_api
    if gpu_api.device_count() > 0:
        return
    import pytest
    case = None
    with pytest.raises(Exception):
        gpu_api.GpuSolver(case)
