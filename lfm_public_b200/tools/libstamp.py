"""Identity of the DEVICE code of a built library: SHA-256 of the `.nv_fatbin` section of the ELF file.

The whole-file hash of liblfmgpu.so changes from build to build (nvcc leaves the name of a temporary file in the symbol
string table) although the cubins are byte-identical; the fat binary section holds exactly what runs on the GPU, so a
profiler capture stamped with this value (scripts/ncu_traffic.py -> profiles/ncu_traffic.json) stays valid for a rebuild of
the same sources and is dropped by bench.py as soon as a kernel changes."""
from __future__ import annotations

import hashlib
import struct


def device_code_sha256(path: str) -> str:
    raw = open(path, "rb").read()
    if raw[:4] != b"\x7fELF" or raw[4] != 2 or raw[5] != 1:
        raise ValueError(f"{path}: not a little-endian ELF64 file")
    e_shoff, = struct.unpack_from("<Q", raw, 0x28)
    e_shentsize, e_shnum, e_shstrndx = struct.unpack_from("<HHH", raw, 0x3A)

    def section(i):
        name, _type, _flags, _addr, off, size = struct.unpack_from("<IIQQQQ", raw, e_shoff + i * e_shentsize)
        return name, off, size

    _, str_off, str_size = section(e_shstrndx)
    names = raw[str_off:str_off + str_size]
    h = hashlib.sha256()
    found = False
    for i in range(e_shnum):
        name, off, size = section(i)
        if names[name:names.index(b"\0", name)] == b".nv_fatbin":
            h.update(raw[off:off + size])
            found = True
    if not found:
        raise ValueError(f"{path}: no .nv_fatbin section (not a CUDA library)")
    return h.hexdigest()


if __name__ == "__main__":
    import sys
    for p in sys.argv[1:]:
        print(device_code_sha256(p), p)
