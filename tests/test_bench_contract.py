"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for the reference arm (CPU) and the CUDA arm."""
import json
import os
import subprocess
import sys

import pytest

import common

BENCH = os.path.join(common.ROOT, "bench.py")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args, timeout=900):
    out = subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, cwd=common.ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.skipif(not common.have_ref(), reason="reference binary not built")
def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-n", "16", "--ref-arm-n", "24"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d) and d["value"] > 0
    # the line names the mesh the CPU arm really advanced (24^3 here), not the GPU arm's
    assert "24^3" in d["config"]["workload"] and d["config"]["total_cells"] == 24 ** 3 and "24^3" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["smaller_sample"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["config"]["workload"]


@pytest.mark.gpu
def test_cuda_arm_line():
    d = _run(["--n", "32", "--steps", "2", "--warmup", "3", "--ref-n", "16", "--ref-steps", "2", "--extras-n", "32"])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["value"] > 0 and d["finite"] and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["kernel"] == "tile_stage" and 0 < r["frac"] < 1.5 and r["unit"] == "GB/s"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 5 * 32 ** 3 * 8 and e["finite"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    x = d["extra"]
    assert "error" not in x, x
    assert x["f32"]["value"] > 0 and all(x["numbering"][k]["value"] > 0 for k in ("morton", "bricks", "lex"))


def test_device_code_stamp(tmp_path):
    """roofline.traffic is only reported for the kernels the committed ncu capture was taken from: the stamp is the SHA-256 of
    the library's .nv_fatbin section (tools/libstamp.py) -- unchanged by what differs between two builds of the same sources
    (nvcc leaves a temporary file name in the symbol strings), changed by any change of the device code."""
    from lfm_public_b200.tools.libstamp import device_code_sha256
    lib = os.path.join(common.ROOT, "lfm_public_b200", "liblfmgpu.so")
    if not os.path.exists(lib):
        pytest.skip("liblfmgpu.so not built")
    stamp = device_code_sha256(lib)
    raw = bytearray(open(lib, "rb").read())
    # a byte of the symbol string table (behind every loaded section): same stamp
    a = bytearray(raw)
    i = a.rindex(b"fatbinData")
    a[i - 8] ^= 1
    pa = tmp_path / "a.so"
    pa.write_bytes(bytes(a))
    assert device_code_sha256(str(pa)) == stamp
    # a byte in the middle of the fat binary: another stamp
    import struct
    e_shoff, = struct.unpack_from("<Q", raw, 0x28)
    e_shentsize, e_shnum, _ = struct.unpack_from("<HHH", raw, 0x3A)
    sizes = [struct.unpack_from("<IIQQQQ", raw, e_shoff + k * e_shentsize)[4:6] for k in range(e_shnum)]
    off, size = max(sizes, key=lambda t: t[1])   # the fat binary is by far the largest section
    b = bytearray(raw)
    b[off + size // 2] ^= 1
    pb = tmp_path / "b.so"
    pb.write_bytes(bytes(b))
    assert device_code_sha256(str(pb)) != stamp
    tr = json.load(open(os.path.join(common.ROOT, "profiles", "ncu_traffic.json")))
    assert all("fatbin_sha256" in v["f64"] for k, v in tr.items() if k.startswith("tile_"))
