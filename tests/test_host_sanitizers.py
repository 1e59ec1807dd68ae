"""The host half of the product under AddressSanitizer + UndefinedBehaviorSanitizer (tests/cpu_probe/sanitize_driver.cpp): the
mesh readers (strict and tolerant, 1 and 3 loader threads) and the voxel-file reader over valid, truncated, mutated and random
files.  A reader may refuse a file; it must never touch memory it does not own, leak, or run into undefined behaviour."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from util import ROOT, mesh_path


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("san") / "sanitize_driver")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-ffp-contract=off", "-pthread", "-o", out,
           os.path.join(ROOT, "tests", "cpu_probe", "sanitize_driver.cpp"), os.path.join(ROOT, "gpview_b200", "csrc", "gpv_host.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr or "ubsan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("this toolchain has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    return out


def _run(driver, *args):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([driver, *args], capture_output=True, text=True, env=env, timeout=600)
    text = r.stdout + r.stderr
    assert r.returncode == 0 and "Sanitizer" not in text and "runtime error" not in text, text[-3000:]
    ok, refused = [int(x) for x in r.stdout.split()[1::2]]
    return ok, refused


def test_mesh_readers_under_sanitizers(driver, tmp_path_factory, tmp_path):
    from test_host_cpu import OBJ_QUIRKS
    rng = np.random.default_rng(11)
    base = tmp_path_factory.getbasetemp()
    for n in ("cessna", "torus", "block"):
        p = mesh_path(n, base)
        shutil.copy(p, tmp_path / os.path.basename(p))
    for k, v in OBJ_QUIRKS.items():
        (tmp_path / ("q_%s.obj" % k)).write_bytes(v.encode())
    off = open(mesh_path("torus", base), "rb").read()
    obj = open(mesh_path("sphere", base), "rb").read()[:40000]
    alpha = b"0123456789 \t\n\r.-+eEvf/#xOF"
    for i in range(40):
        (tmp_path / ("trunc%d.off" % i)).write_bytes(off[:int(rng.integers(0, len(off)))])
        (tmp_path / ("trunc%d.obj" % i)).write_bytes(obj[:int(rng.integers(0, len(obj)))])
        raw = bytes(rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8))
        (tmp_path / ("rnd%d.obj" % i)).write_bytes(raw)
        (tmp_path / ("rnd%d.off" % i)).write_bytes(b"OFF\n" + raw)
    for i in range(120):
        soup = bytes(alpha[j] for j in rng.integers(0, len(alpha), int(rng.integers(1, 600))))
        (tmp_path / ("soup%d.obj" % i)).write_bytes(soup)
        (tmp_path / ("soup%d.off" % i)).write_bytes(b"OFF " + soup)
        for src, ext in ((off[:30000], "off"), (obj[:20000], "obj")):
            a = bytearray(src)
            for _ in range(int(rng.integers(1, 8))):
                a[int(rng.integers(0, len(a)))] = alpha[int(rng.integers(0, len(alpha)))]
            (tmp_path / ("mut%d.%s" % (i, ext))).write_bytes(bytes(a))
    ok, refused = _run(driver, "meshes", str(tmp_path))
    assert ok > 100 and refused > 500, (ok, refused)


def test_voxel_file_reader_under_sanitizers(driver, oracle, tmp_path_factory, tmp_path):
    r = oracle.OracleMesh(mesh_path("torus", tmp_path_factory.getbasetemp())).voxelize(16, 2, oracle.FILL_CERTIFIED, 4)
    (tmp_path / "m0").mkdir()
    r.save(5, str(tmp_path / "m0"))
    cfg = (tmp_path / "m0" / "Obj5VoxelConfig.txt").read_bytes()
    rng = np.random.default_rng(3)
    alpha = b"0123456789 \t\n.-+e"
    n = 120
    for i in range(1, n):
        d = tmp_path / ("m%d" % i)
        shutil.copytree(tmp_path / "m0", d)
        a = bytearray(cfg)
        if i % 4 == 0:
            for _ in range(int(rng.integers(1, 5))):
                a[int(rng.integers(0, len(a)))] = alpha[int(rng.integers(0, len(alpha)))]
        elif i % 4 == 1:
            a = a[:int(rng.integers(0, len(a)))]
        elif i % 4 == 2:
            k = int(rng.integers(0, len(a)))
            a[k:k] = bytes(alpha[j] for j in rng.integers(0, len(alpha), int(rng.integers(1, 30))))
        else:
            f = ["Obj5Level1InOut.raw", "Obj5Level2InOut.raw", "Obj5Level1BoundaryPrefixSum.raw"][i % 3]
            b = (d / f).read_bytes()
            (d / f).write_bytes(b[:int(rng.integers(0, len(b)))])
        (d / "Obj5VoxelConfig.txt").write_bytes(bytes(a))
    ok, refused = _run(driver, "voxels", str(tmp_path), str(n))
    assert ok >= 1 and refused > 50, (ok, refused)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_expand_pool_under_sanitizers(tmp_path, sanitizer):
    """The host-thread pool of the 2-bit Level-2 transfer (gpv_expand.cpp) driven like gpv_voxelize_host drives it -- begin, chunks
    trickling in, end -- 2 x 40 calls from two client threads, every output byte checked: no data race (TSan), no stray write
    (ASan), same bytes as the definition (tests/cpu_probe/expand_probe.cpp)."""
    out = str(tmp_path / "expand_probe")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-fsanitize=" + sanitizer, "-fno-omit-frame-pointer", "-pthread", "-o", out,
           os.path.join(ROOT, "tests", "cpu_probe", "expand_probe.cpp"), os.path.join(ROOT, "gpview_b200", "csrc", "gpv_expand.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("tsan" in r.stderr or "asan" in r.stderr or "sanitize" in r.stderr):
        pytest.skip("this toolchain has no sanitizer runtime")
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, GPV_HOST_THREADS="5", ASAN_OPTIONS="detect_leaks=0", TSAN_OPTIONS="halt_on_error=1")   # (the pool lives as long as the process: not a leak)
    r = subprocess.run([out, "40"], capture_output=True, text=True, env=env, timeout=900)
    text = r.stdout + r.stderr
    assert r.returncode == 0 and "ok" in r.stdout and "Sanitizer" not in text and "runtime error" not in text, text[-3000:]
