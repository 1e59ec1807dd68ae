"""The C-ABI library loads without a GPU and exports exactly the symbols include/gpview_b200.h declares (no compute calls)."""
import os
import re
import subprocess

from util import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "gpview_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)))


def test_header_and_binding_agree(product):
    from gpview_b200 import binding
    declared = header_functions()
    assert sorted(binding.NATIVE_SYMBOLS + binding.COMPAT_SYMBOLS) == declared


def test_library_exports_every_declared_symbol(product):
    out = subprocess.check_output(["nm", "-D", "--defined-only", product.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing


def test_compat_symbols_are_unmangled_c(product):
    out = subprocess.check_output(["nm", "-D", "--defined-only", product.LIB_PATH], text=True)
    for name in ("CUDAClassifyTessellation", "CUDAClassifyTessellationLevel2", "CUDAClassifyInOutLevel2", "THRUSTDeviceFindMax"):
        assert re.search(r" T %s$" % name, out, flags=re.M), name


def test_no_cpu_fallback(product):
    """Without a device the native tier refuses to create a context instead of computing on the host."""
    L = product.lib()
    if L.gpv_device_count() > 0:
        return
    try:
        product.Context(0)
    except product.GpvError as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("gpv_create succeeded without a GPU")


def test_only_sm100a_code_in_the_library(product):
    out = subprocess.run(["cuobjdump", "-lelf", product.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under gpview_b200/ or include/ may import, link or execute it."""
    bad = []
    for base in ("gpview_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    for m in re.finditer(r"^.*(?:import|include|from|dlopen|CDLL|-l|-L).*\boracle\b.*$", txt, flags=re.M):
                        line = m.group(0).strip()
                        if line.startswith(("//", "#", "*", '"""')) or "oracle/gpv_oracle.c gpvo_axis_table" in line:
                            continue
                        bad.append((f, line))
    assert not bad, bad


def test_flags_and_struct_sizes_match_the_header(product):
    """The ctypes mirror (gpview_b200/binding.py) must carry the header's flag values and struct sizes: compile a probe against
    include/gpview_b200.h with the host compiler and compare (no GPU needed)."""
    import ctypes as C
    import tempfile
    from gpview_b200 import binding as B
    src = r"""
#include "gpview_b200.h"
#include <stdio.h>
int main(void) {
    printf("GPV_NORMALS %d\nGPV_NO_LEVEL2 %d\nGPV_KEEP_LISTS %d\nGPV_PROFILE %d\nGPV_GATHER %d\n", GPV_NORMALS, GPV_NO_LEVEL2, GPV_KEEP_LISTS, GPV_PROFILE, GPV_GATHER);
    printf("GPV_SAVE_COMPUTED_ONLY %d\nGPV_BATCH_TOLERANT_LOAD %d\nGPV_LOAD_TOLERANT %u\n", GPV_SAVE_COMPUTED_ONLY, GPV_BATCH_TOLERANT_LOAD, GPV_LOAD_TOLERANT);
    printf("gpv_mesh %zu\ngpv_grid %zu\ngpv_params %zu\ngpv_result %zu\ngpv_host_streams %zu\ngpv_gather_desc %zu\ngpv_batch_stats %zu\ngpv_voxel_file %zu\n",
           sizeof(gpv_mesh), sizeof(gpv_grid), sizeof(gpv_params), sizeof(gpv_result), sizeof(gpv_host_streams), sizeof(gpv_gather_desc), sizeof(gpv_batch_stats), sizeof(gpv_voxel_file));
    printf("GPV_PHASE_COUNT %d\n", (int)GPV_PHASE_COUNT);
    return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = dict(line.split() for line in subprocess.check_output([exe], text=True).splitlines())
    for name in ("GPV_NORMALS", "GPV_NO_LEVEL2", "GPV_KEEP_LISTS", "GPV_PROFILE", "GPV_GATHER", "GPV_SAVE_COMPUTED_ONLY", "GPV_BATCH_TOLERANT_LOAD"):
        assert int(got[name]) == getattr(B, name), name
    assert int(got["GPV_LOAD_TOLERANT"]) == 1  # binding.load_mesh(tolerant=True) passes 1
    flags = [int(got[n]) for n in ("GPV_NORMALS", "GPV_NO_LEVEL2", "GPV_KEEP_LISTS", "GPV_PROFILE", "GPV_GATHER", "GPV_SAVE_COMPUTED_ONLY", "GPV_BATCH_TOLERANT_LOAD")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags), "gpv_params.flags bits must be distinct powers of two"
    mirrors = {"gpv_mesh": B.CMesh, "gpv_grid": B.CGrid, "gpv_params": B.CParams, "gpv_result": B.CResult, "gpv_host_streams": B.CHostStreams,
               "gpv_gather_desc": B.CGatherDesc, "gpv_batch_stats": B.CBatchStats, "gpv_voxel_file": B.CVoxelFile}
    for name, cls in mirrors.items():
        assert int(got[name]) == C.sizeof(cls), (name, got[name], C.sizeof(cls))
    assert int(got["GPV_PHASE_COUNT"]) == len(B.PHASES)


def test_plain_c_example_builds_and_fails_loudly_without_a_gpu(product):
    """tools/example_native.c -- the native tier used from C99 exactly as INTEGRATION.md shows -- compiles with -Wall -Wextra against
    the header, links against the library, reads a mesh with the reference's loader semantics, and (here, without a device)
    stops at gpv_create with the no-CPU-fallback message instead of computing anything on the host."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools"), "example_native"])
    r = subprocess.run([os.path.join(ROOT, "tools", "example_native"), os.path.join(ROOT, "tests", "golden", "meshes", "torus.off"), "32", "4", "/tmp"],
                       capture_output=True, text=True)
    assert "576 triangles, Level-1 grid 32 x 32 x 12" in r.stdout
    if product.lib().gpv_device_count() == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
