#!/usr/bin/env python3
"""A few small voxelizations through every kernel family, for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_gpu.py        (racecheck / initcheck likewise)
Plain, normals + canonical lists, host call with the 2-bit Level-2 transfer, collision structures, a one-rank gathering call."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpview_b200 as gpv  # noqa: E402
from gpview_b200 import binding as B  # noqa: E402


def main():
    ctx = gpv.Context(0)
    for name, l1, l2 in (("torus.off", 32, 4), ("sphere.obj", 16, 16), ("block.off", 24, 2), ("cad.obj", 20, 3), ("sphere.obj", 16, 8)):
        mesh = gpv.load_mesh(os.path.join(ROOT, "tests", "golden", "meshes", name))
        plain = ctx.voxelize(mesh, gpv.Params(l1, l2, 0))
        full = ctx.voxelize(mesh, gpv.Params(l1, l2, gpv.GPV_NORMALS | gpv.GPV_KEEP_LISTS | gpv.GPV_COLLISION))
        assert plain.counts == full.counts
        ctx.collision_boxes()
        if all(int(n) & (int(n) - 1) == 0 for n in full.num_div):
            ctx.build_hierarchy()
        nb, n23, cells = full.nb, full.n23, full.cells
        bufs = [np.zeros(cells, np.uint8), np.zeros(cells, np.int32), np.zeros(nb, np.int32), np.zeros(nb * n23, np.uint8), np.zeros(cells * 3, np.uint8), np.zeros(nb * n23 * 3, np.uint8)]
        hs = B.CHostStreams(*[b.ctypes.data for b in bufs], bufs[3].nbytes, nb)
        r = ctx.voxelize_host(mesh, gpv.Params(l1, l2, gpv.GPV_NORMALS | gpv.GPV_PACKED_L2), hs)
        assert r.counts == plain.counts and np.array_equal(bufs[3], full.level2_inout())
        ctx.gather_create(cells, nb * n23, gpv.GPV_NORMALS)
        ctx.gather_attach_local(ctx, 0, 1)
        g = ctx.voxelize(mesh, gpv.Params(l1, l2, gpv.GPV_GATHER | gpv.GPV_NORMALS))
        assert g.counts == plain.counts
        ctx.gather_detach()
        print(name, l1, l2, "ok", plain.counts)
    ctx.close()


if __name__ == "__main__":
    main()
