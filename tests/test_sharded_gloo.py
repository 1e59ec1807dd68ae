"""N > 1 host logic on CPU: world_size-2 and -3 gloo groups run gpview_b200.sharded (slab plan + gather to rank 0); each rank's
"slab result" is cut out of the oracle's whole-grid result (the definition of a correct slab), the gathered
streams on rank 0 must equal the whole."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import mesh_path


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, path, l1, l2, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpview_b200 import sharded
        from oracle import oraclebind as O
        r = O.OracleMesh(path).voxelize(l1, l2, O.FILL_CERTIFIED | O.NO_NORMALS, 1)
        nx, ny, nz = [int(x) for x in r.num_div]
        plane = nx * ny
        off = np.concatenate([[0], np.cumsum(r.cell_count[r.boundary_index])])
        cuts = sharded.plan_slabs(sharded.layer_cost(r.boundary_index, off, plane, nz), world)
        z0, z1 = cuts[rank], cuts[rank + 1]
        c0, c1 = z0 * plane, z1 * plane
        b0, b1 = int(r.prefix[c0]), int(r.prefix[c1]) if c1 < r.cells else r.nb
        l1s = torch.from_numpy((r.l1_state[c0:c1] * 127).astype(np.uint8))
        pre = torch.from_numpy((r.prefix[c0:c1] - b0).astype(np.int32)).view(torch.uint8)     # slab-local, like the C ABI
        l2s = torch.from_numpy((r.l2_state[b0 * r.n23:b1 * r.n23] * 127).astype(np.uint8))
        pieces = {"l1": (l1s, 1, 0), "prefix": (pre, 4, 0), "l2": (l2s, r.n23, 1)}
        out, sizes = sharded.gather_to_rank0(dist, torch, rank, world, pieces, c1 - c0, b1 - b0)
        if rank == 0:
            ok = (np.array_equal(out["l1"].numpy(), r.l1_state * 127) and np.array_equal(out["prefix"].view(torch.int32).numpy(), r.prefix)
                  and np.array_equal(out["l2"].numpy(), r.l2_state * 127) and int(sizes[:, 1].sum()) == r.nb and cuts[0] == 0 and cuts[-1] == nz)
            q.put(("ok" if ok else "mismatch", cuts))
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("error: %r" % (e,), None))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_reassembles_the_whole_grid(tmp_path_factory, world):
    path = mesh_path("cessna", tmp_path_factory.getbasetemp())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, 32, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, cuts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert status == "ok", status
    assert len(cuts) == world + 1 and all(b > a for a, b in zip(cuts, cuts[1:]))


def test_plan_slabs_balances_cost():
    from gpview_b200 import sharded
    rng = np.random.default_rng(0)
    cost = rng.uniform(0, 1, 256) ** 4
    for world in (1, 2, 4, 8):
        cuts = sharded.plan_slabs(cost, world)
        assert cuts[0] == 0 and cuts[-1] == 256 and len(cuts) == world + 1
        per = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
        assert max(per) <= cost.sum() / world + cost.max() + 1e-9
    assert sharded.plan_slabs(np.zeros(8), 8) == list(range(9))
    with pytest.raises(ValueError):
        sharded.plan_slabs(np.ones(4), 5)


def test_rebalance_converges_on_a_skewed_cost():
    """sharded.rebalance: ranks whose true cost per layer differs from the a-priori model (here: a quadratic profile the model knows
    nothing about) converge to slabs of equal measured time in a few steps; cuts stay strictly increasing and cover [0, nz]."""
    import numpy as np
    from gpview_b200 import sharded
    nz, world = 256, 8
    true = 1.0 + 40.0 * np.exp(-((np.arange(nz) - 128.0) / 18.0) ** 2) + 0.02 * np.arange(nz)   # what the GPU would take per layer
    layer = np.ones(nz)                                                                           # what the model believes
    cuts = sharded.plan_slabs(layer, world)
    for _ in range(6):
        t = np.array([true[cuts[r]:cuts[r + 1]].sum() for r in range(world)])
        if t.max() < 1.05 * t.mean():
            break
        layer, cuts = sharded.rebalance(layer, cuts, t)
        assert cuts[0] == 0 and cuts[-1] == nz and all(b > a for a, b in zip(cuts, cuts[1:]))
    t = np.array([true[cuts[r]:cuts[r + 1]].sum() for r in range(world)])
    assert t.max() < 1.25 * t.mean(), (cuts, t)


def test_column_ownership_balances_level2_and_matches_the_measured_shares():
    """GPV_GATHER shares the Level-2 refinement out by Level-1 column (struct Own in gpv_kernels.cuh; sharded.column_owner is its
    host-side statement).  On the headline model (fixture boundary cells of cessna 256/16): every column has exactly one owner, the
    ranks' shares of boundary cells and of (cell, triangle) pairs -- the two things Level 2 costs -- stay within a few percent of equal
    at 2 / 4 / 8 ranks, and the cell counts are the ones the ranks of the committed 4- and 8-GPU runs reported."""
    import json
    import os
    import numpy as np
    from gpview_b200 import sharded
    from util import ROOT, golden
    info, z = golden("cessna_256_16")
    nx, ny, nz = info["num_div"]
    col = z["boundary_index"].astype(np.int64) % (nx * ny)
    tris = z["tri_count_boundary"].astype(np.int64)
    allcols = np.arange(nx * ny)
    for world in (2, 4, 8):
        own_all = sharded.column_owner(allcols, world, 16, nx)
        assert own_all.min() == 0 and own_all.max() == world - 1
        assert np.bincount(own_all, minlength=world).max() - np.bincount(own_all, minlength=world).min() <= ny      # columns: equal up to row ends
        own = sharded.column_owner(col, world, 16, nx)
        cells = np.bincount(own, minlength=world)
        pairs = np.bincount(own, weights=tris, minlength=world)
        assert cells.sum() == info["l1_boundary"]
        assert cells.max() <= 1.03 * cells.mean() and pairs.max() <= 1.06 * pairs.mean(), (world, cells, pairs)
        line = os.path.join(ROOT, "profiles", "r02_scale_n%d.json" % world)
        if os.path.exists(line):   # the ranks' own reports of that run (bench.py: phase_ms_per_rank[r].l2_cells)
            got = [r["l2_cells"] for r in json.loads(open(line).read().strip().splitlines()[-1])["phase_ms_per_rank"]]
            assert got == [int(c) for c in cells], (world, got, cells)
    # groups of 256 / n2^2 columns for smaller n2; one rank owns everything at world 1
    assert sharded.column_owner(allcols, 1, 16, nx).max() == 0
    g4 = sharded.column_owner(allcols, 4, 4, nx)
    assert all(len(set(g4[k:k + 16])) == 1 for k in range(0, 16 * 10, 16))
