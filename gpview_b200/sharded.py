"""z-slab sharding of one voxelization over the GPUs of a box -- torch.distributed plumbing only.

(The recommended multi-GPU path is GPV_GATHER in the C ABI -- DESIGN.md 7: no collective, Level-2 shared out by Level-1 column,
`column_owner` below is its ownership rule.  What follows is the NCCL send/recv gather of z-slabs it is measured against, and the
slab plan the end-to-end leg of bench.py uses.)

The linear cell index is z-major (idx = r*ny*nx + q*nx + p, cuda/CUDAClassifyTessellation.cu:393), so a slab [z0,z1) owns a
contiguous byte range of Level1InOut / Level1BoundaryPrefixSum, and -- boundary cells being numbered in ascending index
order -- a contiguous range of the Level-2 blocks.  Sharding is therefore: pick cuts, run the C ABI on each slab
(gpv_params.z0/z1), concatenate.  The only exchange step is the gather of the slab pieces to rank 0 (sizes by all_gather,
payload by batched point-to-point sends over NCCL/NVLink; with gloo on CPU tensors in the tests) plus adding the boundary
counts of the lower slabs to each slab-local prefix sum.

The Level-1 passes are cheap and replicated on every rank (every rank needs the full column lists for the Level-2 parity
rays anyway: the candidate set of the reference is the whole column, cu:461-463); Level-2, ~80 % of the time, is what
is sharded, so cuts are balanced by the Level-2 cost per z-layer, not by layer count.
"""
import numpy as np


def column_owner(col, world, n2, nx):
    """Which rank of a GPV_GATHER call refines Level-1 column `col` (numpy arrays welcome): the host-side statement of `struct Own`
    in gpv_kernels.cuh.  Columns are dealt out in groups of max(1, 256 // n2**2) consecutive columns, group k of grid row j to rank
    (k + j) % world -- neighbouring columns cost about the same, so this balances the Level-2 work without a cost model, and the skew
    by the row keeps walls along x or y from landing on one rank."""
    group = max(1, 256 // (n2 * n2))
    return (col // group + col // nx) % world


def plan_slabs(cost_per_layer, world):
    """Cut nz layers into `world` contiguous non-empty slabs of ~equal cost.  Deterministic: every rank computes the same
    cuts from the same replicated Level-1 pre-pass, no communication needed.  Returns world+1 cut positions."""
    cost = np.asarray(cost_per_layer, np.float64)
    nz = len(cost)
    if world > nz:
        raise ValueError("more ranks (%d) than z-layers (%d)" % (world, nz))
    c = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        z = int(np.searchsorted(c, c[-1] * r / world))
        z = min(max(z, cuts[-1] + 1), nz - (world - r))
        cuts.append(z)
    cuts.append(nz)
    return cuts


def layer_cost(boundary_index, cell_off, plane, nz, per_cell_overhead=8.0):
    """Level-2 cost model per z-layer: (triangles in the cell list + a constant for the parity rays) per boundary cell."""
    bi = np.asarray(boundary_index, np.int64)
    off = np.asarray(cell_off, np.int64)
    return np.bincount(bi // plane, weights=(off[1:] - off[:-1]) + per_cell_overhead, minlength=nz)


def rebalance(layer_cost_now, cuts, seconds_per_rank):
    """One step of measured-time load balancing: rescale every slab's layer costs so that the slab's cost equals the time its
    rank measured, then re-cut.  `layer_cost_now` (nz floats, > 0) is updated in place and returned with the new cuts.  Every rank
    calls this with the same all-gathered times, so every rank gets the same cuts (no data moves)."""
    layer = layer_cost_now
    world = len(cuts) - 1
    for r in range(world):
        sl = slice(cuts[r], cuts[r + 1])
        layer[sl] *= max(float(seconds_per_rank[r]), 1e-9) / max(float(layer[sl].sum()), 1e-30)
    return layer, plan_slabs(layer, world)


def gather_to_rank0(dist, torch, rank, world, pieces, cells, n_boundary, out_cache=None):
    """Concatenate the slab pieces on rank 0.

    pieces: dict name -> (uint8 tensor holding this rank's bytes, bytes per unit, unit kind) with unit kind 0 = per cell,
            1 = per boundary cell (e.g. {"l1": (t, 1, 0), "prefix": (t, 4, 0), "l2": (t, n2**3, 1)}).
    Returns (dict name -> uint8 tensor with the whole stream, counts[world, 2]) on rank 0, ({}, counts) elsewhere.
    The "prefix" stream (int32) is made global by adding the boundary counts of the lower slabs."""
    device = next(iter(pieces.values()))[0].device
    sizes = torch.tensor([cells, n_boundary], device=device, dtype=torch.int64)
    allsz = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(allsz, sizes)
    allsz = torch.stack(allsz).cpu().numpy()
    out = out_cache if out_cache is not None else {}
    ops = []
    for name, (mine, unit, which) in pieces.items():
        if rank == 0:
            tot = int(allsz[:, which].sum()) * unit
            if name not in out or out[name].numel() != tot:
                out[name] = torch.empty(tot, dtype=torch.uint8, device=device)
            o = 0
            for r in range(world):
                nby = int(allsz[r][which]) * unit
                if r == 0:
                    out[name][o:o + nby].copy_(mine[:nby], non_blocking=True)
                elif nby:
                    ops.append(dist.P2POp(dist.irecv, out[name][o:o + nby], r))
                o += nby
        else:
            nby = int(allsz[rank][which]) * unit
            if nby:
                ops.append(dist.P2POp(dist.isend, mine[:nby], 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank == 0 and "prefix" in out:
        pre = out["prefix"].view(torch.int32)
        o = base = 0
        for r in range(world):
            ncell = int(allsz[r][0])
            if base:
                pre[o:o + ncell] += base
            o += ncell
            base += int(allsz[r][1])
    return (out if rank == 0 else {}), allsz


def wrap_device_bytes(torch, ptr, nbytes):
    """A uint8 torch tensor aliasing `nbytes` of device memory at `ptr` (views of the ctx-owned result buffers)."""
    if not nbytes:
        return torch.empty(0, dtype=torch.uint8, device="cuda")

    class _P:
        pass
    p = _P()
    p.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(p, device="cuda")
