"""Consumer side of the file contract: the ObjN* sets a voxelization run writes (gpv_save / Object::SaveVoxelization), as a
torch Dataset for 3-D CNN training -- the use the reference states for its voxelizer (README.md:22-27 of idealab-isu/GPView).

    ds = VoxelFolder("out", resolution="dense")        # every ObjNVoxelConfig.txt under out/
    x, meta = ds[0]                                    # uint8 tensor [1, D, H, W] (z, y, x), values 0 / 1 / 2

`resolution="level1"` gives the Level-1 grid; "dense" the effective-resolution grid (Level-1 cells expanded, boundary cells
replaced by their Level-2 blocks: gpv_expand_dense).  States are returned as 0 outside / 1 inside / 2 boundary (the reference's
in-memory encoding, App. A.8), `occupancy=True` folds them to 0 / 1.  Reading is host-only (no GPU needed)."""
import os
import re

import numpy as np

from . import binding as B


def list_object_ids(directory):
    """Object ids of the COMPLETE sets in `directory`, ascending: an ObjNVoxelConfig.txt that parses and every stream it implies
    present with the size it implies (gpv_check_voxels) -- a truncated set of an interrupted run is not listed."""
    ids = []
    for name in os.listdir(directory):
        m = re.fullmatch(r"Obj(-?\d+)VoxelConfig\.txt", name)
        if m and B.check_voxels(directory, int(m.group(1))):
            ids.append(int(m.group(1)))
    return sorted(ids)


def load_grid(directory, obj_id, resolution="dense", occupancy=False):
    """One model as a numpy array [D, H, W] (z, y, x) plus its VoxelConfig fields."""
    v = B.load_voxels(directory, obj_id)
    nx, ny, nz = v["num_div"]
    if resolution == "level1" or v["level2_inout"] is None:
        grid = v["level1_inout"].reshape(nz, ny, nx)
    elif resolution == "dense":
        grid = B.expand_dense(v["level1_inout"], v["prefix_sum"], v["level2_inout"], v["num_div"], v["num_div2"][0])
    else:
        raise ValueError("resolution is 'level1' or 'dense'")
    grid = grid // 127                      # file bytes 0 / 127 / 254 -> states 0 / 1 / 2
    if occupancy:
        grid = (grid > 0).astype(np.uint8)
    meta = {k: v[k] for k in ("name", "num_div", "num_div2", "grid_size", "counts")}
    return grid, meta


class VoxelFolder:
    """torch.utils.data.Dataset over the voxel sets of a directory (torch is imported on first use)."""

    def __init__(self, directory, resolution="dense", occupancy=False, ids=None):
        self.directory, self.resolution, self.occupancy = directory, resolution, occupancy
        self.ids = list(ids) if ids is not None else list_object_ids(directory)

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        import torch
        grid, meta = load_grid(self.directory, self.ids[i], self.resolution, self.occupancy)
        return torch.from_numpy(np.ascontiguousarray(grid)).unsqueeze(0), meta
