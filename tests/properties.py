"""Size-independent properties of a voxelization: what must hold for ANY mesh and resolution, so that the
full-size BASELINE.json configurations -- where the CPU oracle takes minutes -- are still checked on the GPU box.
Test infrastructure only.  Streams are in the file encoding (0 outside / 127 inside / 254 boundary)."""
import numpy as np


def _histogram(a, chunk=1 << 26):
    h = np.zeros(256, np.int64)
    for s in range(0, a.size, chunk):
        h += np.bincount(a[s:s + chunk], minlength=256)
    return h


def check_stream_structure(l1, prefix, bidx, l2, counts, n23, chunk=1 << 26):
    """Internal consistency of the streams of one result (src/Object.cpp:3258-3284, 3353-3378):
    states are 0/127/254; prefix = exclusive running count of boundary cells; boundary_index = their ascending linear indices;
    the four counts are the histograms of the two state streams; Level-2 holds n2^3 voxels per boundary cell.
    Works through the grid in chunks: the 620 M-cell configuration must not need tens of GB of temporaries."""
    l1, prefix, bidx = np.asarray(l1), np.asarray(prefix), np.asarray(bidx)
    assert l1.dtype == np.uint8 and prefix.dtype == np.int32 and bidx.dtype == np.int32
    hist1 = _histogram(l1, chunk)
    assert hist1[0] + hist1[127] + hist1[254] == l1.size, "Level-1 states outside {0,127,254}"
    nb = int(hist1[254])
    assert [int(hist1[127]), nb] == list(counts[:2]), ("Level-1 counts", counts[:2], int(hist1[127]), nb)
    assert prefix.size == l1.size and bidx.size == nb, ("stream sizes", prefix.size, l1.size, bidx.size, nb)
    base = 0
    for s in range(0, l1.size, chunk):
        flag = l1[s:s + chunk] == 254
        run = np.cumsum(flag, dtype=np.int64)
        assert np.array_equal(prefix[s:s + chunk], base + run - flag), "prefix is not the exclusive scan of the boundary flags (cells %d..)" % s
        where = np.flatnonzero(flag) + s
        assert np.array_equal(bidx[base:base + where.size], where), "boundary_index is not the ascending list of boundary cells (cells %d..)" % s
        base += int(run[-1]) if run.size else 0
    assert base == nb
    if l2 is not None:
        l2 = np.asarray(l2)
        assert l2.dtype == np.uint8 and l2.size == nb * n23, ("Level-2 size", l2.size, nb, n23)
        hist2 = _histogram(l2, chunk)
        assert hist2[0] + hist2[127] + hist2[254] == l2.size, "Level-2 states outside {0,127,254}"
        assert [int(hist2[127]), int(hist2[254])] == list(counts[2:4]), ("Level-2 counts", counts[2:4], int(hist2[127]), int(hist2[254]))
    return nb


def mesh_volume(tris, chunk=1 << 20):
    """|signed volume| of a closed triangle mesh (float64, divergence theorem; chunked: 10 M triangles stay below 1 GB)."""
    t = np.asarray(tris).reshape(-1, 3, 3)
    total = 0.0
    for a in range(0, len(t), chunk):
        c = t[a:a + chunk].astype(np.float64)
        total += float(np.einsum("ij,ij->i", c[:, 0], np.cross(c[:, 1], c[:, 2])).sum())
    return abs(total) / 6.0


def rotated(tris, angles=(0.37, 1.13, 2.41), chunk=1 << 20):
    """The mesh turned into generic position (fixed rotation about x, y, z; float32 vertices).  The synthetic bodies are
    surfaces of revolution with meridians in the x = 0, y = 0 and x = +-y planes and symmetric bounding boxes, so whole
    lines of ray origins run exactly along mesh edges -- where the reference's inclusive-edge Moller-Trumbore counts a
    crossing twice (App. A.5).  That is reference behaviour (and bit-exact here), but it breaks volume arguments."""
    a, b, c = angles
    rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    rt = (rx @ ry @ rz).T
    v = np.asarray(tris, np.float32).reshape(-1, 3)
    out = np.empty_like(v)
    for s in range(0, len(v), chunk):
        out[s:s + chunk] = (v[s:s + chunk].astype(np.float64) @ rt).astype(np.float32)
    return out.reshape(-1, 9)


def epsilon_blind_fraction(tris, chunk=1 << 20):
    """Share of the mesh's projected (x, y) area that the reference's parity rays cannot see: Moller-Trumbore rejects a triangle
    when |det| < 1e-6, an ABSOLUTE threshold on twice its projected area (cu:117, src/TriRayIntersection.cpp:93; App. A.5).
    On a unit-sized body that is harmless at 1 M triangles (pole slivers only) but not at 10 M: a third to a half of the CAD
    body's triangles fall below it and the reference's own solid fill loses 0.5 % (generic position) to 11 % (axis-aligned)
    of the volume.  The product reproduces that bit for bit; volume arguments only apply where this fraction is small."""
    t = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    blind = total = 0.0
    for a in range(0, len(t), chunk):
        c = t[a:a + chunk]
        e1, e2 = c[:, 1] - c[:, 0], c[:, 2] - c[:, 0]
        det = np.abs(e1[:, 0] * (-e2[:, 1]) + e1[:, 1] * e2[:, 0]).astype(np.float64)
        total += float(det.sum())
        blind += float(det[det < 1e-6].sum())
    return blind / total if total > 0 else 0.0


def check_volume_bracket(tris, grid_size, grid_size2, n23, counts, rel_tol=2e-3):
    """For a closed 2-manifold the solid's volume V is bracketed by the occupancy at both levels:
        inside cells lie wholly inside (centre inside by parity, no triangle touches the box)  ->  inside * v <= V
        every point of the solid lies in an inside or a boundary cell                        ->  V <= (inside + boundary) * v
    and the Level-2 bracket (inside cells count n2^3 voxels each) is nested in the Level-1 one.  rel_tol absorbs f32 cell sizes
    and the reference's inclusive-edge ray semantics (a ray through a shared edge counts twice, App. A.5)."""
    V = mesh_volume(tris)
    v1 = float(np.prod(np.asarray(grid_size, np.float64)))
    lo1, hi1 = counts[0] * v1, (counts[0] + counts[1]) * v1
    assert lo1 <= V * (1 + rel_tol) and V * (1 - rel_tol) <= hi1, ("Level-1 volume bracket", lo1, V, hi1)
    out = {"volume": V, "l1": (lo1, hi1)}
    if n23:
        v2 = float(np.prod(np.asarray(grid_size2, np.float64)))
        lo2 = (counts[0] * n23 + counts[2]) * v2
        hi2 = lo2 + counts[3] * v2
        assert lo2 <= V * (1 + rel_tol) and V * (1 - rel_tol) <= hi2, ("Level-2 volume bracket", lo2, V, hi2)
        assert lo1 <= lo2 * (1 + rel_tol) and hi2 <= hi1 * (1 + rel_tol), ("Level-2 bracket is not nested in Level-1", lo1, lo2, hi2, hi1)
        out["l2"] = (lo2, hi2)
    return out
