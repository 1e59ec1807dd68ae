"""Deterministic synthetic meshes for the benchmark configs and the parity tests.

All generators return (V float32 [nv,3], F int32 [nf,3]) closed 2-manifolds with outward winding; vertices are rounded to
float32 before any use.  write_obj/write_off print with %.9g so that the reference-semantics loaders (strtof) read back the
exact same float32 values.  Harness code (numpy): mesh synthesis is not on the hot path.
"""
import numpy as np

SEED_BASE = 20240607


def uv_sphere(n_lon=1000, n_rings=501, r=1.0):
    """UV sphere: n_lon longitudes, n_rings rings incl. the two poles -> 2*n_lon*(n_rings-2) triangles (1000 x 501 -> 998,000;
    use n_rings=502 for exactly 1,000,000)."""
    th = np.linspace(0.0, np.pi, n_rings)[1:-1]
    ph = np.arange(n_lon) * (2.0 * np.pi / n_lon)
    T, P = np.meshgrid(th, ph, indexing="ij")
    V = np.stack([r * np.sin(T) * np.cos(P), r * np.sin(T) * np.sin(P), r * np.cos(T)], -1).reshape(-1, 3)
    V = np.concatenate([[[0, 0, r]], V, [[0, 0, -r]]]).astype(np.float32)
    nr = n_rings - 2
    idx = lambda i, j: 1 + i * n_lon + (j % n_lon)
    j = np.arange(n_lon)
    F = [np.stack([np.zeros(n_lon, int), idx(0, j), idx(0, j + 1)], -1)]
    for i in range(nr - 1):
        a, b, c, d = idx(i, j), idx(i, j + 1), idx(i + 1, j), idx(i + 1, j + 1)
        F.append(np.stack([a, c, d], -1)); F.append(np.stack([a, d, b], -1))
    south = 1 + nr * n_lon
    F.append(np.stack([np.full(n_lon, south), idx(nr - 1, j + 1), idx(nr - 1, j)], -1))
    return V, np.concatenate(F).astype(np.int32)


def torus(n_major=1000, n_minor=500, R=1.0, r=0.4):
    """Torus: n_major x n_minor quads -> 2*n_major*n_minor triangles (1000 x 500 -> 1,000,000)."""
    u = np.arange(n_major) * (2 * np.pi / n_major)
    v = np.arange(n_minor) * (2 * np.pi / n_minor)
    U, W = np.meshgrid(u, v, indexing="ij")
    V = np.stack([(R + r * np.cos(W)) * np.cos(U), (R + r * np.cos(W)) * np.sin(U), r * np.sin(W)], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n_major), np.arange(n_minor), indexing="ij")
    idx = lambda a, b: ((a % n_major) * n_minor + (b % n_minor)).reshape(-1)
    a, b, c, d = idx(i, j), idx(i + 1, j), idx(i, j + 1), idx(i + 1, j + 1)
    F = np.concatenate([np.stack([a, b, d], -1), np.stack([a, d, c], -1)]).astype(np.int32)
    return V, F


def cad_body(nu=2500, nv=2000, seed=SEED_BASE, amp=0.05):
    """"NURBS-tessellated CAD" stand-in: a genus-0 body whose radius is a smooth tensor-product B-spline-like perturbation
    (periodic cubic blend of a 12 x 8 control lattice, U(-amp,amp)*r from the seed) evaluated on an nu x nv parameter grid
    -> 2*nu*(nv-1) triangles (2500 x 2001 -> 10,000,000)."""
    rng = np.random.default_rng(seed)
    cu, cv = 12, 8
    ctrl = rng.uniform(-amp, amp, (cu, cv))

    def blend(t, n):  # periodic uniform cubic B-spline basis evaluated at t in [0,n)
        i = np.floor(t).astype(int); f = t - i
        b = np.stack([(1 - f) ** 3, 3 * f ** 3 - 6 * f ** 2 + 4, -3 * f ** 3 + 3 * f ** 2 + 3 * f + 1, f ** 3], -1) / 6.0
        return i, b

    u = np.arange(nu) * (cu / nu)
    v = np.linspace(0.0, cv - 3.0, nv + 1)
    iu, bu = blend(u, cu)
    iv, bv = blend(v, cv)
    Ru = np.zeros((nu, cv))
    for k in range(4):
        Ru += bu[:, k:k + 1] * ctrl[(iu + k) % cu, :]
    Rad = np.zeros((nu, nv + 1))
    for k in range(4):
        Rad += Ru[:, np.minimum(iv + k, cv - 1)] * bv[None, :, k]
    th = np.linspace(0.0, np.pi, nv + 1)
    ph = np.arange(nu) * (2 * np.pi / nu)
    rr = 1.0 + Rad
    rr[:, 0] = rr[:, 0].mean(); rr[:, -1] = rr[:, -1].mean()
    X = rr * np.sin(th)[None, :] * np.cos(ph)[:, None]
    Y = rr * np.sin(th)[None, :] * np.sin(ph)[:, None]
    Z = 1.3 * rr * np.cos(th)[None, :]
    ring = np.stack([X[:, 1:-1], Y[:, 1:-1], Z[:, 1:-1]], -1)  # [nu, nv-1, 3]
    V = np.concatenate([[[0, 0, Z[0, 0]]], ring.transpose(1, 0, 2).reshape(-1, 3), [[0, 0, Z[0, -1]]]]).astype(np.float32)
    nr = nv - 1
    idx = lambda i, j: 1 + i * nu + (j % nu)
    j = np.arange(nu)
    F = [np.stack([np.zeros(nu, int), idx(0, j), idx(0, j + 1)], -1)]
    for i in range(nr - 1):
        a, b, c, d = idx(i, j), idx(i, j + 1), idx(i + 1, j), idx(i + 1, j + 1)
        F.append(np.stack([a, c, d], -1)); F.append(np.stack([a, d, b], -1))
    south = 1 + nr * nu
    F.append(np.stack([np.full(nu, south), idx(nr - 1, j + 1), idx(nr - 1, j)], -1))
    return V, np.concatenate(F).astype(np.int32)


def drilled_block(seed=SEED_BASE, n_seg=48, n_grid=14):
    """Watertight block with 1-4 through-holes along z (config 5: ~5k triangles with the defaults).  The top and bottom faces
    are triangulated as a structured n_grid x n_grid lattice whose cells inside a hole are removed and stitched to the hole
    rim, which keeps the mesh a closed 2-manifold without a general polygon triangulator."""
    rng = np.random.default_rng(seed)
    sx, sy, sz = rng.uniform(0.8, 1.2, 3)
    nh = int(rng.integers(1, 5))
    # holes on distinct lattice-aligned square pads so that they never overlap
    pads = [(i, j) for i in range(2) for j in range(2)]
    rng.shuffle(pads)
    V, F = [], []

    def add(v):
        V.append(v); return len(V) - 1

    m = n_grid // 2  # lattice cells per pad side
    for (pi, pj) in pads:
        x0, y0 = -sx / 2 + pi * sx / 2, -sy / 2 + pj * sy / 2
        w, h = sx / 2, sy / 2
        hole = len([p for p in pads[:nh] if p == (pi, pj)]) > 0
        if hole:
            rad = rng.uniform(0.12, 0.3) * min(w, h)
            cx0, cy0 = x0 + w / 2 + rng.uniform(-0.1, 0.1) * w, y0 + h / 2 + rng.uniform(-0.1, 0.1) * h
        for zsign, z in ((1, sz / 2), (-1, -sz / 2)):
            if not hole:
                ids = [[add((x0 + w * a / m, y0 + h * b / m, z)) for b in range(m + 1)] for a in range(m + 1)]
                for a in range(m):
                    for b in range(m):
                        q = [ids[a][b], ids[a + 1][b], ids[a + 1][b + 1], ids[a][b + 1]]
                        F.extend([(q[0], q[1], q[2]), (q[0], q[2], q[3])] if zsign > 0 else [(q[0], q[2], q[1]), (q[0], q[3], q[2])])
            else:
                # square pad boundary (4*m points, counter-clockwise) fanned to the hole rim (n_seg points)
                bnd = [(x0 + w * a / m, y0) for a in range(m)] + [(x0 + w, y0 + h * b / m) for b in range(m)] + \
                      [(x0 + w - w * a / m, y0 + h) for a in range(m)] + [(x0, y0 + h - h * b / m) for b in range(m)]
                bi = [add((p[0], p[1], z)) for p in bnd]
                ang0 = np.arctan2(bnd[0][1] - cy0, bnd[0][0] - cx0)
                ri = [add((cx0 + rad * np.cos(ang0 + 2 * np.pi * k / n_seg), cy0 + rad * np.sin(ang0 + 2 * np.pi * k / n_seg), z)) for k in range(n_seg)]
                nb_, a, b = len(bi), 0, 0
                # advance along whichever loop is "behind" in angle; produces a closed triangle strip between the two loops
                angb = np.unwrap([np.arctan2(p[1] - cy0, p[0] - cx0) for p in bnd] + [np.arctan2(bnd[0][1] - cy0, bnd[0][0] - cx0)])
                angr = ang0 + 2 * np.pi * np.arange(n_seg + 1) / n_seg
                if angb[-1] < angb[0]:
                    angb = angb + 0
                while a < nb_ or b < n_seg:
                    if b >= n_seg or (a < nb_ and angb[a + 1] <= angr[b + 1]):
                        t = (bi[a % nb_], bi[(a + 1) % nb_], ri[b % n_seg]); a += 1
                    else:
                        t = (ri[b % n_seg], bi[a % nb_], ri[(b + 1) % n_seg]); b += 1
                    F.append(t if zsign > 0 else (t[0], t[2], t[1]))
                if zsign > 0:
                    top_r = ri
                else:
                    bot_r = ri
        if hole:  # cylinder wall, normals pointing into the hole (outward from the solid)
            for k in range(n_seg):
                a, b, c, d = top_r[k], top_r[(k + 1) % n_seg], bot_r[k], bot_r[(k + 1) % n_seg]
                F.extend([(a, b, d), (a, d, c)])
    # outer side walls: stitch the top and bottom outer boundaries (lattice points on the block's perimeter)
    V = np.array(V, np.float64)
    F = np.array(F, np.int64)
    V, F = _weld(V, F)
    per_top = _perimeter(V, sx, sy, sz / 2)
    per_bot = _perimeter(V, sx, sy, -sz / 2)
    side = []
    n = len(per_top)
    for k in range(n):
        a, b, c, d = per_top[k], per_top[(k + 1) % n], per_bot[k], per_bot[(k + 1) % n]
        side.extend([(a, c, d), (a, d, b)])
    F = np.concatenate([F, np.array(side, np.int64)])
    return V.astype(np.float32), F.astype(np.int32)


def _weld(V, F, tol=1e-9):
    key = np.round(V / tol).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    return V[first], inv.reshape(-1)[F]


def _perimeter(V, sx, sy, z):
    on = np.where(np.abs(V[:, 2] - z) < 1e-9)[0]
    edge = on[(np.abs(np.abs(V[on, 0]) - sx / 2) < 1e-9) | (np.abs(np.abs(V[on, 1]) - sy / 2) < 1e-9)]
    ang = np.arctan2(V[edge, 1] / sy, V[edge, 0] / sx)
    return list(edge[np.argsort(ang)])


def triangles(V, F):
    """flat [nf, 9] float32 in the reference's layout (Object::CreateFlatTriangleData)."""
    return np.ascontiguousarray(V[F].reshape(-1, 9), np.float32)


def write_obj(path, V, F):
    with open(path, "w") as f:
        for v in V:
            f.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for t in F:
            f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def write_off(path, V, F):
    with open(path, "w") as f:
        f.write("OFF\n%d %d 0\n" % (len(V), len(F)))
        for v in V:
            f.write("%.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for t in F:
            f.write("3 %d %d %d\n" % (t[0], t[1], t[2]))


def is_closed_manifold(F):
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    fwd = {}
    for a, b in e:
        fwd[(a, b)] = fwd.get((a, b), 0) + 1
    return all(v == 1 and fwd.get((b, a), 0) == 1 for (a, b), v in fwd.items())
