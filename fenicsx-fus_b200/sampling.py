"""Point evaluation of a GLL field after the time loop.

Host-side post-processing, the step the reference's examples run right after `rk4`
(`cpp/mwe/parallel_eval_line/main.cpp:47-106`, `cpp/mwe/parallel_eval_surface/main.cpp:50-104`,
`python/src/fenicsxfus/utils.py:10-47`):
find the cell containing each point, then evaluate the tensor-product Lagrange expansion there.

`compute_eval_params(mesh, points)` mirrors the reference helper of the same name: it returns the
points that fall inside the (local) mesh and one containing cell for each; `eval_function`
mirrors `fem::Function::eval`.  Works for any mesh object with `x` (nverts,3) and `xdofmap`
(ncells,8 in tensor vertex order): BoxMesh, HexMesh or a partition's local mesh.
"""
import numpy as np

from . import capi


def _trilinear(X, xi):
    """x(xi) and Jacobian for cells X (n, 2^d, d) at reference points xi (n, d), d = 2 or 3
    (bilinear quadrilaterals, trilinear hexahedra; vertex v = a + 2b [+ 4c])."""
    n, d = X.shape[0], xi.shape[1]
    x = np.zeros((n, d))
    J = np.zeros((n, d, d))
    for v in range(1 << d):
        b = [(v >> k) & 1 for k in range(d)]
        l = [xi[:, k] if b[k] else 1.0 - xi[:, k] for k in range(d)]
        dl = [1.0 if b[k] else -1.0 for k in range(d)]
        w = np.prod(l, axis=0)
        x += w[:, None] * X[:, v]
        g = np.stack([dl[k] * np.prod([l[j] for j in range(d) if j != k], axis=0)
                      for k in range(d)], axis=1)
        J += X[:, v, :, None] * g[:, None, :]
    return x, J


def pull_back(X, pts, iters=30, tol=1e-14):
    """Reference coordinates of physical points pts (n,d) in cells X (n,2^d,d) by Newton."""
    xi = np.full((X.shape[0], pts.shape[1]), 0.5)
    for _ in range(iters):
        x, J = _trilinear(X, xi)
        r = pts - x
        dxi = np.linalg.solve(J, r[:, :, None])[:, :, 0]
        xi += dxi
        if np.abs(dxi).max() < tol:
            break
    return xi


def compute_eval_params(mesh, points, padding=1e-12):
    """points: (3, n) like the reference helper (or (n, 3)).  Returns (points_on_proc (m,3),
    cells (m,), reference coordinates (m,3), indices of the kept points (m,))."""
    pts = np.asarray(points, dtype=np.float64)
    d = getattr(mesh, "dim", 3)                                # 2: quadrilateral meshes (z ignored)
    if pts.ndim != 2 or not ({3, d} & set(pts.shape)):
        raise ValueError("points must be (3, n) or (n, 3)" + (" or (2, n) / (n, 2)" if d == 2 else ""))
    if pts.shape[0] in (3, d) and pts.shape[1] not in (3, d):
        pts = pts.T
    pts = np.ascontiguousarray(pts[:, :d])
    X = mesh.x[mesh.xdofmap][:, :, :d]                         # (nc, 2^d, d)
    lo, hi = X.min(axis=1) - padding, X.max(axis=1) + padding
    # uniform background grid over the mesh bounding box: cells registered in every bin they touch
    glo, ghi = lo.min(axis=0), hi.max(axis=0)
    nc = X.shape[0]
    nb = max(1, int(round(nc ** (1.0 / d))))
    h = np.where(ghi > glo, (ghi - glo) / nb, 1.0)
    b0 = np.clip(((lo - glo) / h).astype(np.int64), 0, nb - 1)
    b1 = np.clip(((hi - glo) / h).astype(np.int64), 0, nb - 1)
    bins = {}
    import itertools
    for c in range(nc):
        for key in itertools.product(*(range(b0[c, k], b1[c, k] + 1) for k in range(d))):
            bins.setdefault(key, []).append(c)
    keep, cells, xis = [], [], []
    pb = np.clip(((pts - glo) / h).astype(np.int64), 0, nb - 1)
    inside_box = np.all((pts >= glo) & (pts <= ghi), axis=1)
    for n in range(pts.shape[0]):
        if not inside_box[n]:
            continue
        cand = [c for c in bins.get(tuple(pb[n]), ())
                if np.all(pts[n] >= lo[c]) and np.all(pts[n] <= hi[c])]
        if not cand:
            continue
        cand = np.array(cand)
        xi = pull_back(X[cand], np.repeat(pts[n][None, :], cand.size, axis=0))
        ok = np.all((xi > -1e-10) & (xi < 1 + 1e-10), axis=1)
        if ok.any():
            k = int(np.flatnonzero(ok)[0])
            keep.append(n)
            cells.append(int(cand[k]))
            xis.append(np.clip(xi[k], 0.0, 1.0))
    keep = np.array(keep, dtype=np.int64)
    return (pts[keep], np.array(cells, dtype=np.int32),
            np.array(xis).reshape(-1, d), keep)


def lagrange_1d(P, s):
    """Values phi_i(s) of the degree-P Lagrange basis on the GLL nodes (Basix order) at s (n,)."""
    pts, wts = np.zeros(P + 1), np.zeros(P + 1)
    capi.check(capi.load().fus_gll(P, pts, wts), "fus_gll")
    s = np.asarray(s, dtype=np.float64)
    out = np.ones((s.size, P + 1))
    for i in range(P + 1):
        for j in range(P + 1):
            if j != i:
                out[:, i] *= (s - pts[j]) / (pts[i] - pts[j])
    return out


def eval_function(V, u, cells, xi):
    """u at reference points xi (m,d) of the given cells: sum_i u[dofmap[c,i]] phi_i0 phi_i1 [phi_i2]
    (fem::Function::eval)."""
    P, N = V.P, V.P + 1
    if xi.shape[1] == 2:
        l0, l1 = (lagrange_1d(P, xi[:, k]) for k in range(2))
        return np.einsum("mab,ma,mb->m", np.asarray(u)[V.dofmap[cells]].reshape(-1, N, N), l0, l1)
    l0, l1, l2 = (lagrange_1d(P, xi[:, d]) for d in range(3))
    coeff = np.asarray(u)[V.dofmap[cells]].reshape(-1, N, N, N)
    return np.einsum("mabc,ma,mb,mc->m", coeff, l0, l1, l2)


def eval_line(V, u, start, end, num_points=100):
    """Sample u on a straight line (the reference's parallel_eval_line example).  Returns
    (points kept, values)."""
    tt = np.linspace(0.0, 1.0, num_points)[:, None]
    pts = (1 - tt) * np.asarray(start, float)[None, :] + tt * np.asarray(end, float)[None, :]
    pk, cells, xi, _ = compute_eval_params(V.mesh, pts)
    return pk, eval_function(V, u, cells, xi)


def eval_surface(V, u, origin, edge_u, edge_v, num_points=100, path=None):
    """Sample u on the parallelogram origin + s*edge_u + t*edge_v, s,t in [0,1], on a
    num_points x num_points grid (the reference's parallel_eval_surface example,
    main.cpp:50-62: a 100 x 100 grid over the plane z = 0 of a rectangle).  Points outside the
    local mesh are dropped.  Returns (points kept, values); with `path` the rows "x,y[,z],value" are
    APPENDED to that text file, as each rank of the reference appends to surface_data.txt
    (main.cpp:91-104)."""
    num = (num_points, num_points) if np.isscalar(num_points) else tuple(num_points)
    o, eu, ev = (np.asarray(a, float) for a in (origin, edge_u, edge_v))
    ss, tt = np.meshgrid(np.linspace(0.0, 1.0, num[0]), np.linspace(0.0, 1.0, num[1]), indexing="xy")
    pts = o[None, :] + ss.reshape(-1, 1) * eu[None, :] + tt.reshape(-1, 1) * ev[None, :]
    if pts.shape[1] == 2:
        pts = np.concatenate([pts, np.zeros((pts.shape[0], 1))], axis=1)
    pk, cells, xi, _ = compute_eval_params(V.mesh, pts)
    vals = eval_function(V, u, cells, xi)
    if path is not None:
        ncoord = 2 if getattr(V.mesh, "dim", 3) == 2 else 3
        with open(path, "a") as f:
            for p, v in zip(pk, vals):
                f.write(",".join(repr(float(c)) for c in p[:ncoord]) + "," + repr(float(v)) + "\n")
    return pk, vals

