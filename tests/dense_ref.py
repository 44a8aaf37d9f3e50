"""Independent dense O(N^6) evaluation of the GLL-quadrature weak forms (test helper).

Builds element matrices directly from the definition with numpy polynomials -- no sum
factorisation, no shared tables with the oracle -- in the spirit of the reference's own
acceptance method "SF operator vs an independent evaluation of the same weak form"
(cpp/fenicsx-sf/tests/test_operators3d/main.cpp:100-166).
"""
import numpy as np
from numpy.polynomial import legendre as L
from numpy.polynomial import polynomial as Pn


def gll_basix_order(m):
    """GLL nodes/weights on [0,1], order [0,1,interior...]: roots of (1-x^2) P'_{m-1}."""
    n = m - 1
    c = np.zeros(n + 1)
    c[n] = 1.0
    interior = np.sort(np.real(L.legroots(L.legder(c))))
    x = np.concatenate([[-1.0], interior, [1.0]])
    w = 2.0 / (n * (n + 1) * L.legval(x, c) ** 2)
    x01, w01 = 0.5 * (x + 1), 0.5 * w
    order = [0, m - 1] + list(range(1, m - 1))
    return x01[order], w01[order]


def lagrange_tables(pts):
    """phi[q,i], dphi[q,i] of the Lagrange basis on pts evaluated at pts, via monomial coefficients."""
    N = len(pts)
    phi = np.zeros((N, N))
    dphi = np.zeros((N, N))
    for i in range(N):
        others = np.delete(pts, i)
        coef = Pn.polyfromroots(others) / np.prod(pts[i] - others)
        phi[:, i] = Pn.polyval(pts, coef)
        dphi[:, i] = Pn.polyval(pts, Pn.polyder(coef))
    return phi, dphi


def element_matrices(P, X, coeff):
    """Dense element stiffness K_e (Nd x Nd) and lumped mass diagonal for one trilinear cell.
    X: (8,3) vertices in tensor order (x fastest)."""
    N = P + 1
    pts, wts = gll_basix_order(N)
    phi1, dphi1 = lagrange_tables(pts)
    Nd = N ** 3
    # reference gradients of the 3-D basis at all points: grad[q, i, d]
    grad = np.zeros((Nd, Nd, 3))
    for q0 in range(N):
        for q1 in range(N):
            for q2 in range(N):
                q = (q0 * N + q1) * N + q2
                a0, a1, a2 = phi1[q0], phi1[q1], phi1[q2]
                d0, d1, d2 = dphi1[q0], dphi1[q1], dphi1[q2]
                grad[q, :, 0] = np.einsum("a,b,c->abc", d0, a1, a2).reshape(-1)
                grad[q, :, 1] = np.einsum("a,b,c->abc", a0, d1, a2).reshape(-1)
                grad[q, :, 2] = np.einsum("a,b,c->abc", a0, a1, d2).reshape(-1)
    K = np.zeros((Nd, Nd))
    mdiag = np.zeros(Nd)
    for q0 in range(N):
        for q1 in range(N):
            for q2 in range(N):
                q = (q0 * N + q1) * N + q2
                xi = np.array([pts[q0], pts[q1], pts[q2]])
                J = np.zeros((3, 3))
                for v in range(8):
                    b = [(v >> d) & 1 for d in range(3)]
                    for j in range(3):
                        g = 1.0
                        for d in range(3):
                            if d == j:
                                g *= 1.0 if b[d] else -1.0
                            else:
                                g *= xi[d] if b[d] else 1 - xi[d]
                        J[:, j] += X[v] * g
                w = wts[q0] * wts[q1] * wts[q2] * abs(np.linalg.det(J))
                Jinv = np.linalg.inv(J)
                pg = grad[q] @ Jinv          # physical gradients (Nd,3)
                K += coeff * w * (pg @ pg.T)
                mdiag[q] = w
    return K, mdiag


def dense_stiffness_apply(P, xg, xd, dofmap, coeffs, x):
    y = np.zeros_like(x)
    for c in range(xd.shape[0]):
        K, _ = element_matrices(P, xg[xd[c]], coeffs[c])
        y[dofmap[c]] += K @ x[dofmap[c]]
    return y


def element_matrices_2d(P, X, coeff):
    """Dense element stiffness (Nd x Nd, Nd = N^2) and lumped mass diagonal of one bilinear
    quadrilateral; X: (4,2) vertices in tensor order v = a + 2b."""
    N = P + 1
    pts, wts = gll_basix_order(N)
    phi1, dphi1 = lagrange_tables(pts)
    Nd = N * N
    K = np.zeros((Nd, Nd))
    mdiag = np.zeros(Nd)
    for q0 in range(N):
        for q1 in range(N):
            q = q0 * N + q1
            grad = np.stack([np.outer(dphi1[q0], phi1[q1]).reshape(-1),
                             np.outer(phi1[q0], dphi1[q1]).reshape(-1)], axis=1)      # (Nd, 2)
            xi = (pts[q0], pts[q1])
            J = np.zeros((2, 2))
            for v in range(4):
                b = (v & 1, v >> 1)
                l = [xi[d] if b[d] else 1 - xi[d] for d in range(2)]
                s = [1.0 if b[d] else -1.0 for d in range(2)]
                J[:, 0] += X[v] * s[0] * l[1]
                J[:, 1] += X[v] * l[0] * s[1]
            w = wts[q0] * wts[q1] * abs(np.linalg.det(J))
            pg = grad @ np.linalg.inv(J)
            K += coeff * w * (pg @ pg.T)
            mdiag[q] = w
    return K, mdiag
