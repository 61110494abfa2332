"""Five-point essential-matrix solver -- CPU oracle for ``cvFivePoint``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the reference's vendored solver
``src/MiniCVNative/fivepoint.cpp:233-339`` (exported as ``cvFivePoint``,
``MiniCVNative.cpp:368-382``; the same algorithm sits behind ``findEssentialMat``):

  1. Q (5x9) from the correspondences, column order as ``fivepoint.cpp:241-249``;
     null-space basis = rows 5..8 of Vt of OpenCV's full SVD (``:251-254``) -- restated
     with OpenCV's Jacobi SVD incl. its pseudo-random completion (oracle/linalg_cv.py), so
     the basis, and therefore the SIGN of every returned E, matches the reference.
  2. the ten cubic constraints det(E) = 0 and 2 E E^T E - tr(E E^T) E = 0 on
     E = x X + y Y + z Z + W (``getCoeffMat``, ``:10-231``).  The reference expands them
     through a generated 200-row table; here they are formed by polynomial arithmetic,
     which yields the same polynomials (checked numerically in the tests).
  3. Gauss-Jordan on the ten leading monomials (``:260``: inv * tail), the 3x3 polynomial
     matrix B(z) from rows 4..9 (``:262-281``), det B(z) = degree-10 polynomial
     (``:283-294``), its real roots (``:296-303``: |Im| <= 1e-10).
  4. per root: null vector of B(z) -> (x, y) (``:316-327``), E normalised to unit
     Frobenius norm (``:329-333``).
The order of the returned matrices follows numpy's root order, not cv::solvePoly's; tests
compare as sets.
"""
from __future__ import annotations

import numpy as np

from . import linalg_cv

# monomials of degree <= 3 in (x, y, z) as exponent triples, in Nister's elimination order:
# ten leading monomials, then [x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1]
MONOMIALS = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0),
             (1, 1, 1), (1, 1, 0),
             (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2),
             (0, 0, 1), (0, 0, 0)]
MONO_INDEX = {m: i for i, m in enumerate(MONOMIALS)}


class Poly:
    """Sparse polynomial in (x, y, z): dict exponent-triple -> coefficient."""
    __slots__ = ("c",)

    def __init__(self, c=None):
        self.c = dict(c) if c else {}

    def __add__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) + v
        return Poly(r)

    def __sub__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) - v
        return Poly(r)

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly({k: v * o for k, v in self.c.items()})
        r = {}
        for (a, b, c), v in self.c.items():
            for (d, e, f), w in o.c.items():
                k = (a + d, b + e, c + f)
                r[k] = r.get(k, 0.0) + v * w
        return Poly(r)

    __rmul__ = __mul__

    def vector(self):
        out = np.zeros(20)
        for k, v in self.c.items():
            out[MONO_INDEX[k]] = v
        return out


def null_space_basis(pa: np.ndarray, pb: np.ndarray) -> np.ndarray:
    """Rows 5..8 of Vt of the 5x9 constraint matrix -> EE (4 x 9), reference :236-254."""
    q1 = np.asarray(pa, np.float64).reshape(-1, 2)
    q2 = np.asarray(pb, np.float64).reshape(-1, 2)
    n = q1.shape[0]
    Q = np.empty((n, 9))
    Q[:, 0] = q1[:, 0] * q2[:, 0]
    Q[:, 1] = q1[:, 1] * q2[:, 0]
    Q[:, 2] = q2[:, 0]
    Q[:, 3] = q1[:, 0] * q2[:, 1]
    Q[:, 4] = q1[:, 1] * q2[:, 1]
    Q[:, 5] = q2[:, 1]
    Q[:, 6] = q1[:, 0]
    Q[:, 7] = q1[:, 1]
    Q[:, 8] = 1.0
    _, _, vt = linalg_cv.svd_full(Q)
    return vt[5:9].copy()


def constraint_matrix(EE: np.ndarray) -> np.ndarray:
    """10 x 20 coefficient matrix of the cubic constraints over MONOMIALS."""
    X, Y, Z, W = [EE[i].reshape(3, 3) for i in range(4)]
    x = Poly({(1, 0, 0): 1.0})
    y = Poly({(0, 1, 0): 1.0})
    z = Poly({(0, 0, 1): 1.0})
    one = Poly({(0, 0, 0): 1.0})
    E = [[x * X[i, j] + y * Y[i, j] + z * Z[i, j] + one * W[i, j] for j in range(3)] for i in range(3)]
    det = (E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1])
           - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0])
           + E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]))
    EEt = [[E[i][0] * E[j][0] + E[i][1] * E[j][1] + E[i][2] * E[j][2] for j in range(3)] for i in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    rows = [det.vector()]
    for i in range(3):
        for j in range(3):
            p = (EEt[i][0] * E[0][j] + EEt[i][1] * E[1][j] + EEt[i][2] * E[2][j]) * 2.0 - tr * E[i][j]
            rows.append(p.vector())
    return np.array(rows)


def _eliminate(A: np.ndarray) -> np.ndarray:
    """inv(A[:, :10]) @ A[:, 10:] (reference :260)."""
    return np.linalg.solve(A[:, :10], A[:, 10:])


def action_polynomials(C: np.ndarray):
    """B(z): 3 rows x [x-coeffs(4), y-coeffs(4), 1-coeffs(5)] in descending powers of z
    (reference :262-281), from rows (4,5), (6,7), (8,9) of the eliminated system."""
    B = np.zeros((3, 13))
    for i in range(3):
        r1 = C[2 * i + 4]
        r2 = C[2 * i + 5]
        row1 = np.zeros(13)
        row2 = np.zeros(13)
        row1[1:4] = r1[0:3]
        row1[5:8] = r1[3:6]
        row1[9:13] = r1[6:10]
        row2[0:3] = r2[0:3]
        row2[4:7] = r2[3:6]
        row2[8:12] = r2[6:10]
        B[i] = row1 - row2
    return B


def det_polynomial(B: np.ndarray) -> np.ndarray:
    """det B(z) as 11 coefficients, highest power first."""
    p = [[B[i, 0:4], B[i, 4:8], B[i, 8:13]] for i in range(3)]
    mul = np.polymul
    sub = np.polysub
    add = np.polyadd
    d = add(sub(mul(p[0][0], sub(mul(p[1][1], p[2][2]), mul(p[1][2], p[2][1]))),
                mul(p[0][1], sub(mul(p[1][0], p[2][2]), mul(p[1][2], p[2][0])))),
            mul(p[0][2], sub(mul(p[1][0], p[2][1]), mul(p[1][1], p[2][0]))))
    out = np.zeros(11)
    out[11 - len(d):] = d
    return out


def five_point(pa: np.ndarray, pb: np.ndarray) -> np.ndarray:
    """All real essential matrices through 5 correspondences: array [k, 3, 3], unit Frobenius."""
    EE = null_space_basis(pa, pb)
    A = constraint_matrix(EE)
    C = _eliminate(A)
    B = action_polynomials(C)
    coeffs = det_polynomial(B)
    roots = np.roots(coeffs)
    out = []
    for r in roots:
        if abs(r.imag) > 1e-10:
            continue
        z1 = r.real
        z2, z3, z4 = z1 * z1, z1 ** 3, z1 ** 4
        bz = np.empty((3, 3))
        for j in range(3):
            br = B[j]
            bz[j, 0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3]
            bz[j, 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7]
            bz[j, 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12]
        xy1 = linalg_cv.solve_z(bz)
        if abs(xy1[2]) < 1e-10:
            continue
        xs = xy1[0] / xy1[2]
        ys = xy1[1] / xy1[2]
        ev = EE[0] * xs + EE[1] * ys + EE[2] * z1 + EE[3]
        ev = ev / np.linalg.norm(ev)
        out.append(ev.reshape(3, 3))
    return np.array(out).reshape(-1, 3, 3)
