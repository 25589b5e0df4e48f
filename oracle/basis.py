"""GDM 1D basis: Lagrange polynomials through integer nodes (oracle, test-only).

Restates `include/gdm/fe.h:55-336` (`GDM::generate_polynomials_1D`) from its
generator spec `scripts/create_coefficients.py:15-39`: for odd degree p and
variant v in 0..p-1 the cell is [0,1] and the p+1 nodes sit at  k - v,
k = 0..p ; basis k is the Lagrange polynomial that is 1 at node k.
Coefficients are produced exactly (fractions) and handed out highest power
first, i.e. in the order the reference's table is written in (`fe.h:73`).
"""
from fractions import Fraction
import numpy as np


def lagrange_nodes(p, v):
    """Node positions of variant v relative to the cell [0,1] (`create_coefficients.py:22-27`)."""
    return [k - v for k in range(p + 1)]


def _polymul(a, b):
    out = [Fraction(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def lagrange_coefficients(p, v):
    """Exact monomial coefficients, shape [basis k][power high..low] (as written in fe.h)."""
    nodes = lagrange_nodes(p, v)
    table = []
    for k in range(p + 1):
        poly = [Fraction(1)]  # lowest power first while multiplying
        for j in range(p + 1):
            if j == k:
                continue
            d = Fraction(nodes[k] - nodes[j])
            poly = _polymul(poly, [Fraction(-nodes[j]) / d, Fraction(1) / d])
        table.append(list(reversed(poly)))
    return table


def generate_polynomials_1D(p):
    """All variants: list[v][k] of float coefficient arrays, LOWEST power first.

    Mirrors the return value of `GDM::generate_polynomials_1D` (`fe.h:323-333`:
    the table rows are reversed before building `Polynomials::Polynomial`, whose
    coefficient vector is lowest power first).
    """
    assert p % 2 == 1 and p >= 1
    out = []
    for v in range(p):
        coeffs = lagrange_coefficients(p, v)
        out.append([np.array([float(c) for c in reversed(row)]) for row in coeffs])
    return out


def _horner(c_low_first, x, n_der):
    """Value and derivatives of a polynomial by Horner (deal.II `Polynomial::value`)."""
    c = np.array(c_low_first, dtype=float)
    res = []
    for _ in range(n_der + 1):
        acc = np.zeros_like(np.asarray(x, dtype=float))
        for a in c[::-1]:
            acc = acc * x + a
        res.append(acc)
        c = np.array([i * c[i] for i in range(1, len(c))]) if len(c) > 1 else np.array([0.0])
    return res


def basis_values(p, v, x, n_der=1):
    """[der][k][point] values of all p+1 basis functions of variant v at points x."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    polys = generate_polynomials_1D(p)[v]
    out = np.zeros((n_der + 1, p + 1, x.size))
    for k, c in enumerate(polys):
        vals = _horner(c, x, n_der)
        for d in range(n_der + 1):
            out[d, k] = vals[d]
    return out


def gauss_legendre_01(n):
    """`QGauss<1>(n)` on [0,1]: ascending points, weights summing to 1."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w
