"""
Point-charge "molecules" with closed-form Coulomb potentials (the known answers of the reference's
tests/calculators/test_values_direct.py:28-121): unit-edge dimer, equilateral triangle, square,
regular tetrahedron, each with alternating, all-positive and all-negative charges.
V_i = sum_{j != i} q_j / r_ij.
"""
import math

import numpy as np

_S2, _S3 = math.sqrt(2.0), math.sqrt(3.0)

GEOMETRIES = {
    "dimer": [[0, 0, 0], [1, 0, 0]],
    "triangle": [[0, 0, 0], [1, 0, 0], [0.5, _S3 / 2, 0]],
    "square": [[0.5, 0.5, 0], [0.5, -0.5, 0], [-0.5, 0.5, 0], [-0.5, -0.5, 0]],
    "tetrahedron": [[0, 0, 0], [1, 0, 0], [0.5, _S3 / 2, 0], [0.5, _S3 / 6, _S2 / _S3]],
}
ALTERNATING = {"dimer": [1, -1], "triangle": [1, -1, 0], "square": [1, -1, -1, 1], "tetrahedron": [1, -1, 1, -1]}
# closed forms for the alternating charges; uniform charges: (n-1)/edge with the square's diagonal
EXACT_ALTERNATING = {
    "dimer": [-1, 1], "triangle": [-1, 1, 0],
    "square": [q * (1 / _S2 - 2) for q in (1, -1, -1, 1)], "tetrahedron": [-1, 1, -1, 1],
}
EXACT_POSITIVE = {"dimer": 1.0, "triangle": 2.0, "square": 2 + 1 / _S2, "tetrahedron": 3.0}


def molecule(name: str, sign: str):
    """positions (n,3), charges (n,1), exact potentials (n,1) as float64 numpy arrays"""
    pos = np.asarray(GEOMETRIES[name], dtype=np.float64)
    n = len(pos)
    if sign == "alternating":
        q, v = np.asarray(ALTERNATING[name], dtype=np.float64), np.asarray(EXACT_ALTERNATING[name], dtype=np.float64)
    else:
        s = 1.0 if sign == "positive" else -1.0
        q, v = s * np.ones(n), s * EXACT_POSITIVE[name] * np.ones(n)
    return pos, q.reshape(-1, 1), v.reshape(-1, 1)


def rotations():
    phi, theta = 0.82321, 1.23456
    rz = np.array([[math.cos(phi), -math.sin(phi), 0], [math.sin(phi), math.cos(phi), 0], [0, 0, 1.0]])
    ry = np.array([[math.cos(theta), 0, math.sin(theta)], [0, 1.0, 0], [-math.sin(theta), 0, math.cos(theta)]])
    eye = np.eye(3)
    return [eye, rz, ry @ rz, -eye, -(ry @ rz)]


def all_pairs(pos: np.ndarray, full: bool):
    """every pair once (half list) or in both orders (full list), with distances"""
    n = len(pos)
    pairs = [(i, j) for i in range(n) for j in range(n) if (i != j if full else i < j)]
    idx = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    d = np.linalg.norm(pos[idx[:, 1]] - pos[idx[:, 0]], axis=1)
    return idx, d
