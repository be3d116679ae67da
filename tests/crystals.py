"""
Crystal structures with known Madelung constants -- the reference's own known-answer set
(values and structures: /root/reference tests/helpers.py:19-139; literature:
doi:10.1021/ic2023852).  Returned as numpy: positions, charges (N,1), cell, madelung, n_formula.
"""
import math

import numpy as np

S3 = math.sqrt(3)


def crystal(name):
    if name == "CsCl":
        return (np.array([[0, 0, 0], [0.5, 0.5, 0.5]]), [-1.0, 1.0], np.eye(3), 2.0353610945260, 1)
    if name == "NaCl_primitive":
        return (np.array([[0.0, 0, 0], [1.0, 0, 0]]), [1.0, -1.0],
                np.array([[0, 1.0, 1], [1, 0, 1], [1, 1, 0]]), 1.7475645946, 1)
    if name == "NaCl_cubic":
        pos = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]])
        return pos, [1.0, -1, -1, -1, 1, 1, 1, -1], 2 * np.eye(3), 1.7475645946, 4
    if name == "zincblende":
        return (np.array([[0, 0, 0], [0.5, 0.5, 0.5]]), [1.0, -1],
                np.array([[0, 1.0, 1], [1, 0, 1], [1, 1, 0]]), 2 * 1.6380550533 / S3, 1)
    if name == "wurtzite":
        u = 3 / 8
        c = math.sqrt(1 / u)
        pos = np.array([[0.5, 0.5 / S3, 0.0], [0.5, 0.5 / S3, u * c], [0.5, -0.5 / S3, 0.5 * c],
                        [0.5, -0.5 / S3, (0.5 + u) * c]])
        cell = np.array([[0.5, -0.5 * S3, 0], [0.5, 0.5 * S3, 0], [0, 0, c]])
        return pos, [1.0, -1, 1, -1], cell, 1.64132 / (u * c), 2
    if name == "fluorite":
        pos = np.array([[1 / 4, 1 / 4, 1 / 4], [3 / 4, 3 / 4, 3 / 4], [0, 0, 0]])
        return pos, [-1.0, -1, 2], np.array([[1.0, 1, 0], [1, 0, 1], [0, 1, 1]]) / 2.0, 11.6365752270768, 1
    if name == "cu2o":
        pos = np.array([[0, 0, 0], [1 / 2, 1 / 2, 1 / 2], [1 / 4, 1 / 4, 1 / 4], [1 / 4, 3 / 4, 3 / 4],
                        [3 / 4, 1 / 4, 3 / 4], [3 / 4, 3 / 4, 1 / 4]])
        return pos, [-2.0, -2, 1, 1, 1, 1], np.eye(3), 10.2594570330750, 2
    raise ValueError(name)


NAMES = ["CsCl", "NaCl_primitive", "NaCl_cubic", "zincblende", "wurtzite", "cu2o", "fluorite"]


def get(name, scale=1.0):
    pos, q, cell, madelung, units = crystal(name)
    return (np.asarray(pos, dtype=np.float64) * scale, np.asarray(q, dtype=np.float64).reshape(-1, 1),
            np.asarray(cell, dtype=np.float64) * scale, madelung / scale, units)
