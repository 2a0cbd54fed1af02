"""Oracle restatement of the skeleton adjacency (numpy, float64).

Follows the reference: datasets/graph.py:9-44 (edge2mat / normalize_digraph /
get_spatial_graph / Graph), datasets/ntu_rgbd.py:3-35 (NTU edge list, 1-based in the source),
datasets/kinetics.py:24-46 (OpenPose-18 edge list).  Test infrastructure only.
"""
import numpy as np

# (child, parent) pairs, 1-based as printed in datasets/ntu_rgbd.py:5-30.
_NTU_PAIRS_1BASED = (
    (1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9),
    (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18),
    (20, 19), (22, 23), (23, 8), (24, 25), (25, 12),
)
# 0-based, datasets/kinetics.py:26-44.
_KINETICS_PAIRS = (
    (4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2),
    (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14),
)

SKELETONS = {
    "ntu": (25, tuple((a - 1, b - 1) for a, b in _NTU_PAIRS_1BASED)),
    "kinetics": (18, _KINETICS_PAIRS),
}


def _links_to_matrix(links, n):
    """datasets/graph.py:9-13 -- an (i, j) link sets entry [j, i]."""
    m = np.zeros((n, n), dtype=np.float64)
    for src, dst in links:
        m[dst, src] = 1.0
    return m


def _column_normalise(m):
    """datasets/graph.py:16-24 -- right-multiply by diag(1 / column sum), skipping empty columns."""
    col = m.sum(axis=0)
    d = np.zeros((m.shape[1], m.shape[1]), dtype=np.float64)
    for k in range(m.shape[1]):
        if col[k] > 0:
            d[k, k] = col[k] ** (-1)
    return m @ d


def adjacency(name):
    """(3, V, V) float64: self links, normalised inward, normalised outward (graph.py:27-44)."""
    n, inward = SKELETONS[name]
    outward = [(b, a) for a, b in inward]
    eye = _links_to_matrix([(k, k) for k in range(n)], n)
    return np.stack(
        (eye, _column_normalise(_links_to_matrix(inward, n)), _column_normalise(_links_to_matrix(outward, n)))
    )
