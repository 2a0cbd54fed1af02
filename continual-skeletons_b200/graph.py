"""Skeleton graphs -> 3-partition adjacency ``A (3, V, V)`` float64.

Same result, bit for bit, as the reference's ``datasets/graph.py:9-44`` with the edge lists of
``datasets/ntu_rgbd.py:3-35`` (25 joints) and ``datasets/kinetics.py:24-46`` (18 joints): partition
0 = self links, 1 = column-normalised inward links, 2 = column-normalised outward links, where a
link (i, j) sets entry [j, i].  Built here from integer (source, target) index arrays.
"""
import numpy as np

# NTU RGB+D: joint -> joint it points to (toward the spine), 1-based joint numbers.
_NTU_TOWARD = {
    1: 2, 2: 21, 3: 21, 4: 3, 5: 21, 6: 5, 7: 6, 8: 7, 9: 21, 10: 9, 11: 10, 12: 11, 13: 1,
    14: 13, 15: 14, 16: 15, 17: 1, 18: 17, 19: 18, 20: 19, 22: 23, 23: 8, 24: 25, 25: 12,
}
# OpenPose-18 (Kinetics-skeleton), 0-based.
_KINETICS_TOWARD = [
    (4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2), (5, 1),
    (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14),
]


class Graph:
    """``graph.A`` as the reference exposes it (datasets/graph.py:35-44)."""

    def __init__(self, inward, num_node):
        self.num_node = int(num_node)
        self.inward = [(int(a), int(b)) for a, b in inward]
        self.outward = [(b, a) for a, b in self.inward]
        self.self_link = [(k, k) for k in range(self.num_node)]
        self.neighbor = self.inward + self.outward
        self.A = np.stack([self._partition(self.self_link, False), self._partition(self.inward, True),
                           self._partition(self.outward, True)])

    def _partition(self, links, normalise):
        n = self.num_node
        src = np.fromiter((a for a, _ in links), dtype=np.int64, count=len(links))
        dst = np.fromiter((b for _, b in links), dtype=np.int64, count=len(links))
        m = np.zeros((n, n), dtype=np.float64)
        m[dst, src] = 1.0
        if normalise:
            col = m.sum(axis=0)
            inv = np.zeros(n, dtype=np.float64)
            nz = col > 0
            inv[nz] = col[nz] ** (-1)
            m = m @ np.diag(inv)
        return m


def ntu_graph():
    return Graph([(a - 1, b - 1) for a, b in _NTU_TOWARD.items()], 25)


def kinetics_graph():
    return Graph(_KINETICS_TOWARD, 18)
