"""Edge-list builders for the lattices named in BASELINE.json and used by the reference's
tests/examples.  An edge list is `[((a, b), J), ...]`, exactly what
`QmcIsingGraph::new_with_rng` takes (qmc_ising.rs:131-148)."""
from __future__ import annotations


def one_d_periodic(l: int, j: float = 1.0):
    """tests/convert_test.rs:5-7"""
    return [((i, (i + 1) % l), j) for i in range(l)]


def small_qmc_ring():
    """examples/small_qmc.rs:5"""
    return [((0, 1), -1.0), ((1, 2), 1.0), ((2, 3), 1.0), ((3, 0), 1.0)]


def square_periodic(l: int, j: float = 1.0):
    """examples/crash_check.rs:13-32: f(i,j) = j*l + i; all right bonds, then all down bonds,
    sites enumerated i-major."""
    f = lambda i, jj: jj * l + i
    idx = [(i, jj) for i in range(l) for jj in range(l)]
    right = [((f(i, jj), f((i + 1) % l, jj)), j) for (i, jj) in idx]
    down = [((f(i, jj), f(i, (jj + 1) % l)), j) for (i, jj) in idx]
    return right + down


def two_d_periodic_mixed(l: int):
    """tests/longitudinal_crash.rs:5-23 (right bonds -1, down bonds +1/-1 by column parity)"""
    f = lambda i, jj: jj * l + i
    idx = [(i, jj) for i in range(l) for jj in range(l)]
    right = [((f(i, jj), f((i + 1) % l, jj)), -1.0) for (i, jj) in idx]
    down = [((f(i, jj), f(i, (jj + 1) % l)), 1.0 if i % 2 == 0 else -1.0) for (i, jj) in idx]
    return right + down


def two_unit_cell():
    """tests/longitudinal_crash.rs:25-37"""
    return [((0, 1), -1.0), ((1, 2), 1.0), ((2, 3), 1.0), ((3, 0), 1.0), ((1, 7), 1.0),
            ((4, 5), -1.0), ((5, 6), 1.0), ((6, 7), 1.0), ((7, 4), 1.0)]


def triangular_periodic(l: int, j: float = 1.0):
    """BASELINE.json config #5: bonds (i,j)-(i+1,j), (i,j)-(i,j+1), (i,j)-(i+1,j+1), periodic."""
    f = lambda i, jj: jj * l + i
    idx = [(i, jj) for i in range(l) for jj in range(l)]
    e = [((f(i, jj), f((i + 1) % l, jj)), j) for (i, jj) in idx]
    e += [((f(i, jj), f(i, (jj + 1) % l)), j) for (i, jj) in idx]
    e += [((f(i, jj), f((i + 1) % l, (jj + 1) % l)), j) for (i, jj) in idx]
    return e


def bathroom_unit_cells(l: int):
    """classical/graph.rs:605-625: four-site unit cells (one +1 bond per cell), joined right and down"""
    edges = []
    for x in range(l):
        for y in range(l):
            for i in range(4):
                va, vb = y * l * 4 + x * 4 + i, y * l * 4 + x * 4 + (i + 1) % 4
                edges.append(((min(va, vb), max(va, vb)), 1.0 if i == 0 else -1.0))
            va, vb = y * l * 4 + x * 4 + 1, y * l * 4 + ((x + 1) % l) * 4 + 3
            edges.append(((min(va, vb), max(va, vb)), -1.0))
            va, vb = y * l * 4 + x * 4, ((y + 1) % l) * l * 4 + x * 4 + 2
            edges.append(((min(va, vb), max(va, vb)), -1.0))
    return edges


def nvars_of(edges) -> int:
    """qmc_ising.rs:92"""
    return max(max(a, b) for (a, b), _ in edges) + 1
