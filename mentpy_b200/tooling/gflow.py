"""Generalised flow (Mhalla & Perdrix, arXiv:0709.2670) -- the role of mentpy/mbqc/flow.py:191-251.

Layer by layer from the outputs: a non-output vertex u joins layer k when some set g(u) of already
placed, non-input vertices has odd(g(u)) restricted to the unplaced vertices equal to {u}; that is a
linear system over GF(2) in the adjacency matrix."""
import warnings
from typing import Dict, List, Set

import numpy as np

from .gf2 import gf2_solve


def find_gflow(graph, input_nodes, output_nodes):
    """-> (g, partial_order, depth, layers): g(u) = correction set, partial_order(u, v) = u is measured
    before v, layers[v] = distance from the outputs; (None, None, None, None) when no gflow exists."""
    nodes = list(graph.nodes())
    index = {v: i for i, v in enumerate(nodes)}
    adj = np.zeros((len(nodes), len(nodes)), dtype=np.uint8)
    for a, b in graph.edges():
        adj[index[a], index[b]] = adj[index[b], index[a]] = 1
    inputs, placed = set(input_nodes), set(output_nodes)
    layer: Dict[int, int] = {v: 0 for v in output_nodes}
    g: Dict[int, Set[int]] = {}
    k = 1
    while len(placed) < len(nodes):
        rest = [v for v in nodes if v not in placed]
        cand = [v for v in nodes if v in placed and v not in inputs]
        sub = adj[np.ix_([index[v] for v in rest], [index[v] for v in cand])] if cand else np.zeros((len(rest), 0), np.uint8)
        found = {}
        for r, u in enumerate(rest):
            rhs = np.zeros(len(rest), dtype=np.uint8)
            rhs[r] = 1
            x = gf2_solve(sub, rhs)
            if x is not None:
                found[u] = {cand[i] for i in np.nonzero(x)[0]}
        if not found:
            warnings.warn("No gflow exists for this graph.", UserWarning, stacklevel=2)
            return None, None, None, None
        for u, corr in found.items():
            g[u], layer[u] = corr, k
        placed |= set(found)
        k += 1
    return (lambda v: g[v]), (lambda u, v: layer[u] > layer[v]), max(layer.values()), layer


def verify_gflow(graph, input_nodes, output_nodes, g, layers) -> bool:
    """The three gflow conditions (XY-plane measurements): corrections lie in the future and outside
    the inputs, u is in odd(g(u)), and every other odd neighbour of g(u) is in the future."""
    def odd(s: Set[int]) -> List[int]:
        return [v for v in graph.nodes() if sum(1 for w in graph.neighbors(v) if w in s) % 2]

    for u in graph.nodes():
        if u in output_nodes:
            continue
        corr = g(u)
        if not corr or any(v in input_nodes or not layers[u] > layers[v] for v in corr):
            return False
        o = odd(corr)
        if u not in o or any(v != u and not layers[u] > layers[v] for v in o):
            return False
    return True
