"""Insertion-ordered undirected graph used to describe MBQC resource states.

Host-side mirror of the role `GraphState(nx.Graph)` plays in the reference
(mentpy/mbqc/states/graphstate.py:14-68).  The reference inherits every container behaviour from
networkx; the only behaviour the simulator path depends on is *iteration order* (nodes appear in
first-insertion order, relabelling keeps that order) because `trainable_nodes`, tie-breaking in
the measurement order and the stacking helpers are all defined through it
(mentpy/mbqc/mbqcircuit.py:71-87, :110-113, :390-422, :561-611).  This class keeps exactly that
contract with plain dicts and no third-party dependency; any object exposing `nodes()` and
`edges()` (e.g. a networkx graph) can be converted with `GraphState.from_any`.
"""
from typing import Dict, Hashable, Iterable, Iterator, List, Tuple


class GraphState:
    def __init__(self, edges: Iterable[Tuple[Hashable, Hashable]] = ()):
        self._adj: Dict[Hashable, Dict[Hashable, None]] = {}
        self.add_edges_from(edges)

    # -- construction ---------------------------------------------------------------------------
    @classmethod
    def from_any(cls, g) -> "GraphState":
        if isinstance(g, cls):
            return g
        out = cls()
        for v in g.nodes():
            out.add_node(v)
        for e in g.edges():
            out.add_edge(e[0], e[1])
        return out

    def add_node(self, v) -> None:
        if v not in self._adj:
            self._adj[v] = {}

    def add_nodes_from(self, vs) -> None:
        for v in vs:
            self.add_node(v)

    def add_edge(self, u, v) -> None:
        self.add_node(u)
        self.add_node(v)
        self._adj[u][v] = None
        self._adj[v][u] = None

    def add_edges_from(self, edges, **_ignored) -> None:
        for e in edges:
            self.add_edge(e[0], e[1])

    def remove_edge(self, u, v) -> None:
        del self._adj[u][v]
        if u != v:
            del self._adj[v][u]

    def remove_node(self, v) -> None:
        for w in list(self._adj[v]):
            if w != v:
                del self._adj[w][v]
        del self._adj[v]

    def copy(self) -> "GraphState":
        out = GraphState()
        for v, nb in self._adj.items():
            out._adj[v] = dict(nb)
        return out

    # -- queries --------------------------------------------------------------------------------
    def nodes(self) -> List[Hashable]:
        return list(self._adj)

    def edges(self) -> List[Tuple[Hashable, Hashable]]:
        seen, out = set(), []
        for u, nb in self._adj.items():
            for v in nb:
                if v not in seen:
                    out.append((u, v))
            seen.add(u)
        return out

    def neighbors(self, v) -> Iterator[Hashable]:
        return iter(list(self._adj[v]))

    def has_edge(self, u, v) -> bool:
        return u in self._adj and v in self._adj[u]

    def has_node(self, v) -> bool:
        return v in self._adj

    def degree(self, v) -> int:
        return len(self._adj[v])

    def number_of_nodes(self) -> int:
        return len(self._adj)

    def number_of_edges(self) -> int:
        return len(self.edges())

    def subgraph(self, keep) -> "GraphState":
        keep_set = set(keep)
        out = GraphState()
        for v in self._adj:
            if v in keep_set:
                out.add_node(v)
        for u, v in self.edges():
            if u in keep_set and v in keep_set:
                out.add_edge(u, v)
        return out

    def relabeled(self, mapping) -> "GraphState":
        """New graph, same node order, labels pushed through `mapping`."""
        out = GraphState()
        for v in self._adj:
            out.add_node(mapping.get(v, v))
        for u, v in self.edges():
            out.add_edge(mapping.get(u, u), mapping.get(v, v))
        return out

    def index_mapping(self):
        return {v: i for i, v in enumerate(self._adj)}

    def __len__(self) -> int:
        return len(self._adj)

    def __iter__(self):
        return iter(self._adj)

    def __contains__(self, v) -> bool:
        return v in self._adj

    def __repr__(self) -> str:
        return f"GraphState with {self.number_of_nodes()} nodes and {self.number_of_edges()} edges."


def disjoint_union(g: GraphState, h: GraphState) -> GraphState:
    """Integer-relabel both graphs in their node order (g first) and take the union."""
    out = g.relabeled({v: i for i, v in enumerate(g.nodes())})
    off = len(g)
    hh = h.relabeled({v: off + i for i, v in enumerate(h.nodes())})
    for v in hh.nodes():
        out.add_node(v)
    for u, v in hh.edges():
        out.add_edge(u, v)
    return out


def contract_into(g: GraphState, keep, gone) -> GraphState:
    """Copy of g with node `gone` merged into `keep` (no self loops); node order otherwise kept."""
    out = g.copy()
    moved = [w for w in g._adj[gone]]
    out.remove_node(gone)
    for w in moved:
        if w == keep or w == gone:
            continue
        out.add_edge(keep, w)
    return out
