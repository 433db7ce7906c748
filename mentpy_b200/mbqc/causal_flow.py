"""Causal flow (Mhalla-Perdrix) and the layer structure derived from it.

Host-side, integer-exact mirror of mentpy/mbqc/flow.py:115-185 (`find_cflow`, `causal_flow_aux`)
and of the layer / order bookkeeping in `Flow.__init__` (flow.py:61-85).  The measurement order,
and therefore every slot index the CUDA plan uses, is derived from the layer numbers computed
here, so the candidate sets are built with the same Python set operations in the same sequence as
the reference (set iteration order can pick between equivalent flow edges in degenerate graphs).
gflow / pflow (flow.py:191-262) are not on the simulator path of any supported template and are
out of scope.
"""
from typing import Callable, Dict, List, Optional, Tuple


def find_cflow(graph, input_nodes, output_nodes):
    """Returns (flow_fn, partial_order_fn, depth, layer_dict) or four Nones if no causal flow.

    depth is the largest flow *target label* (a reference quirk, flow.py:147)."""
    if len(input_nodes) != len(output_nodes):
        raise ValueError(
            f"Cannot find flow or gflow. Input ({len(input_nodes)}) and output "
            f"({len(output_nodes)}) nodes have different size."
        )
    nodes = list(graph.nodes())
    node_set = set(nodes)
    layer: Dict[int, int] = {v: 0 for v in nodes}
    pending = {v: 0 for v in nodes}  # number of not-yet-corrected neighbours
    succ: Dict[int, int] = {}
    inputs, done = set(input_nodes), set(output_nodes)

    frontier = set()
    not_out = set(node_set - set(output_nodes))
    for v in set(output_nodes) - set(input_nodes):
        pending[v] = len(set(graph.neighbors(v)) & not_out)
        if pending[v] == 1:
            frontier = frontier.union({v})

    level = 1
    while True:
        nxt = set()
        for v in frontier:
            cand = set(graph.neighbors(v)) & (node_set - done)
            if len(cand) != 1:
                continue
            u = cand.pop()
            succ[u] = v
            layer[u] = level
            done.add(u)
            if u not in inputs:
                pending[u] = len(set(graph.neighbors(u)) & (node_set - done))
                if pending[u] == 1:
                    nxt.add(u)
            for w in set(graph.neighbors(u)):
                if pending[w] > 0:
                    pending[w] -= 1
                    if pending[w] == 1:
                        nxt.add(w)
        if not nxt:
            break
        frontier = nxt
        level += 1

    if len(succ) != len(nodes) - len(output_nodes):
        return None, None, None, None
    return (lambda x: succ[x]), (lambda a, b: layer[a] > layer[b]), max(succ.values()), layer


class Flow:
    """Flow object attached to a circuit as `gflow` (flow.py:58-109)."""

    def __init__(self, graph, input_nodes, output_nodes):
        self.graph = graph
        self.input_nodes = input_nodes
        self.output_nodes = output_nodes
        self.func, self.partial_order, self.depth, self.layers_dict = find_cflow(
            graph, input_nodes, output_nodes
        )
        self.layers: Optional[List[List[int]]] = None
        self.measurement_order: Optional[List[int]] = None
        if self.layers_dict is not None:
            top = max(self.layers_dict.values())
            by_level = [[v for v, lv in self.layers_dict.items() if lv == j] for j in range(top + 1)]
            self.layers = by_level[::-1]
            order = [v for group in self.layers for v in group]
            for v in reversed(list(input_nodes)):
                order.remove(v)
                order.insert(0, v)
            self.measurement_order = order

    def __call__(self, node):
        return self.func(node)

    def __repr__(self):
        return f"Flow(n={self.graph.number_of_nodes()})"

    def _correction_sources(self):
        """node -> (X sources, Z sources) under the rule of pennylane_simulator.py:145-153."""
        if self.func is None:
            raise ValueError("the graph has no causal flow")
        order = self.measurement_order
        pos = {v: i for i, v in enumerate(order)}
        xs = {v: [] for v in order}
        zs = {v: [] for v in order}
        for node in order:
            if node in self.output_nodes:
                continue
            tgt = self.func(node)
            xs[tgt].append(node)
            for nb in self.graph.neighbors(tgt):
                if nb != node and pos[nb] > pos[node]:
                    zs[nb].append(node)
        return xs, zs

    def adapt_angle(self, angle, node, previous_outcomes):
        """XY-plane angle of `node` given the outcomes (dict node -> 0/1) measured so far:
        (-1)^a angle + b pi, a / b = parity of the outcomes that put an X / Z on `node`.
        (A stub in the reference, flow.py:105-109; the rule is pennylane_simulator.py:145-153.)"""
        xs, zs = self._correction_sources()
        a = sum(int(previous_outcomes.get(i, 0)) for i in xs[node]) % 2
        b = sum(int(previous_outcomes.get(i, 0)) for i in zs[node]) % 2
        return (-1) ** a * angle + b * 3.141592653589793

    def adapt_angles(self, angles, outcomes):
        """Adapted angles for all measured nodes in measurement order (angles: same order)."""
        measured = [v for v in self.measurement_order if v not in self.output_nodes]
        return [self.adapt_angle(a, v, outcomes) for a, v in zip(angles, measured)]


def check_if_flow(graph, input_nodes, output_nodes, flow: Callable, partial_order: Callable) -> bool:
    """True iff (flow, partial_order) is a causal flow of the open graph; reports which of the
    three flow conditions fails for which node (flow.py:603-623)."""
    outs = set(output_nodes)
    ok = True
    for v in [n for n in graph.nodes() if n not in outs]:
        fv = flow(v)
        around = list(graph.neighbors(fv))
        problems = []
        if v not in around:
            problems.append(f"Condition 1 failed for node {v}. {v} not in {around}")
        if not partial_order(v, fv):
            problems.append(f"Condition 2 failed for node {v}. {v} ≮ {fv}")
        late = [w for w in set(around) - {v} if not partial_order(v, w)]
        if late:
            problems.append(f"Condition 3 failed for node {v}: " + ", ".join(f"{v} ≮ {w}" for w in late))
        for msg in problems:
            print(msg)
        ok = ok and not problems
    return bool(ok)
