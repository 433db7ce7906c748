"""MBQC pattern container: open graph + measurement table + flow-derived measurement order.

Host-side mirror of mentpy/mbqc/mbqcircuit.py:22-448 (MBQCircuit) and :459-611
(merge / vstack / hstack).  Nothing here is accelerated -- it runs once per pattern -- but every
list it produces feeds the CUDA plan (mentpy_b200/plan.py), so the orderings are kept identical
to the reference and pinned by tests/test_host_indexing.py against tables dumped from it:

  * nodes are relabelled to 0..N-1 by sorted label, node *order* is kept      (:71-87)
  * default measurement table: non-outputs in node order, then outputs=None  (:110-113)
  * trainable_nodes follow the measurement-table order, not measurement order (:325-349)
  * measurement order = flow layers descending, ties in node order, inputs first (:390-422)
"""
import copy
from functools import reduce
from typing import Callable, Dict, List, Optional

import numpy as np

from .causal_flow import Flow, check_if_flow, find_cflow
from .graph import GraphState, contract_into, disjoint_union
from .measurement import ControlMent, Ment

__all__ = ["MBQCircuit", "merge", "hstack", "vstack"]


class MBQCircuit:
    def __init__(
        self,
        graph,
        input_nodes: List[int] = [],
        output_nodes: List[int] = [],
        measurements: Optional[Dict[int, Optional[Ment]]] = None,
        default_measurement: Optional[Ment] = Ment("XY"),
        flow: Optional[Callable] = None,
        partial_order: Optional[Callable] = None,
        measurement_order: Optional[List[int]] = None,
        relabel_indices: bool = True,
    ) -> None:
        graph = GraphState.from_any(graph)
        if relabel_indices:
            ranked = sorted(graph.nodes())
            fwd = {v: i for i, v in enumerate(ranked)}
            back = {i: v for i, v in enumerate(ranked)}
            graph = graph.relabeled(fwd)
            input_nodes = [fwd[v] for v in input_nodes]
            output_nodes = [fwd[v] for v in output_nodes]
            if flow is not None:
                user_flow = flow
                flow = lambda x: fwd[user_flow(back[x])]  # noqa: E731
            if partial_order is not None:
                user_po = partial_order
                partial_order = lambda a, b: user_po(back[a], back[b])  # noqa: E731
            if measurement_order is not None:
                measurement_order = [fwd[v] for v in measurement_order]
            if measurements is not None:
                measurements = {fwd[k]: m for k, m in measurements.items()}

        self._graph = graph
        present = set(graph.nodes())
        if any(v not in present for v in input_nodes):
            raise ValueError(f"Input nodes {input_nodes} are not in the graph. Graph nodes are {graph.nodes()}")
        if any(v not in present for v in output_nodes):
            raise ValueError(f"Output nodes {output_nodes} are not in the graph. Graph nodes are {graph.nodes()}")
        self._input_nodes = list(input_nodes)
        self._output_nodes = list(output_nodes)
        if not isinstance(default_measurement, Ment):
            raise ValueError(f"Default measurement {default_measurement} is not an instance of Ment.")
        self._default_measurement = default_measurement
        self._outputc = [v for v in graph.nodes() if v not in self._output_nodes]
        self._inputc = [v for v in graph.nodes() if v not in self._input_nodes]

        if measurements is None:
            table = {v: default_measurement for v in self._outputc}
            for v in self._output_nodes:
                table[v] = None
        else:
            stray = [v for v in measurements if v not in present]
            if stray:
                raise ValueError(f"Nodes {stray} are not in the graph.")
            if any(not (m is None or isinstance(m, Ment)) for m in measurements.values()):
                raise ValueError(f"Values {measurements.values()} are not instances of Ment.")
            table = dict(measurements)
            for v in graph.nodes():
                if v not in table:
                    table[v] = default_measurement if v in self._outputc else None
        self._measurements = table
        self._flow = self._partial_order = None
        self._refresh()

        if flow is None or partial_order is None:
            flow, partial_order, _depth, _layers = find_cflow(graph, self._input_nodes, self._output_nodes)
            self.gflow = Flow(graph, self._input_nodes, self._output_nodes)
        else:
            check_if_flow(graph, self._input_nodes, self._output_nodes, flow, partial_order)
            if not hasattr(self, "gflow"):
                self.gflow = Flow(graph, self._input_nodes, self._output_nodes)
        self._flow = flow
        self._partial_order = partial_order
        if measurement_order is None and flow is not None:
            measurement_order = self.calculate_order()
        self._quantum_output_nodes = [v for v, m in self._measurements.items() if m is None]
        self._measurement_order = measurement_order

    # -- derived tables -------------------------------------------------------------------------
    def _refresh(self) -> None:
        """Rebuild trainable / plane / output tables from the measurement table (:325-373)."""
        trainable, controlled, planes, q_out, c_out = [], [], {}, [], []
        for v, m in self._measurements.items():
            if m is None:
                planes[v] = ""
                if v in self._output_nodes:
                    q_out.append(v)
                continue
            if isinstance(m, ControlMent):
                controlled.append(v)
            if m.is_trainable():
                trainable.append(v)
            planes[v] = m.plane() if isinstance(m, ControlMent) else m.plane  # ControlMent: plane of the false branch
            owned = copy.deepcopy(m)
            owned.node_id = v
            self._measurements[v] = owned
            if v in self._output_nodes:
                c_out.append(v)
        self._trainable_nodes = trainable
        self._controlled_nodes = controlled
        self._planes = planes
        self._quantum_output_nodes = q_out
        self._classical_output_nodes = c_out
        if controlled and getattr(self, "_partial_order", None) is not None:
            self._partial_order = _order_with_conditions(controlled, self._measurements, self._partial_order)

    def calculate_order(self) -> List[int]:
        """Layers descending (count of strictly-later nodes), ties in node order, inputs first."""
        nodes = self._graph.nodes()
        later = [sum(1 for b in nodes if self._partial_order(a, b)) for a in nodes]
        groups: Dict[int, List[int]] = {}
        for v, cnt in zip(nodes, later):
            groups.setdefault(cnt, []).append(v)
        self._sorted_labels = [groups[c] for c in sorted(groups, reverse=True)]
        order = [v for grp in self._sorted_labels for v in grp]
        for v in reversed(self._input_nodes):
            order.remove(v)
            order.insert(0, v)
        return order

    # -- container protocol ---------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._graph)

    def __repr__(self) -> str:
        return f"MBQCircuit with {self._graph.number_of_nodes()} qubits."

    def __getattr__(self, name):
        # only reached when normal lookup fails: fall through to the graph, then the flow object
        if name.startswith("__") or name in ("_graph", "gflow"):
            raise AttributeError(name)
        try:
            return getattr(self._graph, name)
        except AttributeError:
            try:
                return getattr(self.gflow, name)
            except AttributeError:
                raise AttributeError(f"Attribute {name} not found in MBQCircuit.")

    def __getitem__(self, node):
        try:
            return self._measurements[node]
        except KeyError:
            raise ValueError(f"Node {node} is not in the graph.")

    def __setitem__(self, node, ment) -> None:
        if node not in self._graph:
            raise ValueError(f"Node {node} is not in the graph.")
        if not isinstance(ment, Ment):
            raise ValueError(f"Value {ment} is not a Measurement object.")
        self._measurements[node] = ment
        self._refresh()
        if isinstance(ment, ControlMent):
            # the controlled node must come after the nodes its condition reads (mbqcircuit.py:191-193)
            self._measurement_order = self.calculate_order()

    def __delitem__(self, node) -> None:
        if node not in self._graph:
            raise ValueError(f"Node {node} is not in the graph.")
        self._measurements[node] = None

    # -- properties (same names as the reference) -----------------------------------------------
    graph = property(lambda self: self._graph)
    input_nodes = property(lambda self: self._input_nodes)
    output_nodes = property(lambda self: self._output_nodes)
    quantum_output_nodes = property(lambda self: self._quantum_output_nodes)
    classical_output_nodes = property(lambda self: self._classical_output_nodes)
    controlled_nodes = property(lambda self: self._controlled_nodes)
    planes = property(lambda self: self._planes)
    flow = property(lambda self: self._flow)
    partial_order = property(lambda self: self._partial_order)
    outputc = property(lambda self: self._outputc)
    inputc = property(lambda self: self._inputc)

    @property
    def depth(self) -> int:
        return self.gflow.depth

    @property
    def measurements(self) -> Dict[int, Optional[Ment]]:
        return self._measurements

    @measurements.setter
    def measurements(self, table: Dict[int, Ment]) -> None:
        if any(v not in self._graph for v in table):
            raise ValueError(f"Nodes {table.keys()} are not in the graph.")
        if any(not isinstance(m, Ment) for m in table.values()):
            raise ValueError(f"Values {table.values()} are not Measurement objects.")
        self._measurements = table
        self._refresh()

    @property
    def trainable_nodes(self) -> List[int]:
        return self._trainable_nodes

    @trainable_nodes.setter
    def trainable_nodes(self, nodes: List[int]) -> None:
        if any(v not in self._graph for v in nodes):
            raise ValueError(f"Trainable nodes {nodes} are not in the graph. Graph nodes are {self._graph.nodes()}")
        self._trainable_nodes = nodes

    @property
    def measurement_order(self) -> List[int]:
        return self._measurement_order

    @measurement_order.setter
    def measurement_order(self, order: List[int]) -> None:
        po = self._partial_order
        for i in range(len(order)):
            for j in range(i + 1, len(order)):
                if po(order[i], order[j]):
                    raise ValueError(f"Invalid measurement order {order}.")
        self._measurement_order = order

    def ordered_layers(self, train_indices: bool = False):
        if self.gflow.func is None:
            return None
        if train_indices:
            return [[self._trainable_nodes.index(v) for v in layer] for layer in self.gflow.layers[:-1]]
        return self.gflow.layers

    # -- graph edits re-derive everything with default measurements (:424-448) ------------------
    def add_edge(self, u, v) -> None:
        self._graph.add_edge(u, v)
        try:
            self.__init__(self._graph, self._input_nodes, self._output_nodes)
        except Exception as exc:
            self._graph.remove_edge(u, v)
            raise ValueError(f"Cannot add edge between {u} and {v}.\n" + str(exc))

    def add_edges_from(self, edges, **kwargs) -> None:
        grown = self._graph.copy()
        grown.add_edges_from(edges)
        try:
            self.__init__(grown, self._input_nodes, self._output_nodes)
        except Exception as exc:
            raise ValueError(f"Cannot add edges {edges}.\n" + str(exc))


# ---------------------------------------------------------------------------------------------
# composition (mbqcircuit.py:459-611)
# ---------------------------------------------------------------------------------------------
def _shifted_table(state: MBQCircuit, key) -> Dict[int, Optional[Ment]]:
    return {key(v): m for v, m in state.measurements.items()}


def _vstack2(a: MBQCircuit, b: MBQCircuit) -> MBQCircuit:
    off = len(a.graph)
    table = dict(a.measurements)
    table.update(_shifted_table(b, lambda v: v + off))
    return MBQCircuit(
        disjoint_union(a.graph, b.graph),
        a.input_nodes + [v + off for v in b.input_nodes],
        a.output_nodes + [v + off for v in b.output_nodes],
        measurements=table,
    )


def _hstack2(a: MBQCircuit, b: MBQCircuit) -> MBQCircuit:
    if len(a.output_nodes) != len(b.input_nodes):
        raise ValueError(
            "The output of the first state must be the same size as the input of the second state."
        )
    g = disjoint_union(a.graph, b.graph)
    pos_a = {v: i for i, v in enumerate(a.graph.nodes())}
    pos_b = {v: i for i, v in enumerate(b.graph.nodes())}
    off = len(pos_a)
    for out_a, in_b in zip(a.output_nodes, b.input_nodes):
        g.add_edge(pos_a[out_a], pos_b[in_b] + off)
    inputs = [pos_a[v] for v in a.input_nodes]
    outputs = [pos_b[v] + off for v in b.output_nodes]
    table = _shifted_table(a, lambda v: pos_a[v])
    table.update(_shifted_table(b, lambda v: pos_b[v] + off))
    for out_a, in_b in zip(a.output_nodes, b.input_nodes):
        # the reference indexes by raw label here (mbqcircuit.py:604-607); circuits are always
        # relabelled to 0..N-1 in node order, where label == position
        g.add_edge(out_a, in_b + off)
        g = contract_into(g, keep=in_b + off, gone=out_a)
        del table[out_a]
    return MBQCircuit(g, inputs, outputs, measurements=table)


def vstack(states) -> MBQCircuit:
    """Side by side: inputs/outputs of all blocks are kept."""
    if len(states) == 0:
        raise ValueError("Cannot vertically stack an empty list of states.")
    return reduce(_vstack2, states)


def hstack(states) -> MBQCircuit:
    """In sequence: outputs of each block are identified with the inputs of the next."""
    if len(states) == 0:
        raise ValueError("Cannot horizontally stack an empty list of states.")
    return reduce(_hstack2, states)


def merge(a: MBQCircuit, b: MBQCircuit, along=[]) -> MBQCircuit:
    """Identify output i of `a` with input j of `b` for every (i, j) in `along`."""
    for i, j in along:
        if i not in a.output_nodes or j not in b.input_nodes:
            raise ValueError(f"Cannot merge states at indices {i} and {j}")
    off = len(a.graph)
    g = disjoint_union(a.graph, b.graph)
    outs_a, ins_b = zip(*along)
    inputs = a.input_nodes + [v + off for v in b.input_nodes if v not in ins_b]
    outputs, used = [], []
    for v in a.output_nodes:
        if v in outs_a:
            slot = b.input_nodes.index(ins_b[outs_a.index(v)])
            used.append(slot)
            outputs.append(b.output_nodes[slot] + off)
        else:
            outputs.append(v)
    outputs += [v + off for s, v in enumerate(b.output_nodes) if s not in used]
    table = dict(a.measurements)
    table.update(_shifted_table(b, lambda v: v + off))
    for i, j in along:
        g.add_edge(i, j + off)
        g = contract_into(g, keep=j + off, gone=i)
        del table[i]
    return MBQCircuit(g, inputs, outputs, measurements=table)


def _order_with_conditions(controlled, measurements, base):
    """Partial order extended by the outcome dependencies of the controlled measurements
    (mbqcircuit.py:614-633): whatever precedes (or is) a condition node of c precedes everything from c on."""
    def order(i, j):
        for c in controlled:
            cond = measurements[c].condition
            reads = cond.cond_nodes  # a plain bool condition has none: AttributeError, as in the reference
            i_feeds_c = i in reads or any(base(i, r) for r in reads)
            j_from_c = j == c or base(c, j)
            if i_feeds_c and j_from_c and i != j:
                return True
            if j in reads and i == c:
                return False
        return base(i, j)

    return order
