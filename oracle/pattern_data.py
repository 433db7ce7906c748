"""TEST INFRASTRUCTURE ONLY -- plain-data description of an MBQC pattern for the oracles.

The oracles never import the product package and never import the reference: they work on this
neutral record, which can be filled from either side (both expose the attribute names of
mentpy/mbqc/mbqcircuit.py:226-300: graph, input_nodes, output_nodes, measurements,
trainable_nodes, measurement_order, quantum_output_nodes) or from a golden JSON fixture.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass
class PatternData:
    n_nodes: int
    edges: List[Tuple[int, int]]
    input_nodes: List[int]
    output_nodes: List[int]
    # node -> (plane, fixed_angle|None) for measured nodes, None for unmeasured (quantum output)
    measurements: Dict[int, Optional[Tuple[str, Optional[float]]]]
    trainable_nodes: List[int]
    measurement_order: List[int]
    quantum_output_nodes: List[int] = field(default_factory=list)

    def neighbors(self, v: int) -> List[int]:
        out = []
        for a, b in self.edges:
            if a == v:
                out.append(b)
            elif b == v:
                out.append(a)
        return out

    def to_json(self) -> dict:
        return {
            "n_nodes": self.n_nodes,
            "edges": [list(e) for e in self.edges],
            "input_nodes": list(self.input_nodes),
            "output_nodes": list(self.output_nodes),
            "measurements": {
                str(k): (None if v is None else [v[0], v[1]]) for k, v in self.measurements.items()
            },
            "trainable_nodes": list(self.trainable_nodes),
            "measurement_order": None if self.measurement_order is None else list(self.measurement_order),
            "quantum_output_nodes": list(self.quantum_output_nodes),
        }

    @staticmethod
    def from_json(d: dict) -> "PatternData":
        meas = {}
        for k, v in d["measurements"].items():
            # JSON turns the two-angle tuple of an XYZ node into a list: restore it
            meas[int(k)] = None if v is None else (v[0], tuple(v[1]) if isinstance(v[1], list) else v[1])
        return PatternData(
            n_nodes=int(d["n_nodes"]),
            edges=[(int(a), int(b)) for a, b in d["edges"]],
            input_nodes=[int(x) for x in d["input_nodes"]],
            output_nodes=[int(x) for x in d["output_nodes"]],
            measurements=meas,
            trainable_nodes=[int(x) for x in d["trainable_nodes"]],
            measurement_order=None if d["measurement_order"] is None else [int(x) for x in d["measurement_order"]],
            quantum_output_nodes=[int(x) for x in d.get("quantum_output_nodes", [])],
        )

    @staticmethod
    def from_circuit(circ) -> "PatternData":
        """Fill from an MBQCircuit-like object (the reference's or the product's host mirror)."""
        meas = {}
        for node, m in circ.measurements.items():
            if m is None:
                meas[int(node)] = None
            else:
                ang = m.angle
                # XYZ carries two angles (ment.py:239-251); every other plane one
                meas[int(node)] = (str(m.plane), None if ang is None else
                                   (tuple(float(x) for x in ang) if isinstance(ang, (tuple, list)) else float(ang)))
        qout = getattr(circ, "quantum_output_nodes", None)
        if qout is None:
            qout = [n for n in circ.output_nodes if meas.get(n) is None]
        order = circ.measurement_order
        return PatternData(
            n_nodes=len(list(circ.graph.nodes())),
            edges=[(int(a), int(b)) for a, b in circ.graph.edges()],
            input_nodes=[int(x) for x in circ.input_nodes],
            output_nodes=[int(x) for x in circ.output_nodes],
            measurements=meas,
            trainable_nodes=[int(x) for x in circ.trainable_nodes],
            measurement_order=None if order is None else [int(x) for x in order],
            quantum_output_nodes=[int(x) for x in qout],
        )
