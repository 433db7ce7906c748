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
    # outcome-controlled measurements (operators/controlled_ment.py:14-113): node -> {"reads": [nodes],
    # "table": [0/1 per assignment of the read outcomes, first read node = lowest index bit],
    # "true": (plane, fixed|None), "false": (plane, fixed|None)}; `measurements[node]` holds the false branch
    controls: Dict[int, dict] = field(default_factory=dict)

    def control_branch(self, node: int, outcomes: Dict[int, int]):
        """(plane, fixed angle|None) of a controlled node for the outcome record so far."""
        c = self.controls[node]
        idx = sum((int(outcomes[r]) & 1) << i for i, r in enumerate(c["reads"]))
        return tuple(c["true"]) if c["table"][idx] else tuple(c["false"])

    def neighbors(self, v: int) -> List[int]:
        out = []
        for a, b in self.edges:
            if a == v:
                out.append(b)
            elif b == v:
                out.append(a)
        return out

    def to_json(self) -> dict:
        d = {
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
        if self.controls:  # only patterns with controlled measurements carry the key
            d["controls"] = {str(k): v for k, v in self.controls.items()}
        return d

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
            controls={int(k): {"reads": [int(r) for r in v["reads"]], "table": [int(t) for t in v["table"]],
                               "true": (v["true"][0], _ang(v["true"][1])), "false": (v["false"][0], _ang(v["false"][1]))}
                      for k, v in d.get("controls", {}).items()},
        )

    @staticmethod
    def from_circuit(circ) -> "PatternData":
        """Fill from an MBQCircuit-like object (the reference's or the product's host mirror)."""
        meas, controls = {}, {}
        for node, m in circ.measurements.items():
            if m is None:
                meas[int(node)] = None
            elif hasattr(m, "_true_ment"):  # ControlMent (either side): tabulate the condition
                cond = m.condition
                reads = [int(r) for r in _condition_reads(cond)]
                table = [int(bool(cond({r: (idx >> i) & 1 for i, r in enumerate(reads)}))) for idx in range(1 << len(reads))]
                tm = m._true_ment
                controls[int(node)] = {"reads": reads, "table": table, "true": (str(tm.plane), _ang(tm.angle)),
                                       "false": (str(m._plane), _ang(m._angle))}
                meas[int(node)] = (str(m._plane), _ang(m._angle))
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
            controls=controls,
        )


def _ang(a):
    if a is None:
        return None
    return tuple(float(x) for x in a) if isinstance(a, (tuple, list)) else float(a)


class _Probe(dict):
    def __init__(self, values, seen):
        super().__init__(values)
        self._seen = seen

    def __getitem__(self, key):
        self._seen.add(key)
        return dict.get(self, key, 0)


def _condition_reads(cond):
    """Nodes a condition reads, by probing (the reference's cond_nodes loses nodes under `~`, ment.py:118-119)."""
    seen = set(getattr(cond, "cond_nodes", ()) or ())
    while True:
        before = set(seen)
        keys = sorted(before)
        for idx in range(1 << len(keys)):
            cond(_Probe({k: (idx >> i) & 1 for i, k in enumerate(keys)}, seen))
        if seen == before:
            return sorted(seen)
