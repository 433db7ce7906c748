"""Pattern generators used by the benchmark configs and the parity tests.

Host-side mirror of mentpy/mbqc/templates.py:16-295 (linear_cluster, many_wires, grid_cluster,
muta, spturb).  Node numbering is row-major (wire after wire), inputs are the first node of every wire and
outputs the last -- which is what fixes `measurement_order` and `trainable_nodes` for
BASELINE.json's configs.  from_pauli needs the GF(2) Pauli
algebra (galois) and is out of scope.
"""
from typing import List

from .circuit import MBQCircuit, hstack, merge
from .graph import GraphState
from .measurement import Ment

TRIANGLE_WIRE = 5  # nodes per wire in one MuTA block


def _wire_offsets(lengths: List[int]) -> List[int]:
    offs, total = [], 0
    for n in lengths:
        offs.append(total)
        total += n
    return offs


def _wires(lengths: List[int]) -> GraphState:
    g = GraphState()
    for start, n in zip(_wire_offsets(lengths), lengths):
        g.add_edges_from((start + j, start + j + 1) for j in range(n - 1))
    return g


def linear_cluster(n: int, **kwargs) -> MBQCircuit:
    """1D cluster of n qubits, input node 0, output node n-1."""
    return MBQCircuit(_wires([n]), input_nodes=[0], output_nodes=[n - 1], **kwargs)


def many_wires(n_wires: List[int], **kwargs) -> MBQCircuit:
    """Disconnected wires of the given lengths (each > 1)."""
    if not all(isinstance(n, int) and n > 1 for n in n_wires):
        raise ValueError("n_wires must be a list of integers greater than 1")
    offs = _wire_offsets(n_wires)
    return MBQCircuit(
        _wires(n_wires),
        input_nodes=list(offs),
        output_nodes=[o + n - 1 for o, n in zip(offs, n_wires)],
        **kwargs,
    )


def grid_cluster(n: int, m: int, periodic: bool = False, **kwargs) -> MBQCircuit:
    """n rows x m columns 2D cluster; `periodic` closes the rows into a cylinder."""
    g = _wires([m] * n)
    for r in range(n - 1):
        g.add_edges_from((r * m + c, (r + 1) * m + c) for c in range(m))
    if periodic and n > 1:
        for c in range(m):
            g.add_edge(c, (n - 1) * m + c)
    return MBQCircuit(
        g,
        input_nodes=[r * m for r in range(n)],
        output_nodes=[r * m + m - 1 for r in range(n)],
        **kwargs,
    )


def muta(n_wires: int, n_layers: int, **kwargs) -> MBQCircuit:
    """Multiple Triangle Ansatz: per block, wire `t` carries a node tied to two nodes of every
    other wire; blocks for t = 0..n_wires-1 are chained, and the chain is repeated n_layers times.
    `one_column=True` keeps only the t = 0 block."""
    opts = {"restrict_trainable": True, "one_column": False}
    opts.update(kwargs)
    column = None
    for t in range(n_wires):
        if opts["one_column"] and t != 0:
            break
        block = many_wires([TRIANGLE_WIRE] * n_wires)
        if opts["restrict_trainable"]:
            # kept for API fidelity: every later edit re-derives the table, so this has no
            # lasting effect (mentpy/mbqc/templates.py:186-191 and mbqcircuit.py:424-432)
            block.trainable_nodes = list(
                set(block.trainable_nodes) - set(v - 1 for v in block.output_nodes)
            )
        apex = TRIANGLE_WIRE * t + 1
        for other in range(n_wires):
            if other != t:
                block.add_edge(apex, TRIANGLE_WIRE * other)
                block.add_edge(apex, TRIANGLE_WIRE * other + 2)
        column = block if column is None else hstack((column, block))
    stacked = column
    for _ in range(1, n_layers):
        stacked = hstack((stacked, column))
    return stacked


def spturb(n_qubits: int, n_layers: int, periodic: bool = False, **kwargs) -> MBQCircuit:
    """Symmetry-protected-topological perturbator ansatz (mentpy/mbqc/templates.py:219-295).

    Every layer is a column of 3-node wires whose middle node is trainable (the rest measured in
    X), followed by two sweeps of two-wire "symmetry blocks" glued onto the outputs of wires i and
    i + 2: a 5-2-5 wire triple whose middle wire is tied to node 2 of the outer wires and carries
    the block's one trainable node; the second sweep measures the odd nodes of the outer wires in Y.
    n_qubits * n_layers + 2 * n_layers * (n_qubits if periodic else n_qubits - 2) trainable nodes."""
    if n_qubits < 4:
        raise ValueError("n_qubits must be greater than 4")
    frame = many_wires([5, 2, 5]).graph
    frame.add_edge(2, 6)
    frame.add_edge(9, 6)
    y_nodes = {v: Ment(plane="Y") for v in (1, 3, 8, 10)}
    blocks = [
        MBQCircuit(frame, input_nodes=[0, 7], output_nodes=[4, 11], measurements={5: Ment(plane="XY")},
                   default_measurement=Ment(plane="X")),
        MBQCircuit(frame, input_nodes=[0, 7], output_nodes=[4, 11], measurements={5: Ment(plane="XY"), **y_nodes},
                   default_measurement=Ment(plane="X")),
    ]

    def column():
        return many_wires([3] * n_qubits, measurements={3 * i + 1: Ment() for i in range(n_qubits)},
                          default_measurement=Ment(plane="X"))

    n_blocks = n_qubits if periodic else n_qubits - 2
    ansatz = None
    for _layer in range(n_layers):
        ansatz = column() if ansatz is None else hstack((ansatz, column()))
        for block in blocks:
            for i in range(n_blocks):
                first, second = ansatz.output_nodes[i], ansatz.output_nodes[(i + 2) % n_qubits]
                ansatz = merge(ansatz, block, [(first, 0), (second, 7)])
    return ansatz
