"""Lower an MBQCircuit + window + schedule to the flat plan the CUDA kernels execute.

This is the host half of what NumpySimulatorSV/DM.__init__ do per simulator
(mentpy/simulators/np_simulator_sv.py:38-128, np_simulator_dm.py:33-115): choose the schedule,
promote the default window, validate sizes (same exceptions), find which qubits sit in the first
window and which CZs act inside it -- plus the bookkeeping the reference redoes on every
measurement (np_simulator_sv.py:130-142, :207-223: who is in the window, who is the new qubit,
which of its neighbours are present), resolved once here into per-step records.

Slot recycling: qubits are never shifted.  The first window puts schedule[p] at slot w-1-p (the
reference's big-endian layout); afterwards the qubit appended after measurement m takes the slot
the measured qubit just freed.  The output permutation (np_simulator_sv.py:286-290,
np_simulator_dm.py:267-273) becomes a slot list.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from .mbqc.measurement import Ment, condition_reads

_PLANE_CODE = {"XY": _lib.PLANE_XY, "X": _lib.PLANE_XY, "Y": _lib.PLANE_XY,
               "XZ": _lib.PLANE_XZ, "YZ": _lib.PLANE_YZ, "Z": _lib.PLANE_Z, "XYZ": _lib.PLANE_XYZ}


@dataclass
class StepRecord:
    node: int
    slot: int
    angle_idx: int          # index into trainable_nodes, or -1
    plane: int
    fixed_angle: Optional[float]
    fixed_cos: float
    fixed_sin: float
    append: bool
    new_node: Optional[int]
    nbr_mask: int
    dropped_neighbours: List[int] = field(default_factory=list)
    fixed_z: float = 0.0    # plane XYZ: Z component of the measurement axis
    # outcome-controlled measurement (ControlMent): the fields above describe the FALSE branch, alt_* the
    # TRUE branch; cond_mask selects earlier outcomes (bit j = j+1 measurements back), cond_table is the
    # condition's truth table over them (lowest selected bit = lowest index bit); column = the node's angle column
    cond_mask: int = 0
    cond_table: int = 0
    alt_plane: int = 0
    alt_angle_idx: int = -1
    alt_cos: float = 1.0
    alt_sin: float = 0.0
    alt_z: float = 0.0
    column: int = -1


@dataclass
class LoweredPlan:
    window: int
    n_nodes: int
    schedule: List[int]
    schedule_measure: List[int]
    steps: List[StepRecord]
    input_nodes: List[int]
    input_slot: List[int]
    init_cz_mask: List[int]
    output_nodes: List[int]          # order of the output index, first = MSB
    output_slot: List[int]
    n_angles: int
    mixed: bool
    first_window: List[int]
    first_slots: Dict[int, int] = field(default_factory=dict)
    dev_windows: Optional[List[List[int]]] = None  # dev_mode: window content after every measurement

    def window_nodes_after(self, n_done: int) -> List[int]:
        """Reference window content (position 0 first) after n_done measurements."""
        if self.dev_windows is not None:
            return list(self.dev_windows[n_done])
        return self.schedule[n_done: n_done + self.window]

    def slot_of_after(self, n_done: int) -> Dict[int, int]:
        slot = {v: s for v, s in self.first_slots.items()}
        for st in self.steps[:n_done]:
            del slot[st.node]
            if st.append:
                slot[st.new_node] = st.slot
        return slot


def _fixed_cos_sin(plane: str, angle):
    # exact values for the Pauli planes (the reference uses the Pauli matrices directly,
    # mentpy/operators/ment.py:233-235); np.cos/np.sin otherwise, as ment.py:230 does
    if plane == "X":
        return 1.0, 0.0
    if plane == "Y":
        return 0.0, 1.0
    return float(np.cos(angle)), float(np.sin(angle))


def _is_controlled(ment) -> bool:
    """ControlMent of this package or of the reference (a mentpy.MBQCircuit is accepted as it is)."""
    return hasattr(ment, "_true_ment") and hasattr(ment, "condition")


def _dev_mode_order(circuit, schedule, w, n_meas, wires):
    """Measurement sequence and window contents of the reference's dev_mode scheduling
    (np_simulator_sv.py:173-203, np_simulator_dm.py:160-201): measure the FIRST node of the window
    that has no later neighbour in its wire, or whose first later wire-neighbour is already in the
    window; drop it from the window and append the next node of the schedule."""
    if wires is None:
        raise TypeError("dev_mode needs the wires of the pattern (wires=[[...], ...])")
    n_nodes = len(schedule)

    def later_in_wire(node):
        for wire in wires:
            if node in wire:
                return [v for v in wire if circuit.graph.has_edge(node, v) and v > node]
        raise ValueError(f"Node {node} is in no wire.")

    win, seq, windows = list(schedule[:w]), [], [list(schedule[:w])]
    while len(seq) < n_meas:
        pick = None
        for node in win:
            fut = later_in_wire(node)
            if not fut or fut[0] in win:
                pick = node
                break
        if pick is None:
            raise ValueError("WTF")  # the reference's own message (np_simulator_sv.py:268)
        if circuit[pick] is None:
            raise ValueError(f"dev_mode scheduling reaches node {pick}, which has no measurement.")
        seq.append(pick)
        win.remove(pick)
        if len(seq) + w <= n_nodes:
            win.append(schedule[len(seq) + w - 1])
        windows.append(list(win))
    return seq, windows


def lower(circuit, window_size: int = 1, schedule: Optional[Sequence[int]] = None,
          mixed: bool = False, slot_order: str = "msb", shard_bits: int = 0,
          dev_mode: bool = False, wires=None) -> LoweredPlan:
    """slot_order: "msb" puts window position p at slot w-1-p (reference layout; the measured slots
    then cycle w-1, w-2, ..., 0 and the register kernel takes its unrolled path); "lsb" puts it at
    slot p (exercises the kernels' generic slot path; results are identical); "shard-last" (with
    shard_bits = g) is "msb" inside the w-g local slots and keeps the g window positions measured
    LAST in the shard slots, so the first w-g measurements of a sharded run are local AND hit high,
    coalesced slots first."""
    nodes = list(circuit.graph.nodes())
    n_nodes = len(nodes)
    outputs_excluded = circuit.quantum_output_nodes if mixed else circuit.output_nodes

    if not mixed:
        for v in nodes:
            m = circuit[v]
            if m is not None and (_is_controlled(m) or m.plane not in ("X", "Y", "XY")):
                raise ValueError(f"Node {v} has plane {m.plane}, but only XY plane is supported.")

    if schedule is not None:
        schedule = list(schedule)
    elif circuit.measurement_order is not None:
        schedule = list(circuit.measurement_order)
        if window_size == 1 and circuit.flow is not None:
            window_size = len(circuit.input_nodes) + 1
    else:
        raise ValueError(
            "Schedule must be provided for numpy simulator as the MBQCircuit does not have a flow."
        )
    schedule_measure = [v for v in schedule if v not in outputs_excluded]

    n_in = len(circuit.input_nodes)
    if n_in > window_size:
        raise ValueError(
            f"Input state has {n_in} qubits, but window size is set to {window_size}."
            " Input state must have at most as many qubits as the window size minus one."
        )
    if window_size > len(schedule_measure):
        raise ValueError(
            f"Window size is set to {window_size}, but schedule only has {len(schedule_measure)} measurements."
        )
    if sorted(schedule) != sorted(nodes):
        raise ValueError("The schedule must visit every node of the graph exactly once.")
    n_meas = len(schedule_measure)
    dev_windows = None
    if dev_mode:
        # the window decides the order: the nodes enter in schedule order, the measured one is chosen
        # by the wire rule (any window position), so unmeasured outputs may sit anywhere
        schedule_measure, dev_windows = _dev_mode_order(circuit, schedule, window_size, n_meas, wires)
    elif schedule[:n_meas] != schedule_measure:
        # the reference would measure window position 0 with another node's Ment here
        # (np_simulator_sv.py:169-172 vs :130-135); refuse instead of reproducing garbage
        raise ValueError("Unmeasured output nodes must come last in the schedule.")
    if set(schedule[:n_in]) != set(circuit.input_nodes):
        raise ValueError(
            f"Both lists must have the same elements, but source={circuit.input_nodes} "
            f"and target={schedule[:n_in]}"
        )
    if not mixed:
        for v in circuit.output_nodes:
            if circuit[v] is not None:
                raise NotImplementedError(
                    "Measured output nodes are not supported on the state-vector path "
                    "(the reference indexes them by node label, np_simulator_sv.py:279-284)."
                )
    if window_size > _lib.MAX_WINDOW:
        raise NotImplementedError(f"window_size {window_size} > {_lib.MAX_WINDOW}")

    w = window_size
    first_window = schedule[:w]
    if slot_order not in ("msb", "lsb", "shard-last"):
        raise ValueError("slot_order must be 'msb', 'lsb' or 'shard-last'")
    n_local = w - max(int(shard_bits), 0)

    def first_slot(p: int) -> int:
        if slot_order == "msb":
            return w - 1 - p
        if slot_order == "lsb":
            return p
        return n_local - 1 - p if p < n_local else p

    slot_of: Dict[int, int] = {v: first_slot(p) for p, v in enumerate(first_window)}
    init_cz = [0] * w
    for a, b in circuit.graph.edges():
        if a in slot_of and b in slot_of and a != b:
            lo, hi = sorted((slot_of[a], slot_of[b]))
            init_cz[lo] ^= 1 << hi
    input_slot = [slot_of[v] for v in circuit.input_nodes]
    first_slots = dict(slot_of)

    trainable = list(circuit.trainable_nodes)
    steps: List[StepRecord] = []

    def branch_spec(node, b):
        """(plane code, angle column | -1, fixed angle, cos, sin, z) of one plain measurement."""
        if b.plane not in _PLANE_CODE or b.plane == "Z":
            raise NotImplementedError(f"Node {node}: plane {b.plane} is not supported in a controlled measurement.")
        if b.plane == "XYZ":
            if not isinstance(b.angle, tuple):
                raise TypeError(f"Invalid argument type. Expected tuple but got {float if b.angle is None else type(b.angle)}")
            t1, t2 = b.angle
            return (_PLANE_CODE["XYZ"], -1, b.angle, float(np.cos(t1) * np.cos(t2)), float(np.sin(t1) * np.cos(t2)),
                    float(np.sin(t2)))
        if b.is_trainable():
            return _PLANE_CODE[b.plane], trainable.index(node), None, 1.0, 0.0, 0.0
        fc, fs = _fixed_cos_sin(b.plane, b.angle)
        return _PLANE_CODE[b.plane], -1, b.angle, fc, fs, 0.0

    for m, node in enumerate(schedule_measure):
        ment = circuit[node]
        ctl = None
        if _is_controlled(ment):
            # controlled_ment.py:96-113 through np_simulator_dm.py:307-346: only the density-matrix simulator
            # evaluates conditions, and only for nodes that receive an angle (a fixed-fixed ControlMent raises)
            if not mixed:
                raise ValueError(f"Node {node} has plane {ment.plane}, but only XY plane is supported.")
            if node not in trainable:
                raise ValueError("ControlledMent is not trainable, so angle must be None.")
            cond = ment.condition
            reads = condition_reads(cond)
            pos = {}
            for r in reads:
                if r not in schedule_measure[:m]:
                    raise ValueError(f"Node {node}: the condition reads node {r}, which is not measured before it.")
                pos[r] = m - 1 - schedule_measure.index(r)
            if len(reads) > 5 or any(d > 31 for d in pos.values()):
                raise NotImplementedError(f"Node {node}: a condition reads at most 5 outcomes, at most 32 measurements back.")
            by_bit = sorted(reads, key=lambda r: pos[r])  # lowest history bit first = lowest table-index bit
            table = 0
            for idx in range(1 << len(by_bit)):
                if cond({r: (idx >> i) & 1 for i, r in enumerate(by_bit)}):
                    table |= 1 << idx
            true_ment, false_ment = ment._true_ment, Ment(ment._angle, ment._plane)
            if not true_ment.is_trainable():
                # taking a fixed true branch makes the reference raise: ControlMent.get_povm hands the node's
                # angle to it (controlled_ment.py:109-111 -> ment.py:222-226); refused here for every sample
                shown = true_ment.angle
                raise ValueError(f"Measurement has a fixed angle of {round(shown, 4) if isinstance(shown, (int, float)) else shown}")
            f_spec, t_spec = branch_spec(node, false_ment), branch_spec(node, true_ment)
            if not by_bit:  # constant condition: a plain step
                ment = true_ment if table & 1 else false_ment
            else:
                ctl = (sum(1 << pos[r] for r in by_bit), table, f_spec, t_spec)
        if ctl is not None:
            mask_c, table, (pl, aidx, fixed, fc, fs, fz), (apl, aaidx, _af, afc, afs, afz) = ctl
            slot = slot_of.pop(node)
            done = m + 1
            append = done + w <= n_nodes
            new_node, mask, dropped = None, 0, []
            if append:
                new_node = schedule[done + w - 1]
                for nb in circuit.graph.neighbors(new_node):
                    if nb in slot_of:
                        mask |= 1 << slot_of[nb]
                    elif nb in schedule_measure[:done]:
                        dropped.append(nb)
                slot_of[new_node] = slot
            steps.append(StepRecord(node, slot, aidx, pl, fixed, fc, fs, append, new_node, mask, dropped, fz,
                                    cond_mask=mask_c, cond_table=table, alt_plane=apl, alt_angle_idx=aaidx,
                                    alt_cos=afc, alt_sin=afs, alt_z=afz, column=trainable.index(node)))
            continue
        plane = ment.plane
        if plane not in _PLANE_CODE:
            raise NotImplementedError(f"Node {node}: plane {plane} is not supported on the CUDA path.")
        if plane == "Z" and window_size > _lib.MAX_WINDOW_REG:
            raise NotImplementedError(f"plane-Z measurements cover window_size <= {_lib.MAX_WINDOW_REG}")
        fz = 0.0
        if plane == "Z":  # angle-free; only mode="expectation" runs it (np_simulator_dm.py:327-344)
            angle_idx, fixed, fc, fs = -1, None, 1.0, 0.0
        elif plane == "XYZ":
            # two angles, so only a fixed tuple works in the reference: a trainable XYZ node receives one
            # float from the angle vector and ment.py:240-245 raises
            if node in trainable or not isinstance(ment.angle, tuple):
                got = float if node in trainable else type(ment.angle)
                raise TypeError(f"Invalid argument type. Expected tuple but got {got}")
            t1, t2 = ment.angle
            angle_idx, fixed = -1, ment.angle
            fc, fs, fz = float(np.cos(t1) * np.cos(t2)), float(np.sin(t1) * np.cos(t2)), float(np.sin(t2))
        elif node in trainable:
            angle_idx, fixed, fc, fs = trainable.index(node), None, 1.0, 0.0
        else:
            if ment.angle is None:
                raise ValueError(f"Node {node} is not trainable but has no fixed angle.")
            angle_idx, fixed = -1, ment.angle
            fc, fs = _fixed_cos_sin(plane, ment.angle)
        slot = slot_of.pop(node)
        done = m + 1
        append = done + w <= n_nodes
        new_node, mask, dropped = None, 0, []
        if append:
            new_node = schedule[done + w - 1]
            for nb in circuit.graph.neighbors(new_node):
                if nb in slot_of:
                    mask |= 1 << slot_of[nb]
                elif nb in schedule_measure[:done]:
                    dropped.append(nb)  # already measured: the reference silently skips this CZ
            slot_of[new_node] = slot
        steps.append(StepRecord(node, slot, angle_idx, _PLANE_CODE[plane], fixed, fc, fs, append,
                                new_node, mask, dropped, fz))

    remaining = schedule[n_meas:] if not dev_mode else [v for v in schedule if v not in schedule_measure]
    # np_simulator_dm.py:267-273 reorders to quantum_output_nodes; np_simulator_sv.py:286-290 leaves
    # the window order alone when it already equals quantum_output_nodes and otherwise reorders to
    # output_nodes (the two lists differ in order for merged circuits) -- reproduced as is
    if mixed or list(circuit.quantum_output_nodes) == remaining:
        out_order = list(circuit.quantum_output_nodes)
    else:
        out_order = list(circuit.output_nodes)
    if set(out_order) != set(remaining):
        raise ValueError(f"Both lists must have the same elements, but source={remaining} and target={out_order}")
    output_slot = [slot_of[v] for v in out_order]
    if len(out_order) > _lib.MAX_IO or n_in > _lib.MAX_IO:
        raise NotImplementedError(f"more than {_lib.MAX_IO} input/output qubits")
    return LoweredPlan(w, n_nodes, schedule, schedule_measure, steps, list(circuit.input_nodes),
                       input_slot, init_cz, out_order, output_slot, len(trainable), mixed,
                       first_window, first_slots, dev_windows)


@dataclass
class FeedForward:
    """Flow corrections of one measurement step in the form the kernels consume: the outcome of
    this step, if 1, is remembered in a shift register (bit d of it = outcome of step m-1-d);
    xdep / zdep select the earlier outcomes that put a pending X / Z on THIS step's qubit
    (theta' = (-1)^a theta + b pi), outx / outz the output qubits (bit q = q-th output node) whose
    byproduct toggles when this step's outcome is 1."""
    xdep: int = 0
    zdep: int = 0
    outx: int = 0
    outz: int = 0


def correction_sources(circuit, schedule: Sequence[int]):
    """node -> (x_sources, z_sources): the measured nodes whose outcome 1 toggles a pending X / Z
    on `node`.  The rule of mentpy/simulators/pennylane_simulator.py:145-153: outcome 1 at node i
    applies X to f(i) and Z to every neighbour of f(i) other than i that is measured after i."""
    if circuit.flow is None:
        raise ValueError("outcome sampling needs a causal flow for the byproduct corrections")
    pos = {v: i for i, v in enumerate(schedule)}
    outs = set(circuit.quantum_output_nodes)
    xs: Dict[int, List[int]] = {v: [] for v in schedule}
    zs: Dict[int, List[int]] = {v: [] for v in schedule}
    for node in schedule:
        if node in outs:
            continue
        tgt = circuit.flow(node)
        xs[tgt].append(node)
        for nb in circuit.graph.neighbors(tgt):
            if nb != node and pos[nb] > pos[node]:
                zs[nb].append(node)
    return xs, zs


def feedforward(circuit, plan: LoweredPlan) -> List[FeedForward]:
    """Per-step correction masks for sampled runs (force0=False)."""
    xs, zs = correction_sources(circuit, plan.schedule)
    step_of = {st.node: m for m, st in enumerate(plan.steps)}
    ff = [FeedForward() for _ in plan.steps]
    for m, st in enumerate(plan.steps):
        if st.plane != _lib.PLANE_XY:
            raise NotImplementedError("byproduct corrections are implemented for XY-plane (and X, Y) measurements")
        for src, attr in ((xs[st.node], "xdep"), (zs[st.node], "zdep")):
            for i in src:
                d = m - 1 - step_of[i]
                if not 0 <= d < 32:
                    raise NotImplementedError(f"node {st.node} depends on the outcome of node {i}, {d + 1} steps back (max 32)")
                setattr(ff[m], attr, getattr(ff[m], attr) ^ (1 << d))
    for q, v in enumerate(plan.output_nodes):
        for i in xs[v]:
            ff[step_of[i]].outx ^= 1 << q
        for i in zs[v]:
            ff[step_of[i]].outz ^= 1 << q
    return ff


def window_is_valid(plan: LoweredPlan) -> bool:
    """True when no CZ was dropped because a neighbour had already left the window."""
    return all(not st.dropped_neighbours for st in plan.steps)


# ---------------------------------------------------------------------------------------------
# noise: Kraus set -> block-form coefficients (include/mbqc_b200.h: mbqc_noise)
# ---------------------------------------------------------------------------------------------
def kraus_set(kind: str, p: float = 0.0, gamma: Optional[float] = None, p_gad: float = 0.5):
    """Kraus operators of the single-qubit channels the reference can request from PennyLane
    (mentpy/simulators/pennylane_simulator.py:123-136; PennyLane's published definitions)."""
    eye = np.eye(2, dtype=complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    y = np.array([[0, -1j], [1j, 0]], dtype=complex)
    z = np.array([[1, 0], [0, -1]], dtype=complex)
    g = p if gamma is None else gamma
    if kind == "depolarizing":
        return [np.sqrt(1 - p) * eye, np.sqrt(p / 3) * x, np.sqrt(p / 3) * y, np.sqrt(p / 3) * z]
    if kind == "phase_flip":
        return [np.sqrt(1 - p) * eye, np.sqrt(p) * z]
    if kind == "bit_flip":
        return [np.sqrt(1 - p) * eye, np.sqrt(p) * x]
    damp0 = np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=complex)
    if kind == "amplitude_damping":
        return [damp0, np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex)]
    if kind == "phase_damping":
        return [damp0, np.array([[0, 0], [0, np.sqrt(g)]], dtype=complex)]
    if kind == "generalized_amplitude_damping":
        return [np.sqrt(p_gad) * damp0,
                np.sqrt(p_gad) * np.array([[0, np.sqrt(g)], [0, 0]], dtype=complex),
                np.sqrt(1 - p_gad) * np.array([[np.sqrt(1 - g), 0], [0, 1]], dtype=complex),
                np.sqrt(1 - p_gad) * np.array([[0, 0], [np.sqrt(g), 0]], dtype=complex)]
    raise ValueError(f"Unrecognized circuit noise: {kind}")


def noise_from_kraus(kraus) -> "_lib.Noise":
    """Superoperator sum_k K (x) conj(K) -> the 6 real block coefficients; rejects channels whose
    block form needs more (populations feeding coherences, complex couplings)."""
    s = np.zeros((4, 4), dtype=complex)  # (a,b) <- (c,d): sum K[a,c] conj(K[b,d])
    for k in kraus:
        k = np.asarray(k, dtype=complex)
        s += np.kron(k, np.conj(k))
    allowed = np.zeros((4, 4), dtype=bool)
    for i, j in ((0, 0), (0, 3), (3, 0), (3, 3), (1, 1), (1, 2), (2, 1), (2, 2)):
        allowed[i, j] = True
    if np.abs(s[~allowed]).max() > 1e-14 or np.abs(s.imag).max() > 1e-14:
        raise NotImplementedError("channel does not have the supported block form")
    if abs(s[1, 1] - s[2, 2]) > 1e-14 or abs(s[1, 2] - s[2, 1]) > 1e-14:
        raise NotImplementedError("channel does not have the supported block form")
    if abs(s[0, 0] + s[3, 0] - 1) > 1e-12 or abs(s[0, 3] + s[3, 3] - 1) > 1e-12:
        raise ValueError("channel is not trace preserving")
    nz = _lib.Noise()
    nz.pop[0], nz.pop[1], nz.pop[2], nz.pop[3] = s[0, 0].real, s[0, 3].real, s[3, 0].real, s[3, 3].real
    nz.coh_g, nz.coh_d = s[1, 1].real, s[1, 2].real
    return nz


# ---------------------------------------------------------------------------------------------
# device plan handle
# ---------------------------------------------------------------------------------------------
class DevicePlan:
    """Owns the opaque mbqc_plan* created on the current CUDA device."""

    def __init__(self, plan: LoweredPlan, noise: Optional["_lib.Noise"] = None,
                 n_steps: Optional[int] = None, output_slot: Optional[List[int]] = None, host_only: bool = False):
        """host_only: lowering tables without device allocations (mbqc_plan_create_hostonly) -- such
        a plan cannot be run, only inspected and passed to `jit_compile_check` (no GPU needed)."""
        lib = _lib.load()
        steps = plan.steps if n_steps is None else plan.steps[:n_steps]
        arr = (_lib.Step * max(len(steps), 1))()
        for i, st in enumerate(steps):
            arr[i].slot, arr[i].angle_idx, arr[i].plane = st.slot, st.angle_idx, st.plane
            arr[i].flags = _lib.STEP_APPEND if st.append else 0
            arr[i].fixed_cos, arr[i].fixed_sin = st.fixed_cos, st.fixed_sin
            arr[i].nbr_mask = st.nbr_mask
            arr[i].fixed_z = st.fixed_z
            arr[i].cond_mask, arr[i].cond_table = st.cond_mask, st.cond_table
            arr[i].alt_plane, arr[i].alt_angle_idx = st.alt_plane, st.alt_angle_idx
            arr[i].alt_cos, arr[i].alt_sin, arr[i].alt_z = st.alt_cos, st.alt_sin, st.alt_z
        out_slot = plan.output_slot if output_slot is None else output_slot
        in_arr = (C.c_int32 * max(len(plan.input_slot), 1))(*plan.input_slot)
        cz_arr = (C.c_uint64 * plan.window)(*plan.init_cz_mask)
        out_arr = (C.c_int32 * max(len(out_slot), 1))(*out_slot)
        handle = C.c_void_p()
        create = lib.mbqc_plan_create_hostonly if host_only else lib.mbqc_plan_create
        _lib.check(create(arr, len(steps), plan.window, len(plan.input_slot),
                          len(out_slot), plan.n_angles, in_arr, cz_arr, out_arr,
                          C.byref(noise) if noise is not None else None,
                          C.byref(handle)))
        self.handle = handle
        self.n_out = len(out_slot)
        self.n_in = len(plan.input_slot)
        self.n_steps = len(steps)
        self._lib = lib
        self.has_feedforward = False

    def jit_compile_check(self, out_form: int = 0, cta: int = 128) -> int:
        """Generate and compile (NVRTC, offline) the run-time specialised state-vector kernel of this
        plan; returns the cubin size, 0 when the plan is outside that kernel's scope."""
        n = int(self._lib.mbqc_jit_compile_check(self.handle, out_form, cta))
        if n < 0:
            _lib.check(n)
        return n

    def set_feedforward(self, ff: Sequence[FeedForward]):
        """Attach the correction masks (once, before the plan is used for sampled runs)."""
        if len(ff) != self.n_steps:
            raise ValueError("one FeedForward record per measurement step")
        arr = (_lib.FeedForwardC * max(len(ff), 1))()
        for i, f in enumerate(ff):
            arr[i].xdep, arr[i].zdep, arr[i].outx, arr[i].outz = f.xdep, f.zdep, f.outx, f.outz
        _lib.check(self._lib.mbqc_plan_set_feedforward(self.handle, arr, len(ff)))
        self.has_feedforward = True

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.mbqc_plan_destroy(h)
            except Exception:
                pass
