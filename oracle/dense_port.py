"""TEST INFRASTRUCTURE ONLY -- dense CPU restatement ("port") of the reference hot path.

This is the checker and the CPU baseline, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(mentpy_b200/) must never route through it.

It restates, operator for operator, what the reference numpy simulators do -- including the dense
2^n x 2^n Kronecker-product operators that dominate the reference's run time -- so that timing it
is a fair stand-in for `PatternSimulator(..., backend="numpy-sv"|"numpy-dm").run(angles)`:

  * SV:  mentpy/simulators/np_simulator_sv.py:38-128 (seed), :164-225 (measure), :227-297 (run),
         :299-320 (reset), :322-358 (measure_ment), :360-384 (reorder)
  * DM:  mentpy/simulators/np_simulator_dm.py:33-115, :151-216, :218-283, :285-305, :307-346
  * dense operators: mentpy/operators/gates.py:62-72 (1-qubit embed), :75-98 (swap), :127-143 (CZ)
  * partial traces:  mentpy/calculator/state_ops.py:42-74 (pure: SUM + renormalise), :77-119 (mixed)
  * projectors:      mentpy/operators/ment.py:218-260

Parity pinning: checked in tests/test_oracle_golden.py against golden vectors produced by the
unmodified reference (tests/golden/*.json, generator oracle/gen_golden.py), and -- when the
reference tree is present -- against the imported reference directly.
"""
from functools import reduce

import numpy as np

from .pattern_data import PatternData

_I2 = np.eye(2)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_PLUS = np.array([1.0, 1.0]) / np.sqrt(2)
_KET = (np.array([[1.0], [0.0]]), np.array([[0.0], [1.0]]))


def _kron_all(factors):
    return reduce(np.kron, factors, 1)


def embed_one_qubit(u, pos, n):
    """I x ... x u x ... x I with u at window position `pos` (0 = MSB).  gates.py:62-72."""
    return _kron_all([u if k == pos else _I2 for k in range(n)])


def dense_cz(i, j, n):
    """I - 2 |11><11|_{ij} as a dense matrix.  gates.py:127-143."""
    p11 = _KET[1] @ _KET[1].T
    return _kron_all([_I2] * n) - 2 * _kron_all([p11 if k in (i, j) else _I2 for k in range(n)])


def dense_swap(i, j, n):
    """SWAP_{ij} = sum_{ab} |a b><b a| on (i, j).  gates.py:75-98."""
    total = 0
    for a in (0, 1):
        for b in (0, 1):
            fac = []
            for k in range(n):
                if k == i:
                    fac.append(_KET[a] @ _KET[b].T)
                elif k == j:
                    fac.append(_KET[b] @ _KET[a].T)
                else:
                    fac.append(_I2)
            total = total + _kron_all(fac)
    return total


def observable(plane, angle):
    """2x2 measurement observable.  ment.py:228-251."""
    if plane == "XYZ":
        if not isinstance(angle, tuple):
            raise TypeError(f"Invalid argument type. Expected tuple but got {type(angle)}")
        a1, a2 = angle
        return np.cos(a1) * np.cos(a2) * _X + np.sin(a1) * np.cos(a2) * _Y + np.sin(a2) * _Z
    if plane == "XY":
        return np.cos(angle) * _X + np.sin(angle) * _Y
    if plane in ("X", "Y", "Z"):
        return {"X": _X, "Y": _Y, "Z": _Z}[plane]
    if plane == "XZ":
        return np.cos(angle) * _X + np.sin(angle) * _Z
    if plane == "YZ":
        return np.cos(angle) * _Y + np.sin(angle) * _Z
    raise ValueError(f"Plane {plane} is not supported.")


def projectors(plane, angle):
    """(I +- M)/2.  ment.py:255-260."""
    m = observable(plane, angle)
    return (_I2 + m) / 2, (_I2 - m) / 2


def sum_trace_pure(psi, idx):
    """Reference 'partial trace' of a pure state: SUM over the traced qubit, renormalise.
    state_ops.py:42-74."""
    n = int(np.log2(psi.shape[0]))
    keep = [k for k in range(n) if k != idx]
    t = psi.reshape([2] * n).transpose(keep + [idx]).reshape(-1, 2).sum(axis=1)
    return t / np.linalg.norm(t)


def trace_mixed(rho, idx):
    """True partial trace via dense isometries V_m.  state_ops.py:77-119."""
    n = int(np.log2(rho.shape[0]))
    sigma = np.zeros((2 ** (n - 1), 2 ** (n - 1)), dtype=complex)
    for m in (0, 1):
        v = _kron_all([_KET[m] if k == idx else _I2 for k in range(n)])
        sigma += v.conj().T @ rho @ v
    return sigma


def swap_sequence(source, target):
    """Selection-sort swap list.  np_simulator_sv.py:360-374."""
    src = list(source)
    assert set(src) == set(target)
    swaps = []
    for i, want in enumerate(target):
        if src[i] != want:
            j = src.index(want, i + 1)
            src[i], src[j] = src[j], src[i]
            swaps.append((i, j))
    return swaps


def reorder(state, current, target):
    """np_simulator_sv.py:376-384 / np_simulator_dm.py:364-380."""
    out = state.copy()
    for i, j in swap_sequence(current, target):
        s = dense_swap(i, j, len(current))
        out = s @ out if out.ndim == 1 else s @ out @ s.conj().T
    return out


def default_input(n_inputs):
    """|+>^{|I|}.  pattern_simulator.py:58-61."""
    st = 1
    for _ in range(n_inputs):
        st = np.kron(st, _PLUS)
    return np.asarray(st, dtype=float)


class _DenseBase:
    mixed = False

    def __init__(self, pat: PatternData, input_state=None, window_size=1, schedule=None):
        self.pat = pat
        self.n_total = pat.n_nodes
        out_excl = pat.quantum_output_nodes if self.mixed else pat.output_nodes
        if schedule is not None:
            self.schedule = list(schedule)
        elif pat.measurement_order is not None:
            self.schedule = list(pat.measurement_order)
            if window_size == 1:
                window_size = len(pat.input_nodes) + 1
        else:
            raise ValueError("Schedule must be provided")
        self.schedule_measure = [v for v in self.schedule if v not in out_excl]
        self.window_size = window_size
        n_in = len(pat.input_nodes)
        if n_in > window_size:
            raise ValueError("window too small for the input state")
        if window_size > len(self.schedule_measure):
            raise ValueError("window larger than the number of measurements")
        if input_state is None:
            input_state = default_input(n_in)
        first = self.schedule[:window_size]
        cz = np.eye(2**window_size)
        for a, b in pat.edges:
            if a in first and b in first:
                cz = dense_cz(first.index(a), first.index(b), window_size) @ cz
        self.initial_czs = cz
        self.outcomes = {}
        self.reset(input_state)

    # -- bookkeeping: np_simulator_sv.py:130-142 ------------------------------------------------
    def window(self):
        return self.schedule[self.cm : self.cm + self.window_size]

    def live(self):
        return min(self.window_size, self.n_total - self.cm)

    def reset(self, input_state=None):
        self.cm = 0
        if input_state is not None:
            n_in = len(self.pat.input_nodes)
            st = reorder(
                np.asarray(input_state, dtype=complex), self.pat.input_nodes, self.schedule[:n_in]
            )
            for _ in range(self.window_size - n_in):
                st = np.kron(st, _PLUS)
            self.seed = st
        self.state = self._seed_state(self.seed)
        self.outcomes = {}

    def _angle_for(self, node, angles):
        plane, fixed = self.pat.measurements[node]
        if node in self.pat.trainable_nodes:
            return angles[self.pat.trainable_nodes.index(node)]
        return fixed

    def run_loop(self, angles):
        if len(angles) != len(self.pat.trainable_nodes):
            raise ValueError("Number of angles does not match number of trainable nodes")
        for node in self.schedule_measure:
            self.measure(self._angle_for(node, angles))

    def measure(self, angle):
        if self.cm >= len(self.schedule_measure):
            raise ValueError("No more measurements to be done.")
        node = self.schedule_measure[self.cm]
        plane, _ = self.pat.measurements[node]
        if node in self.pat.controls:
            # controlled_ment.py:96-113: the condition picks the branch; a fixed branch ignores the angle
            plane, fixed = self.pat.control_branch(node, self.outcomes)
            if fixed is not None:
                angle = fixed
        outcome = self._project(plane, angle)
        self.outcomes[node] = outcome
        self.cm += 1
        self._trace_first()
        if self.cm + self.window_size <= self.n_total:
            self._append_plus()
            win = self.window()
            new = win[-1]
            for nb in self.pat.neighbors(new):
                if nb in win:
                    self._apply(dense_cz(win.index(nb), win.index(new), self.window_size))
        return self.state, outcome


class DensePatternSV(_DenseBase):
    """Restates NumpySimulatorSV (np_simulator_sv.py:35-384), force0 only."""

    mixed = False

    def __init__(self, pat, input_state=None, window_size=1, schedule=None):
        for node, m in pat.measurements.items():
            if m is not None and m[0] not in ("X", "Y", "XY"):
                raise ValueError(f"Node {node} has plane {m[0]}, but only XY plane is supported.")
        super().__init__(pat, input_state, window_size, schedule)

    def _seed_state(self, seed):
        return self.initial_czs @ seed

    def _project(self, plane, angle):
        p0, p1 = projectors(plane, angle)
        n = self.live()
        p1e = embed_one_qubit(p1, 0, n)
        p0e = embed_one_qubit(p0, 0, n)
        _prob0 = np.dot(np.conj(self.state), p0e @ self.state)  # computed, unused under force0
        _prob1 = np.dot(np.conj(self.state), p1e @ self.state)
        st = p0e @ self.state
        self.state = st / np.linalg.norm(st)
        return 0

    def _trace_first(self):
        self.state = sum_trace_pure(self.state, 0)

    def _append_plus(self):
        self.state = np.kron(self.state, _PLUS)

    def _apply(self, op):
        self.state = op @ self.state

    def run(self, angles, output_form="sv"):
        self.run_loop(angles)
        current = self.window()
        if self.pat.quantum_output_nodes != current:
            self.state = reorder(self.state, current, self.pat.output_nodes)
        if output_form == "dm":
            return np.outer(self.state, np.conj(self.state))
        return self.state


class DensePatternDM(_DenseBase):
    """Restates NumpySimulatorDM (np_simulator_dm.py:29-380), force0 only, deterministic planes."""

    mixed = True

    def _seed_state(self, seed):
        rho = np.outer(seed, np.conj(seed))
        return self.initial_czs @ rho @ self.initial_czs.conj().T

    def _project(self, plane, angle):
        if plane == "Z":
            raise NotImplementedError("plane Z is sampled by the reference even under force0")
        p0, p1 = projectors(plane, angle)
        n = self.live()
        p0e = embed_one_qubit(p0, 0, n)
        p1e = embed_one_qubit(p1, 0, n)
        prob0 = np.real(np.trace(self.state @ p0e))
        prob1 = np.real(np.trace(self.state @ p1e))
        outcome = 1 if prob0 < 1e-4 else 0  # np_simulator_dm.py:335-338
        if outcome == 0:
            self.state = p0e @ self.state @ p0e.conj().T / prob0
        else:
            self.state = p1e @ self.state @ p1e.conj().T / prob1
        if np.isnan(self.state).any():
            raise ValueError("qstate has nan, you might want to increase the window size")
        return outcome

    def _trace_first(self):
        self.state = trace_mixed(self.state, 0)

    def _append_plus(self):
        self.state = np.kron(self.state, np.outer(_PLUS, _PLUS))

    def _apply(self, op):
        self.state = op @ self.state @ op.conj().T

    def run(self, angles):
        self.run_loop(angles)
        current = [v for v in self.schedule if v not in self.schedule_measure]
        if self.pat.quantum_output_nodes != current:
            self.state = reorder(self.state, current, self.pat.quantum_output_nodes)
        return self.state


def run_sv(pat, angles, input_state=None, window_size=1, schedule=None, output_form="sv"):
    return DensePatternSV(pat, input_state, window_size, schedule).run(angles, output_form)


def run_dm(pat, angles, input_state=None, window_size=1, schedule=None):
    return DensePatternDM(pat, input_state, window_size, schedule).run(angles)
