"""Single-qubit measurement descriptor (`Ment`).

Host-side mirror of mentpy/operators/ment.py:123-260 -- same constructor conventions
(`Ment(angle, plane)` or `Ment(plane)` / `Ment(plane, angle)`), same validation and error types,
same `matrix` / `get_povm` values -- because the plan lowering (mentpy_b200/plan.py) reads plane,
fixed angle and trainability from it and the parity tests compare those against the reference.
`MentOutcome` / `ControlMent` mirror mentpy/operators/ment.py:13-120 and
mentpy/operators/controlled_ment.py:14-113: a measurement whose plane / angle depends on a boolean
function of earlier outcomes.  The lowering (plan.py) tabulates that function over its condition
nodes; the density-matrix kernels pick the branch per sample.
"""
import warnings
from typing import Optional, Tuple, Union

import numpy as np

import operator

PLANES = ("XY", "XZ", "YZ", "XYZ", "X", "Y", "Z")


class MentOutcome:
    """Boolean function of the outcome record `x` (dict node -> 0/1), closed under the arithmetic and
    comparison operators (results are taken mod 2), as ment.py:13-120.  `cond_nodes` = the nodes
    whose outcomes it reads."""

    def __init__(self, outcome, node_id=None, cond_nodes=None):
        if isinstance(outcome, (bool, int)):
            value = bool(outcome % 2)
            outcome = lambda *args, **kwargs: value  # noqa: E731
        self._outcome = outcome
        self._node_id = node_id
        if isinstance(outcome, MentOutcome):
            self._cond_nodes = outcome.cond_nodes
        elif cond_nodes is not None:
            self._cond_nodes = cond_nodes
        else:
            self._cond_nodes = {node_id} if node_id is not None else set()

    node_id = property(lambda self: self._node_id)
    cond_nodes = property(lambda self: self._cond_nodes)

    @node_id.setter
    def node_id(self, value):
        self._node_id = value

    def __repr__(self) -> str:
        return "Measurement Outcome"

    def __call__(self, *args, **kwargs):
        try:
            return self._outcome(*args, **kwargs)
        except Exception:
            raise UserWarning("Could not evaluate callable at given")

    def _combine(self, op, other):
        mine = self._outcome
        if isinstance(other, (bool, int)):
            return MentOutcome(lambda x: bool(op(mine(x), other) % 2), cond_nodes=self._cond_nodes)
        if isinstance(other, MentOutcome):
            theirs = other._outcome
            return MentOutcome(lambda x: bool(op(mine(x), theirs(x)) % 2), cond_nodes=self._cond_nodes | other._cond_nodes)
        if callable(other):
            return MentOutcome(lambda x: bool(op(mine(x), other(x)) % 2), cond_nodes=self._cond_nodes)
        raise TypeError(f"Invalid type {type(other)}")

    def __invert__(self):
        mine = self._outcome
        return MentOutcome(lambda x: not mine(x))


for _name, _op in (("mul", operator.mul), ("add", operator.add), ("sub", operator.sub), ("truediv", operator.truediv),
                   ("floordiv", operator.floordiv), ("mod", operator.mod), ("pow", operator.pow), ("eq", operator.eq),
                   ("ne", operator.ne), ("lt", operator.lt), ("le", operator.le), ("gt", operator.gt), ("ge", operator.ge),
                   ("and", lambda a, b: a and b), ("or", lambda a, b: a or b), ("xor", operator.xor)):
    setattr(MentOutcome, f"__{_name}__", (lambda op: lambda self, other: self._combine(op, other))(_op))
MentOutcome.__hash__ = object.__hash__  # __eq__ builds a new condition, identity hashing stays
_AXIS = {
    "X": np.array([[0, 1], [1, 0]]),
    "Y": np.array([[0, -1j], [1j, 0]]),
    "Z": np.array([[1, 0], [0, -1]]),
}
_PAIR = {"XY": ("X", "Y"), "XZ": ("X", "Z"), "YZ": ("Y", "Z")}


class Ment:
    """Measurement of one qubit in `plane` at `angle` (None = trainable)."""

    def __init__(self, angle: Optional[Union[int, float, tuple, str]] = None,
                 plane: Optional[str] = "XY"):
        if isinstance(angle, str):
            # Ment("XY") or Ment("XY", 0.3): first positional is the plane
            angle, plane = (plane if isinstance(plane, (int, float, tuple)) else None), angle
        elif angle is not None and not isinstance(angle, (int, float, tuple)):
            raise TypeError(f"Invalid argument type. Expected float or str but got {type(angle)}")
        elif plane is None:
            plane = "XY"
        plane = plane.upper()
        if plane not in PLANES:
            raise ValueError(f"Plane {plane} is not supported.")
        if plane == "XYZ":
            warnings.warn("Plane XYZ might be unstable. Use at your own risk.")
        if plane in _AXIS:
            if angle is not None and angle != 0:
                raise ValueError(f"Plane {plane} does not support angle.")
            angle = 0
        self._plane = plane
        self._angle = angle
        self._node_id = -1
        self._outcome = MentOutcome(lambda x, n=-1: x[n])

    @property
    def outcome(self) -> MentOutcome:
        """The node's entry of the outcome record, as a condition for `ControlMent` (ment.py:190-192)."""
        return self._outcome

    @property
    def plane(self) -> str:
        return self._plane

    @property
    def angle(self):
        return self._angle

    @property
    def node_id(self):
        return self._node_id

    @node_id.setter
    def node_id(self, value):
        self._node_id = value
        self._outcome = MentOutcome(lambda x, n=value: x[n], value)

    def set_angle(self, angle) -> "Ment":
        self._angle = angle
        return self

    def copy(self) -> "Ment":
        return Ment(self._angle, self._plane)

    def is_trainable(self) -> bool:
        return self._angle is None and self._plane in ("XY", "XZ", "YZ", "XYZ")

    def _resolve(self, angle):
        if self._angle is None:
            if angle is None:
                raise ValueError("Measurement is trainable, please provide an angle.")
            return angle
        if angle is not None and self._angle != angle:
            raise ValueError(f"Measurement has a fixed angle of {round(self._angle, 4)}")
        return self._angle

    def matrix(self, angle=None, *args, **kwargs) -> np.ndarray:
        """2x2 observable n.sigma of the measurement (ment.py:218-253)."""
        angle = self._resolve(angle)
        if self._plane in _AXIS:
            return _AXIS[self._plane]
        if self._plane in _PAIR:
            a, b = _PAIR[self._plane]
            return np.cos(angle) * _AXIS[a] + np.sin(angle) * _AXIS[b]
        if not isinstance(angle, tuple):
            raise TypeError(f"Invalid argument type. Expected tuple but got {type(angle)}")
        t1, t2 = angle
        return (np.cos(t1) * np.cos(t2) * _AXIS["X"] + np.sin(t1) * np.cos(t2) * _AXIS["Y"]
                + np.sin(t2) * _AXIS["Z"])

    def get_povm(self, angle=None, *args, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
        """Projectors (I +- M)/2 (ment.py:255-260)."""
        m = self.matrix(angle, *args, **kwargs)
        return (np.eye(2) + m) / 2, (np.eye(2) - m) / 2

    def __repr__(self) -> str:
        if isinstance(self._angle, tuple):
            shown = (round(self._angle[0], 4), round(self._angle[1], 4))
        elif isinstance(self._angle, (int, float)):
            shown = round(self._angle, 4)
        else:
            shown = "θ"
        return f"Ment({shown}, {self._plane})"


Measurement = Ment


class _Probe(dict):
    """Outcome record that logs which nodes a condition looks at (missing entries read as 0)."""

    def __init__(self, values, seen):
        super().__init__(values)
        self._seen = seen

    def __getitem__(self, key):
        self._seen.add(key)
        return dict.get(self, key, 0)


def condition_reads(cond):
    """Nodes whose outcomes the condition actually reads, found by evaluating it on probing records
    (`cond_nodes` is not reliable: `~outcome` drops it in the reference, ment.py:118-119)."""
    seen = set(getattr(cond, "cond_nodes", ()) or ())
    while True:
        before = set(seen)
        keys = sorted(before)
        for idx in range(1 << len(keys)):
            cond(_Probe({k: (idx >> i) & 1 for i, k in enumerate(keys)}, seen))
        if seen == before:
            return sorted(seen)


class ControlMent(Ment):
    """Measurement that takes `true_*` when `condition(outcomes)` holds and `false_*` otherwise
    (controlled_ment.py:14-113).  As in the reference `angle` and `plane` are METHODS here: without
    arguments they describe the false branch, with the outcome record they evaluate the condition."""

    def __init__(self, condition=None, true_angle=None, true_plane: Optional[str] = "XY",
                 false_angle=0, false_plane: Optional[str] = "X"):
        super().__init__(false_angle, false_plane)
        self._true_ment = Ment(true_angle, true_plane)
        self._condition = condition

    def __repr__(self) -> str:
        return f"ControlMent(False: {Ment.__repr__(self)}, True: {repr(self._true_ment)})"

    @property
    def condition(self):
        if isinstance(self._condition, bool):
            return lambda x: self._condition
        if isinstance(self._condition, MentOutcome):
            return self._condition
        return None

    @condition.setter
    def condition(self, condition):
        if not isinstance(condition, (bool, MentOutcome)):
            raise TypeError(f"Invalid argument type. Expected bool or MentOutcome but got {type(condition)}")
        self._condition = condition

    true_ment = property(lambda self: self._true_ment)
    false_ment = property(lambda self: Ment(self._angle, self._plane))

    def angle(self, *args, **kwargs):
        if not args and not kwargs:
            if isinstance(self._condition, bool):
                return self._true_ment.angle if self._condition else self._angle
            if self._true_ment.angle is None or self._angle is None:
                return None
            return self._angle
        return self._true_ment.angle if self.condition(*args, **kwargs) else self._angle

    def plane(self, *args, **kwargs):
        if not args and not kwargs:
            return self._plane
        return self._true_ment.plane if self.condition(*args, **kwargs) else self._plane

    def is_trainable(self) -> bool:
        return Ment.is_trainable(self) or self._true_ment.is_trainable()

    def copy(self) -> "ControlMent":
        return ControlMent(self.condition, self._true_ment.angle, self._true_ment.plane, self._angle, self._plane)

    def _branch(self, *args, **kwargs) -> Ment:
        return self._true_ment if self.condition(*args, **kwargs) else Ment(self._angle, self._plane)

    def matrix(self, angle=None, *args, **kwargs) -> np.ndarray:
        if not self.is_trainable() and angle is not None:
            raise ValueError("ControlledMent is not trainable, so angle must be None.")
        branch = self._branch(*args, **kwargs)
        return branch.matrix(angle) if branch.is_trainable() else branch.matrix()

    def get_povm(self, angle=None, *args, **kwargs):
        m = self.matrix(angle, *args, **kwargs)
        return (np.eye(2) + m) / 2, (np.eye(2) - m) / 2


ControlledMent = ControlMent
