"""Single-qubit measurement descriptor (`Ment`).

Host-side mirror of mentpy/operators/ment.py:123-260 -- same constructor conventions
(`Ment(angle, plane)` or `Ment(plane)` / `Ment(plane, angle)`), same validation and error types,
same `matrix` / `get_povm` values -- because the plan lowering (mentpy_b200/plan.py) reads plane,
fixed angle and trainability from it and the parity tests compare those against the reference.
Outcome-conditioned measurements (ControlMent, mentpy/operators/controlled_ment.py) are out of
scope for the accelerated path.
"""
import warnings
from typing import Optional, Tuple, Union

import numpy as np

PLANES = ("XY", "XZ", "YZ", "XYZ", "X", "Y", "Z")
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
