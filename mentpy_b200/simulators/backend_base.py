"""Abstract simulator interface -- the drop-in boundary.

Same surface as mentpy/simulators/base_simulator.py:13-101: constructor (mbqcircuit,
input_state), properties `mbqcircuit`, `input_state`, `outcomes`, `__call__ -> run`, and the
abstract trio `measure`, `run`, `reset`.
"""
import abc
from typing import List

import numpy as np


class BaseSimulator(abc.ABC):
    def __init__(self, mbqcircuit, input_state: np.ndarray = None) -> None:
        self._mbqcirc = mbqcircuit
        self._input_state = input_state
        self._outcomes = {}

    @property
    def mbqcircuit(self):
        return self._mbqcirc

    @property
    def input_state(self) -> np.ndarray:
        return self._input_state

    @input_state.setter
    def input_state(self, value: np.ndarray):
        self._input_state = value

    @property
    def outcomes(self) -> dict:
        return self._outcomes

    @outcomes.setter
    def outcomes(self, value: dict):
        self._outcomes = value

    def __call__(self, angles: List[float], **kwargs):
        return self.run(angles, **kwargs)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__} for {self.mbqcircuit}"

    @abc.abstractmethod
    def measure(self, angle: float, **kwargs):
        """Measure the next qubit of the schedule at `angle`."""

    @abc.abstractmethod
    def run(self, angles: List[float], **kwargs):
        """Run the whole pattern for one angle vector."""

    @abc.abstractmethod
    def reset(self, input_state=None):
        """Back to the seeded window."""
