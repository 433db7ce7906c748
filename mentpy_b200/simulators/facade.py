"""`PatternSimulator` facade with the reference's signature and a CUDA backend table.

Mirror of mentpy/simulators/pattern_simulator.py:19-88: backend name is lower-cased and looked up
in a fixed table (unknown -> ValueError), the default input is |+>^{|I|}, a non-ndarray input is a
TypeError, unknown attributes are forwarded to the backend object.  The table holds the CUDA
backends only -- the product has no numpy path.
"""
from typing import List

import numpy as np

from ..streaming import CudaSimulatorSVStream
from .cuda_backends import CudaSimulatorDM, CudaSimulatorSV

SUPPORTED_BACKENDS = {"cuda-sv": CudaSimulatorSV, "cuda-dm": CudaSimulatorDM,
                      "cuda-sv-stream": CudaSimulatorSVStream}


class PatternSimulator:
    def __init__(self, mbqcircuit, input_state: np.ndarray = None, backend="cuda-sv", *args, **kwargs) -> None:
        backend = backend.lower()
        if backend not in SUPPORTED_BACKENDS:
            raise ValueError(
                f"Backend {backend} not supported. Supported backends are {SUPPORTED_BACKENDS.keys()}"
            )
        if input_state is None:
            input_state = 1
            for _ in range(len(mbqcircuit.input_nodes)):
                input_state = np.kron(input_state, np.array([1, 1]) / np.sqrt(2))
            input_state = np.atleast_1d(input_state)
        elif not isinstance(input_state, np.ndarray):
            raise TypeError(f"Input state must be a numpy array, not {type(input_state)}")
        self.simulator = SUPPORTED_BACKENDS[backend](mbqcircuit, input_state, *args, **kwargs)

    def __getattr__(self, name):
        if name == "simulator":
            raise AttributeError(name)
        return getattr(self.simulator, name)

    def __call__(self, angles: List[float], **kwargs):
        return self.run(angles, **kwargs)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__} ({self.simulator!r})"

    def measure(self, angle: float, **kwargs):
        return self.simulator.measure(angle, **kwargs)

    def run(self, angles: List[float], **kwargs):
        return self.simulator.run(angles, **kwargs)

    def run_batch(self, angles, **kwargs):
        return self.simulator.run_batch(angles, **kwargs)

    def reset(self, input_state: np.ndarray = None):
        return self.simulator.reset(input_state)
