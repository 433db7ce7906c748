"""mentpy_b200 -- B200-native MBQC pattern simulator behind MentPy's simulator API.

Importing the package never needs a GPU (the host-side pattern layer is pure Python); any call
that simulates needs the in-tree CUDA library and a CUDA device and raises otherwise.
"""
from . import mbqc
from .mbqc import (ControlledMent, ControlMent, GraphState, MBQCircuit, Measurement, Ment, MentOutcome, hstack, merge, templates, vstack)
from . import calculator, gates, gradients, optimizers, simulators, tooling, utils
from .tooling import PauliOp
from .simulators import BaseSimulator, CudaSimulatorDM, CudaSimulatorSV, PatternSimulator



def pinned_empty(shape, dtype="float64"):
    """Page-locked host array (numpy view + owning tensor) for zero-staging H2D/D2H in run_batch."""
    import torch

    t = torch.empty(shape, dtype=getattr(torch, dtype)).pin_memory()
    return t


__version__ = "0.1.0"
