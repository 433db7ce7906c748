"""Simulator backends (CUDA only) and the PatternSimulator facade."""
from .backend_base import BaseSimulator
from .cuda_backends import CudaSimulatorDM, CudaSimulatorSV
from .facade import SUPPORTED_BACKENDS, PatternSimulator

__all__ = ["BaseSimulator", "CudaSimulatorSV", "CudaSimulatorDM", "PatternSimulator", "SUPPORTED_BACKENDS"]
