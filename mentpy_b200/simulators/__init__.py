"""Simulator backends (CUDA only) and the PatternSimulator facade."""
from .backend_base import BaseSimulator
from ..streaming import CudaSimulatorSVStream
from .cuda_backends import CudaSimulatorDM, CudaSimulatorSV
from .facade import SUPPORTED_BACKENDS, PatternSimulator

__all__ = ["BaseSimulator", "CudaSimulatorSV", "CudaSimulatorDM", "CudaSimulatorSVStream", "PatternSimulator", "SUPPORTED_BACKENDS"]
