"""Optimisers with the mentpy.optimizers API, running on batched CUDA cost evaluations."""
from .device_loop import adam_optimize_batched, sgd_optimize_batched
from .descent import (AdamOptimizer, BaseOptimizer, BatchedFidelityCost, RCDOptimizer, SGDOptimizer,
                      compute_gradient_variance)

__all__ = ["AdamOptimizer", "SGDOptimizer", "RCDOptimizer", "BaseOptimizer", "BatchedFidelityCost",
           "compute_gradient_variance", "adam_optimize_batched", "sgd_optimize_batched"]
