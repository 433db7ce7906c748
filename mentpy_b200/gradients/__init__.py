"""Gradients of scalar costs over measurement angles (mentpy.gradients API + batched twins)."""
from .shift_rules import (fd_gradient, fd_hessian, get_gradient, get_hessian, psr_gradient,
                          psr_gradient_batched, psr_gradient_dataset, psr_hessian)

__all__ = ["get_gradient", "get_hessian", "psr_gradient", "psr_hessian", "fd_gradient", "fd_hessian",
           "psr_gradient_batched", "psr_gradient_dataset"]
