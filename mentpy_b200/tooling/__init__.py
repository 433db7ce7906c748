"""Pattern tooling around the simulation path: Pauli operators in the
binary symplectic form, generalised flow, the dynamical Lie algebra of a pattern and the
expressivity estimate -- host-side numpy, except that `expressivity` evaluates its pattern
samples through the batched CUDA simulator."""
from .gf2 import gf2_rank, gf2_solve
from .pauli import PauliOp
from .gflow import find_gflow, verify_gflow
from .lie_algebra import (calculate_complete_gens_lie_algebra, calculate_gens_lie_algebra, calculate_lie_algebra, dim_so,
                          dim_sp, dim_su, graph_stabilizers, lie_algebra_completion, remove_repeated_ops)
from .expressivity import (expressivity_with_histogram, haar_probability_density_of_fidelities,
                           sample_probability_density_of_fidelities)

__all__ = ["PauliOp", "gf2_solve", "gf2_rank", "find_gflow", "verify_gflow", "graph_stabilizers",
           "calculate_complete_gens_lie_algebra", "calculate_gens_lie_algebra", "calculate_lie_algebra",
           "lie_algebra_completion", "remove_repeated_ops", "dim_su", "dim_so", "dim_sp",
           "haar_probability_density_of_fidelities", "sample_probability_density_of_fidelities",
           "expressivity_with_histogram"]
