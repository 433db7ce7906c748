"""Expressivity of a pattern: how close the fidelity distribution of its outputs is to Haar
(the role of mentpy/utils/expressivity.py:21-125, whose sampling loop goes through a PennyLane
circuit that no longer exists in the reference).  The samples -- random input states through the
pattern at random angles -- are ONE batched launch of the CUDA simulator."""
from typing import Optional

import numpy as np


def haar_probability_density_of_fidelities(F, n_qubits: int):
    """P_Haar(F) = (N - 1) (1 - F)^(N - 2), N = 2^n (expressivity.py:21-36)."""
    N = int(2**n_qubits)
    return (N - 1) * ((1 - np.asarray(F)) ** (N - 2))


def sample_probability_density_of_fidelities(circuit, n_samples: int = 1000, backend: str = "cuda-sv",
                                              seed: Optional[int] = None, simulator=None):
    """n_samples fidelities |<in|out>|^2 between a Haar-random input state and the pattern's output for
    uniformly random angles (the quantity expressivity.py:90-125 samples)."""
    from ..simulators import PatternSimulator
    from ..utils import generate_haar_random_states

    if len(circuit.input_nodes) != len(circuit.output_nodes):
        raise ValueError("the fidelity between input and output needs as many outputs as inputs")
    rng = np.random.default_rng(seed)
    n_in = len(circuit.input_nodes)
    states = np.asarray(generate_haar_random_states(n_in, n_samples, None if seed is None else int(rng.integers(1 << 31))))
    angles = rng.uniform(0, 2 * np.pi, (n_samples, len(circuit.trainable_nodes)))
    ps = simulator if simulator is not None else PatternSimulator(circuit, backend=backend)
    out = ps.run_batch(angles, input_states=states)
    if out.ndim == 3:  # density matrices
        return np.real(np.einsum("bi,bij,bj->b", states.conj(), out, states))
    return np.abs(np.einsum("bi,bi->b", states.conj(), out)) ** 2


def expressivity_with_histogram(circuit, n_samples: int = 10000, n_bins: int = 1000, method: str = "KL",
                                backend: str = "cuda-sv", seed: Optional[int] = None, samples=None) -> float:
    """Divergence between the histogram of sampled fidelities and the Haar density (expressivity.py:39-87):
    'KL' Kullback-Leibler, 'RE' relative entropy, 'JS' Jensen-Shannon distance."""
    if method not in ("KL", "RE", "JS"):
        raise UserWarning("Unsupported method for calculating expressivity")
    if samples is None:
        samples = sample_probability_density_of_fidelities(circuit, n_samples=n_samples, backend=backend, seed=seed)
    density, edges = np.histogram(samples, bins=n_bins, density=True, range=(0, 1))
    width = np.diff(edges)
    p = density * width
    q = haar_probability_density_of_fidelities(0.5 * (edges[:-1] + edges[1:]), len(circuit.output_nodes)) * width
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(p > 0, p * np.log(p / q), 0.0)
        if method == "RE":
            return float(rel.sum())
        if method == "KL":  # scipy.special.kl_div: x log(x/y) - x + y
            return float((rel - p + q).sum())
        m = 0.5 * (p / p.sum() + q / q.sum())
        pn, qn = p / p.sum(), q / q.sum()
        js = 0.5 * np.where(pn > 0, pn * np.log(pn / m), 0.0).sum() + 0.5 * np.where(qn > 0, qn * np.log(qn / m), 0.0).sum()
        return float(np.sqrt(max(js, 0.0)))
