"""Small helpers used by tests and training scripts (mentpy/utils/generate_data.py:15-30)."""
import numpy as np


def generate_haar_random_states(n_qubits: int, n_samples: int = 1, random_state=None) -> np.ndarray:
    """n_samples Haar-random n-qubit states: first columns of Haar unitaries."""
    from scipy.stats import unitary_group

    rng = np.random.default_rng(random_state) if not isinstance(random_state, np.random.Generator) else random_state
    dim = 2**n_qubits
    if dim == 1:
        return np.ones((n_samples, 1), dtype=complex)
    return np.array([unitary_group.rvs(dim, random_state=rng)[:, 0] for _ in range(n_samples)])
