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


def random_special_unitary(n_qubits: int, random_state=None) -> np.ndarray:
    """Haar-random element of SU(2^n) (the `random_gate` of docs/tutorials/intro-to-mbqml.rst:66)."""
    from scipy.stats import unitary_group

    rng = np.random.default_rng(random_state) if not isinstance(random_state, np.random.Generator) else random_state
    dim = 2**n_qubits
    u = unitary_group.rvs(dim, random_state=rng) if dim > 1 else np.ones((1, 1), dtype=complex)
    return u / np.linalg.det(u) ** (1.0 / dim)


def train_test_split(inputs, targets, test_size: float = 0.3, randomize: bool = False, random_state=None):
    """((x_train, y_train), (x_test, y_test)): the last int(n * test_size) samples form the test
    set (mentpy/utils/generate_data.py:43-57)."""
    inputs, targets = np.asarray(inputs), np.asarray(targets)
    n = len(inputs)
    n_train = n - int(n * test_size)
    if randomize:
        perm = np.random.default_rng(random_state).permutation(n)
        inputs, targets = inputs[perm], targets[perm]
    return (inputs[:n_train], targets[:n_train]), (inputs[n_train:], targets[n_train:])


def generate_random_dataset(unitary: np.ndarray, n_samples: int, test_size: float = 0.3, random_state=None):
    """Haar-random input states and their images under `unitary`, split into train / test
    (mentpy/utils/generate_data.py:33-40) -- as [n, 2^q] arrays, ready for BatchedFidelityCost /
    adam_optimize_batched(dataset=True)."""
    unitary = np.asarray(unitary, dtype=complex)
    n_qubits = int(np.log2(unitary.shape[0]))
    x = np.atleast_2d(generate_haar_random_states(n_qubits, n_samples, random_state=random_state))
    return train_test_split(x, x @ unitary.T, test_size=test_size)


def __getattr__(name):
    # the reference keeps its Lie-algebra / expressivity helpers in mentpy.utils (utils/lie_algebra.py,
    # utils/expressivity.py): same names here, implemented in mentpy_b200.tooling
    from . import tooling

    if name in tooling.__all__:
        return getattr(tooling, name)
    raise AttributeError(name)
