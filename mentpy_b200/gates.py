"""Single-qubit constants with the names of mentpy/operators/gates.py (PauliX, PauliY, PauliZ, ...).

Only the constants: the reference's dense n-qubit builders (`controlled_z`, `arbitrary_qubit_gate`,
`swap_ij`; gates.py:62-143) are exactly what the CUDA path replaces with index arithmetic and are
deliberately not reproduced."""
import numpy as np

PauliX = np.array([[0, 1], [1, 0]], dtype=complex)
PauliY = np.array([[0, -1j], [1j, 0]], dtype=complex)
PauliZ = np.array([[1, 0], [0, -1]], dtype=complex)
Identity = np.eye(2, dtype=complex)
Hadamard = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
qubit_plus = np.array([1, 1], dtype=complex) / np.sqrt(2)
