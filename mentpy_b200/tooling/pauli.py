"""Pauli operators (up to phase) as rows of a binary symplectic matrix [x | z].

Same surface as mentpy/operators/pauliop.py:16-211 -- construction from a matrix, a ';'-separated
string or a list of strings, `txt`, indexing, product, commutator, symplectic product, `append`,
`get_subset` -- on uint8 arrays instead of galois field arrays."""
from typing import List, Union

import numpy as np

_LETTER = {(0, 0): "I", (1, 0): "X", (0, 1): "Z", (1, 1): "Y"}
_BITS = {v: k for k, v in _LETTER.items()}


class PauliOp:
    def __init__(self, op: Union[np.ndarray, str, List[str]]):
        if isinstance(op, np.ndarray):
            if op.ndim != 2 or op.shape[1] % 2:
                raise ValueError("Tableau representation must have an even number of columns")
            self.matrix = np.asarray(op, dtype=np.uint8) & 1
        elif isinstance(op, (str, list)):
            rows = op.replace(" ", "").rstrip(";").split(";") if isinstance(op, str) else list(op)
            if not rows or any(len(r) != len(rows[0]) for r in rows):
                raise ValueError("All Pauli operators must be the same length")
            n = len(rows[0])
            self.matrix = np.zeros((len(rows), 2 * n), dtype=np.uint8)
            for i, word in enumerate(rows):
                for q, ch in enumerate(word):
                    if ch not in _BITS:
                        raise ValueError(f"'{ch}' is not a Pauli letter")
                    self.matrix[i, q], self.matrix[i, q + n] = _BITS[ch]
        else:
            raise ValueError("PauliOp must be initialized with a string or a numpy array")

    # -- views ------------------------------------------------------------------------------------
    @property
    def number_of_qubits(self) -> int:
        return self.matrix.shape[1] // 2

    @property
    def txt(self) -> str:
        n = self.number_of_qubits
        return "\n".join("".join(_LETTER[(int(r[q]), int(r[q + n]))] for q in range(n)) for r in self.matrix)

    def __repr__(self) -> str:
        return self.txt

    def __len__(self) -> int:
        return self.matrix.shape[0]

    def __getitem__(self, key) -> "PauliOp":
        rows = self.matrix[key]
        return PauliOp(rows[None, :] if rows.ndim == 1 else rows)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __contains__(self, item: "PauliOp") -> bool:
        have = {r.tobytes() for r in self.matrix}
        return all(r.tobytes() in have for r in item.matrix)

    def __hash__(self) -> int:
        return hash((self.matrix.shape, self.matrix.tobytes()))

    def __eq__(self, other) -> bool:
        return isinstance(other, PauliOp) and self.matrix.shape == other.matrix.shape and bool(np.all(self.matrix == other.matrix))

    # -- algebra (phases dropped) -------------------------------------------------------------------
    def __mul__(self, other: "PauliOp") -> "PauliOp":
        return PauliOp(self.matrix ^ other.matrix)

    def symplectic_prod(self, other: "PauliOp") -> np.ndarray:
        """[len(self), len(other)] matrix: 1 where the two operators anticommute."""
        n = self.number_of_qubits
        x1, z1 = self.matrix[:, :n].astype(np.int64), self.matrix[:, n:].astype(np.int64)
        x2, z2 = other.matrix[:, :n].astype(np.int64), other.matrix[:, n:].astype(np.int64)
        return ((x1 @ z2.T + z1 @ x2.T) & 1).astype(np.uint8)

    def commutator(self, other: "PauliOp"):
        """0 when the (single) operators commute, their product otherwise."""
        if not np.any(self.symplectic_prod(other)):
            return 0
        return self * other

    def append(self, other: "PauliOp") -> None:
        self.matrix = np.vstack((self.matrix, other.matrix))

    def get_subset(self, indices: List[int]) -> "PauliOp":
        """The operators restricted to the given qubits."""
        n = self.number_of_qubits
        idx = list(indices)
        if max(idx) >= n:
            raise ValueError("Index out of range")
        return PauliOp(self.matrix[:, idx + [i + n for i in idx]])
