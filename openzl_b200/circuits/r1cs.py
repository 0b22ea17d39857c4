"""Rank-1 constraint systems in the layout the device prover takes.

Plays the role of ``constraint::R1CS<F>`` once synthesis is finished
(/root/reference/plugins/arkworks/src/constraint/mod.rs:64-108): ``A z * B z = C z`` with
``z = (1, instance..., witness...)``.  Matrices are CSR with coefficient *indices* into a small
table (gadget-built systems have few distinct constants), which keeps a 2^20-constraint system
numpy-sized and lets the SpMV kernel keep the table cache-resident.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Csr:
    row_ptr: np.ndarray   # uint32[n_rows + 1]
    col_idx: np.ndarray   # uint32[nnz]
    coef_idx: np.ndarray  # uint32[nnz]

    @property
    def n_rows(self) -> int:
        return len(self.row_ptr) - 1

    def transpose(self, n_cols: int) -> "Csr":
        """CSR of the transpose (used by the setup: a_j(tau) = sum_i L_i(tau) A[i][j])."""
        from scipy.sparse import csr_matrix
        m = csr_matrix((self.coef_idx.astype(np.int64) + 1, self.col_idx.astype(np.int64), self.row_ptr.astype(np.int64)),
                       shape=(self.n_rows, n_cols))
        t = m.T.tocsr()
        t.sort_indices()
        return Csr(t.indptr.astype(np.uint32), t.indices.astype(np.uint32), (t.data - 1).astype(np.uint32))


@dataclass
class R1CS:
    modulus: int
    n_constraints: int
    n_instance: int           # includes the leading constant 1 (ark: num_instance_variables)
    n_vars: int               # n_instance + n_witness
    A: Csr
    B: Csr
    C: Csr
    coef_table: List[int] = field(default_factory=list)   # canonical integers

    def matvec(self, M: Csr, z: List[int]) -> List[int]:
        """Reference (slow, pure Python) product used by tests on small systems."""
        p = self.modulus
        out = []
        for r in range(M.n_rows):
            acc = 0
            for k in range(int(M.row_ptr[r]), int(M.row_ptr[r + 1])):
                acc += self.coef_table[int(M.coef_idx[k])] * z[int(M.col_idx[k])]
            out.append(acc % p)
        return out

    def is_satisfied(self, z: List[int]) -> bool:
        """``R1CS::is_satisfied`` (constraint/mod.rs:101-107)."""
        p = self.modulus
        a, b, c = self.matvec(self.A, z), self.matvec(self.B, z), self.matvec(self.C, z)
        return all((x * y - w) % p == 0 for x, y, w in zip(a, b, c))
