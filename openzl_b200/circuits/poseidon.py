"""Poseidon hash-chain circuit: the "Poseidon-preimage eclair circuit" of BASELINE.json.

Semantics follow the reference's COM-generic Poseidon
(/root/reference/openzl-crypto/src/poseidon/mod.rs:156-283 round structure,
hash.rs:93-135 ``Hasher::hash`` = permutation of [domain_tag, x, y] and take the first element,
lfsr.rs / round_constants.rs / mds.rs for the constants) instantiated as the plugin does for
``bn254::Fr`` arity 2: width 3, 8 full + 55 partial rounds, S-box x^5
(/root/reference/plugins/arkworks/src/poseidon/mod.rs:147-159,300-304).

Statement: "I know (x0, y) such that x_{i+1} = H(x_i, y) for i < links and x_links = digest"
with ``digest`` public.  R1CS shape (our own synthesis -- the arkworks constraint synthesizer is
not available offline): every S-box is three constraints (t^2, t^4, t^5), the two pass-through
lanes of a partial round are re-materialised as fresh variables, and each link's output is one
more constraint: 8*9 + 55*5 + 1 = 348 constraints and 348 fresh variables per link.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

from .r1cs import Csr, R1CS

BN254_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class _GrainLFSR:
    SIZE = 80

    def __init__(self, seed):
        self.state = [False] * self.SIZE
        self.head = 0
        for n, bits in seed:
            for i in reversed(range(n)):
                self._push(((bits >> i) & 1) != 0)
        for _ in range(self.SIZE * 2):
            self._update()

    def _push(self, b):
        self.state[self.head] = b
        self.head = (self.head + 1) % self.SIZE
        return b

    def _bit(self, i):
        return self.state[(i + self.head) % self.SIZE]

    def _update(self):
        return self._push(self._bit(62) ^ self._bit(51) ^ self._bit(38) ^ self._bit(23) ^ self._bit(13) ^ self._bit(0))

    def next_bit(self):
        bit = self._update()
        while not bit:
            self._update()
            bit = self._update()
        return self._update()


@dataclass
class PoseidonParams:
    modulus: int
    width: int
    full_rounds: int
    partial_rounds: int
    round_keys: List[int]
    mds: List[List[int]]
    domain_tag: int

    @classmethod
    def generate(cls, modulus: int = BN254_FR, arity: int = 2, full_rounds: int = 8, partial_rounds: int = 55):
        width = arity + 1
        bits = modulus.bit_length()
        lfsr = _GrainLFSR([(2, 1), (4, 0), (12, bits), (12, width), (10, full_rounds), (10, partial_rounds), (30, (1 << 30) - 1)])
        keys = []
        while len(keys) < width * (full_rounds + partial_rounds):
            v = 0
            for _ in range(bits):
                v = (v << 1) | int(lfsr.next_bit())
            if v < modulus:
                keys.append(v)
        mds = [[pow((x + y) % modulus, -1, modulus) for y in range(width, 2 * width)] for x in range(width)]
        return cls(modulus, width, full_rounds, partial_rounds, keys, mds, (1 << arity) - 1)

    def permute(self, state: List[int]) -> List[int]:
        p, t = self.modulus, self.width
        half = self.full_rounds // 2
        state = list(state)
        for rnd in range(self.full_rounds + self.partial_rounds):
            state = [(s + k) % p for s, k in zip(state, self.round_keys[rnd * t:(rnd + 1) * t])]
            if rnd < half or rnd >= half + self.partial_rounds:
                state = [pow(s, 5, p) for s in state]
            else:
                state[0] = pow(state[0], 5, p)
            state = [sum(self.mds[i][j] * state[j] for j in range(t)) % p for i in range(t)]
        return state

    def hash2(self, x: int, y: int) -> int:
        return self.permute([self.domain_tag, x, y])[0]


# symbolic column codes used inside the one-link template
_ONE, _Y, _XIN = -1, -2, -3


class PoseidonChain:
    """R1CS + witness generator for a chain of ``links`` arity-2 Poseidon hashes."""

    def __init__(self, links: int, params: PoseidonParams | None = None):
        self.params = params or PoseidonParams.generate()
        assert self.params.width == 3
        self.links = links
        self.p = self.params.modulus
        self._build_template()

    # ---- one link, symbolically --------------------------------------------------------------
    def _build_template(self):
        P, p = self.params, self.p
        half = P.full_rounds // 2
        rows_A: List[Dict[int, int]] = []
        rows_B: List[Dict[int, int]] = []
        rows_C: List[Dict[int, int]] = []
        self._prog: List[Tuple] = []   # witness program: ("mul", lcA, lcB) | ("lin", lc)
        nloc = 0

        def lc_add(a, b, kb=1):
            out = dict(a)
            for v, c in b.items():
                out[v] = (out.get(v, 0) + kb * c) % p
            return {v: c for v, c in out.items() if c}

        def new_mul(la, lb):
            nonlocal nloc
            v = nloc
            nloc += 1
            rows_A.append(la); rows_B.append(lb); rows_C.append({v: 1})
            self._prog.append(("mul", la, lb))
            return {v: 1}

        def new_lin(la):
            nonlocal nloc
            v = nloc
            nloc += 1
            rows_A.append(la); rows_B.append({_ONE: 1}); rows_C.append({v: 1})
            self._prog.append(("lin", la))
            return {v: 1}

        def sbox(t):
            a = new_mul(t, t)
            b = new_mul(a, a)
            return new_mul(b, t)

        state = [{_ONE: P.domain_tag}, {_XIN: 1}, {_Y: 1}]
        for rnd in range(P.full_rounds + P.partial_rounds):
            keys = P.round_keys[rnd * 3:(rnd + 1) * 3]
            t = [lc_add(state[i], {_ONE: keys[i]}) for i in range(3)]
            full = rnd < half or rnd >= half + P.partial_rounds
            if full:
                u = [sbox(t[i]) for i in range(3)]
            else:
                u = [sbox(t[0]), t[1], t[2]]
            nxt = []
            for i in range(3):
                acc: Dict[int, int] = {}
                for j in range(3):
                    acc = lc_add(acc, u[j], P.mds[i][j])
                nxt.append(acc)
            if not full:
                nxt[1] = new_lin(nxt[1])
                nxt[2] = new_lin(nxt[2])
            state = nxt
        out = new_lin(state[0])   # x_out = first element of the permutation output
        self.vars_per_link = nloc
        self.constraints_per_link = len(rows_A)
        assert self.constraints_per_link == nloc
        # coefficient table + CSR template
        table: Dict[int, int] = {}

        def flat(rows):
            ptr, cols, cidx = [0], [], []
            for r in rows:
                for v in sorted(r):
                    cols.append(v)
                    cidx.append(table.setdefault(r[v] % p, len(table)))
                ptr.append(len(cols))
            return np.array(ptr, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(cidx, dtype=np.int64)

        self._tA, self._tB, self._tC = flat(rows_A), flat(rows_B), flat(rows_C)
        self.coef_table = [0] * len(table)
        for val, idx in table.items():
            self.coef_table[idx] = val

    # ---- variable numbering --------------------------------------------------------------------
    # z = [ONE, digest | x0, y, link0 locals..., link1 locals..., ...]; the LAST link's output local
    # is replaced by the public digest variable (index 1), so that slot is left unused (zero).
    N_INSTANCE = 2

    def _base(self, link: int) -> int:
        return 4 + link * self.vars_per_link

    @property
    def n_vars(self) -> int:
        return 4 + self.links * self.vars_per_link

    @property
    def n_constraints(self) -> int:
        return self.links * self.constraints_per_link

    def r1cs(self) -> R1CS:
        V = self.vars_per_link
        L = self.links
        bases = 4 + V * np.arange(L, dtype=np.int64)
        xin = np.where(np.arange(L) == 0, 2, bases - 1)          # previous link's output local (V-1), or x0
        out_local = V - 1

        def expand(t):
            ptr, cols, cidx = t
            nnz = len(cols)
            g = np.empty((L, nnz), dtype=np.int64)
            loc = cols >= 0
            g[:, loc] = bases[:, None] + cols[loc][None, :]
            g[:, cols == _ONE] = 0
            g[:, cols == _Y] = 3
            g[:, cols == _XIN] = xin[:, None]
            # the last link's output is the public digest
            last_out = loc & (cols == out_local)
            g[L - 1, last_out] = 1
            rp = (ptr[None, :-1] + (nnz * np.arange(L, dtype=np.int64))[:, None]).reshape(-1)
            rp = np.concatenate([rp, [nnz * L]])
            return Csr(rp.astype(np.uint32), g.reshape(-1).astype(np.uint32), np.tile(cidx, L).astype(np.uint32))

        return R1CS(self.p, self.n_constraints, self.N_INSTANCE, self.n_vars, expand(self._tA), expand(self._tB),
                    expand(self._tC), list(self.coef_table))

    # ---- witness ---------------------------------------------------------------------------------
    def assignment(self, x0: int, y: int) -> List[int]:
        """Full assignment z (canonical ints); z[1] is the digest the chain produces."""
        p = self.p
        V = self.vars_per_link
        z = [0] * self.n_vars
        z[0], z[2], z[3] = 1, x0 % p, y % p
        x = x0 % p
        for link in range(self.links):
            base = self._base(link)
            loc = [0] * V

            def ev(lc):
                acc = 0
                for v, c in lc.items():
                    if v >= 0:
                        acc += c * loc[v]
                    elif v == _ONE:
                        acc += c
                    elif v == _Y:
                        acc += c * y
                    else:
                        acc += c * x
                return acc % p

            for k, op in enumerate(self._prog):
                loc[k] = (ev(op[1]) * ev(op[2])) % p if op[0] == "mul" else ev(op[1])
            x = loc[V - 1]
            z[base:base + V] = loc
        z[1] = x
        z[self._base(self.links - 1) + V - 1] = 0   # slot replaced by the public digest
        return z

    def digest(self, x0: int, y: int) -> int:
        x = x0
        for _ in range(self.links):
            x = self.params.hash2(x, y)
        return x
