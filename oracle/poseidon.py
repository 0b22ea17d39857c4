"""Poseidon permutation restatement used ONLY to pin the oracle's Fr arithmetic against the
reference's golden vectors (oracle; test infrastructure only).

Follows ``/root/reference/openzl-crypto/src/poseidon/lfsr.rs:14-100`` (Grain LFSR),
``round_constants.rs:10-59`` (rejection sampling of round constants, big-endian bits),
``mds.rs:84-102`` (Cauchy matrix 1/(x_i + y_j), x_i = i, y_j = t + j) and
``mod.rs:156-283`` (round structure: full = ARK + S-box on all + MDS; partial = ARK on all +
S-box on element 0 + MDS), with S-box x^5
(``/root/reference/plugins/arkworks/src/poseidon/mod.rs:147-159``).
"""
from __future__ import annotations

from .fields import Field


class GrainLFSR:
    SIZE = 80

    def __init__(self, seed):
        self.state = [False] * self.SIZE
        self.head = 0
        for n, bits in seed:
            for i in reversed(range(n)):
                self._set_next(((bits >> i) & 1) != 0)
        for _ in range(self.SIZE * 2):
            self._update()

    def _set_next(self, b):
        self.state[self.head] = b
        self.head = (self.head + 1) % self.SIZE
        return b

    def _bit(self, i):
        return self.state[(i + self.head) % self.SIZE]

    def _update(self):
        return self._set_next(self._bit(62) ^ self._bit(51) ^ self._bit(38) ^ self._bit(23)
                              ^ self._bit(13) ^ self._bit(0))

    def next_bit(self):
        bit = self._update()
        while not bit:
            self._update()
            bit = self._update()
        return self._update()


def generate_round_constants(field: Field, width: int, full_rounds: int, partial_rounds: int):
    lfsr = GrainLFSR([(2, 1), (4, 0), (12, field.bits), (12, width), (10, full_rounds),
                      (10, partial_rounds), (30, (1 << 30) - 1)])
    out = []
    while len(out) < width * (full_rounds + partial_rounds):
        v = 0
        for _ in range(field.bits):
            v = (v << 1) | int(lfsr.next_bit())
        if v < field.p:
            out.append(v)
    return out


def generate_mds(field: Field, t: int):
    return [[field.inv((x + y) % field.p) for y in range(t, 2 * t)] for x in range(t)]


def permute(field: Field, state, round_keys, mds, full_rounds: int, partial_rounds: int):
    p = field.p
    t = len(state)
    half = full_rounds // 2
    state = list(state)
    for rnd in range(full_rounds + partial_rounds):
        keys = round_keys[rnd * t:(rnd + 1) * t]
        state = [(s + k) % p for s, k in zip(state, keys)]
        if rnd < half or rnd >= half + partial_rounds:
            state = [pow(s, 5, p) for s in state]
        else:
            state[0] = pow(state[0], 5, p)
        state = [sum(mds[i][j] * state[j] for j in range(t)) % p for i in range(t)]
    return state
