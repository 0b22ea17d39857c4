"""Radix-2 evaluation domains (oracle; test infrastructure only).

Restates the part of ``ark_poly::{EvaluationDomain, Radix2EvaluationDomain}`` (ark-poly 0.3.0,
``src/domain/radix2/{mod,fft}.rs``; reached via ``pub use poly`` in
``/root/reference/plugins/arkworks/src/lib.rs:70-71``) that ``ark_groth16``'s
``R1CStoQAP::witness_map`` uses: ``new``, ``fft/ifft(_in_place)``, ``coset_fft/coset_ifft``,
``divide_by_vanishing_poly_on_coset_in_place``.  Conventions: natural order in and out,
``X[k] = sum_j x[j] w^(jk)``, ``w = TWO_ADIC_ROOT^(2^(TWO_ADICITY - log n))``, the inverse
multiplies by ``size_inv``, cosets use ``g = F::multiplicative_generator()``, inputs shorter
than the domain are zero-extended.  PARITY UNPINNED by the reference's tests; validated
against the O(n^2) DFT and round trips.
"""
from __future__ import annotations

from .fields import Field


class Radix2Domain:
    def __init__(self, field: Field, num_coeffs: int):
        """``Radix2EvaluationDomain::new``: size = next power of two >= num_coeffs."""
        size = 1
        log = 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > field.two_adicity:
            raise ValueError("ark returns None: log_size_of_group > TWO_ADICITY")
        self.f = field
        self.size = size
        self.log_size = log
        self.group_gen = field.root_of_unity(log)
        self.group_gen_inv = field.inv(self.group_gen)
        self.size_inv = field.inv(size % field.p)
        self.generator_inv = field.inv(field.generator)

    # -- core transform ------------------------------------------------------------------
    def _pad(self, x):
        x = list(x)
        if len(x) > self.size:
            raise ValueError("input longer than the domain")
        return x + [0] * (self.size - len(x))

    def _transform(self, x, w):
        p = self.f.p
        n = self.size
        a = list(x)
        # bit reversal then decimation-in-time butterflies
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wl = pow(w, n // length, p)
            half = length >> 1
            for s in range(0, n, length):
                t = 1
                for k in range(half):
                    u = a[s + k]
                    v = (a[s + k + half] * t) % p
                    a[s + k] = (u + v) % p
                    a[s + k + half] = (u - v) % p
                    t = (t * wl) % p
            length <<= 1
        return a

    def dft_naive(self, x, inverse=False):
        p = self.f.p
        n = self.size
        x = self._pad(x)
        w = self.group_gen_inv if inverse else self.group_gen
        out = []
        for k in range(n):
            wk = pow(w, k, p)
            acc, t = 0, 1
            for j in range(n):
                acc = (acc + x[j] * t) % p
                t = (t * wk) % p
            out.append((acc * self.size_inv) % p if inverse else acc)
        return out

    def fft(self, x):
        return self._transform(self._pad(x), self.group_gen)

    def ifft(self, x):
        p = self.f.p
        return [(v * self.size_inv) % p for v in self._transform(self._pad(x), self.group_gen_inv)]

    @staticmethod
    def distribute_powers(x, g, p):
        out, t = [], 1
        for v in x:
            out.append((v * t) % p)
            t = (t * g) % p
        return out

    def coset_fft(self, x):
        return self.fft(self.distribute_powers(self._pad(x), self.f.generator, self.f.p))

    def coset_ifft(self, x):
        return self.distribute_powers(self.ifft(x), self.generator_inv, self.f.p)

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self.size, self.f.p) - 1) % self.f.p

    def divide_by_vanishing_poly_on_coset(self, evals):
        p = self.f.p
        i = self.f.inv(self.evaluate_vanishing_polynomial(self.f.generator))
        return [(v * i) % p for v in evals]

    def element(self, i):
        return pow(self.group_gen, i, self.f.p)
