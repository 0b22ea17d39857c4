"""Groth16 prover restatement "in the exponent" (oracle; test infrastructure only).

Restates ``ark_groth16::{create_proof, R1CStoQAP::witness_map}`` (ark-groth16 0.3.0, called at
/root/reference/plugins/arkworks/src/groth16.rs:454) with the step order of SURVEY.md section 3.1,
over a KNOWN trapdoor so that every proof element has a known discrete log:

    A = [alpha + <z, a(tau)> + r delta] G1        B = [beta + <z, b(tau)> + s delta] G2
    C = [ (<z_aux, k_aux> + h(tau) Z(tau)) / delta + s A' + r B' - r s delta ] G1

and the verification equation  e(A,B) = e(alpha,beta) e(sum x_j IC_j, gamma) e(C, delta)  becomes
A' B' = alpha beta + gamma * sum x_j ic_j + C' delta  in Fr.  PARITY UNPINNED by the reference (no
reference test proves or verifies); pinned by this algebraic identity and by bit-exact comparison
of the device's proof points with [A']G1, [B']G2, [C']G1 from the CPU group law.
"""
from __future__ import annotations

from typing import List

from .fields import FIELDS
from .ntt import Radix2Domain


def _matvec(rows, coef, z, p, n_rows):
    row_ptr, col, cidx = rows
    out = []
    for r in range(n_rows):
        acc = 0
        for k in range(int(row_ptr[r]), int(row_ptr[r + 1])):
            acc += coef[int(cidx[k])] * z[int(col[k])]
        out.append(acc % p)
    return out


def witness_map(field_name: str, r1cs, z: List[int]) -> List[int]:
    """h coefficients (length = domain size) exactly as ark's witness_map orders the steps."""
    f = FIELDS[field_name]
    p = f.p
    nc, ni = r1cs.n_constraints, r1cs.n_instance
    d = Radix2Domain(f, nc + ni)
    n = d.size
    csr = lambda M: (M.row_ptr, M.col_idx, M.coef_idx)
    a = _matvec(csr(r1cs.A), r1cs.coef_table, z, p, nc) + [0] * (n - nc)
    b = _matvec(csr(r1cs.B), r1cs.coef_table, z, p, nc) + [0] * (n - nc)
    c = _matvec(csr(r1cs.C), r1cs.coef_table, z, p, nc) + [0] * (n - nc)
    for j in range(ni):
        a[nc + j] = z[j]
    a, b, c = d.ifft(a), d.ifft(b), d.ifft(c)
    a, b, c = d.coset_fft(a), d.coset_fft(b), d.coset_fft(c)
    ab = [(x * y - w) % p for x, y, w in zip(a, b, c)]
    ab = d.divide_by_vanishing_poly_on_coset(ab)
    return d.coset_ifft(ab)


def lagrange_at(field_name: str, n: int, tau: int) -> List[int]:
    """L_i(tau) = Z(tau) w^i / (n (tau - w^i))  -- independent of the device's ifft route."""
    f = FIELDS[field_name]
    p = f.p
    d = Radix2Domain(f, n)
    zt = (pow(tau, n, p) - 1) % p
    ninv = f.inv(n % p)
    out, w = [], 1
    for _ in range(n):
        out.append(zt * w % p * ninv % p * f.inv((tau - w) % p) % p)
        w = (w * d.group_gen) % p
    return out


def qap_at_tau(field_name: str, r1cs, tau: int):
    """(a_j(tau), b_j(tau), c_j(tau)) for every variable j, with ark's instance rows folded into a."""
    f = FIELDS[field_name]
    p = f.p
    nc, ni, m = r1cs.n_constraints, r1cs.n_instance, r1cs.n_vars
    n = Radix2Domain(f, nc + ni).size
    L = lagrange_at(field_name, n, tau)
    out = []
    for M in (r1cs.A, r1cs.B, r1cs.C):
        acc = [0] * m
        for r in range(nc):
            for k in range(int(M.row_ptr[r]), int(M.row_ptr[r + 1])):
                j = int(M.col_idx[k])
                acc[j] = (acc[j] + L[r] * r1cs.coef_table[int(M.coef_idx[k])]) % p
        out.append(acc)
    a, b, c = out
    for j in range(ni):
        a[j] = (a[j] + L[nc + j]) % p
    return a, b, c, n


def prove_exponents(field_name: str, r1cs, z: List[int], trapdoor, r: int, s: int, h: List[int] | None = None):
    """Discrete logs (A', B', C') of the proof a correct prover outputs for (z, r, s)."""
    f = FIELDS[field_name]
    p = f.p
    t = trapdoor
    ni, m = r1cs.n_instance, r1cs.n_vars
    a, b, c, n = qap_at_tau(field_name, r1cs, t.tau)
    if h is None:
        h = witness_map(field_name, r1cs, z)
    dinv = f.inv(t.delta)
    zt = (pow(t.tau, n, p) - 1) % p
    za = sum(z[j] * a[j] for j in range(m)) % p
    zb = sum(z[j] * b[j] for j in range(m)) % p
    A = (t.alpha + za + r * t.delta) % p
    B = (t.beta + zb + s * t.delta) % p
    l_acc = sum(z[j] * (t.beta * a[j] + t.alpha * b[j] + c[j]) for j in range(ni, m)) % p * dinv % p
    h_acc = sum(h[i] * pow(t.tau, i, p) for i in range(n - 1)) % p * zt % p * dinv % p
    C = (l_acc + h_acc + s * A + r * B - r * s % p * t.delta) % p
    return A, B, C


def verify_exponents(field_name: str, r1cs, trapdoor, public_inputs: List[int], A: int, B: int, C: int) -> bool:
    """The Groth16 pairing equation with every element replaced by its discrete log."""
    f = FIELDS[field_name]
    p = f.p
    t = trapdoor
    a, b, c, _ = qap_at_tau(field_name, r1cs, t.tau)
    ginv = f.inv(t.gamma)
    x = [1] + list(public_inputs)
    assert len(x) == r1cs.n_instance
    ic = sum(x[j] * ((t.beta * a[j] + t.alpha * b[j] + c[j]) % p) for j in range(r1cs.n_instance)) % p * ginv % p
    return (A * B - (t.alpha * t.beta + ic * t.gamma + C * t.delta)) % p == 0
