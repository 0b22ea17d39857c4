"""Shared helpers for the test-suite (seeded inputs, limb conversions)."""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


def int_to_limbs(v: int, n: int = 4):
    return [(v >> (64 * i)) & MASK64 for i in range(n)]


def limbs_to_int(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def ints_to_array(vals, n: int = 4) -> np.ndarray:
    return np.array([int_to_limbs(v, n) for v in vals], dtype=np.uint64).reshape(len(vals), n)


def random_scalars(n: int, modulus: int, seed: int) -> np.ndarray:
    """Uniform in [0, modulus) by mask-and-reject (how ark's UniformRand samples), seeded."""
    rng = np.random.default_rng(seed)
    bits = modulus.bit_length()
    top_mask = np.uint64((1 << (bits - 192)) - 1)
    mod_limbs = np.array(int_to_limbs(modulus), dtype=np.uint64)
    out = np.zeros((n, 4), dtype=np.uint64)
    todo = np.arange(n)
    while todo.size:
        cand = rng.integers(0, 1 << 64, size=(todo.size, 4), dtype=np.uint64)
        cand[:, 3] &= top_mask
        # lexicographic compare from the top limb: cand < modulus
        lt = np.zeros(todo.size, dtype=bool)
        eq = np.ones(todo.size, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (cand[:, k] < mod_limbs[k])
            eq &= cand[:, k] == mod_limbs[k]
        out[todo[lt]] = cand[lt]
        todo = todo[~lt]
    return out


def scalars_to_ints(arr: np.ndarray):
    return [limbs_to_int(row) for row in arr]


def verify_proof_with_pairings(pairing_name: str, fname: str, r1cs, trapdoor, public_inputs, proof_a, proof_b, proof_c) -> bool:
    """``ProofSystem::verify`` with real pairings (oracle/pairing.py) on a proof given in the C ABI's
    layout -- affine Montgomery limbs as ``ozl_groth16_prove`` writes them.  The verifying key is
    rebuilt from the known trapdoor exactly as ``Groth16.compile`` defines it:
    alpha G1, beta G2, gamma G2, delta G2, gamma_abc_j = (beta a_j + alpha b_j + c_j) / gamma G1."""
    from oracle import fields, groth16 as og, pairing as opair
    f = fields.FIELDS[fname]
    p = f.p
    P = opair.PAIRINGS[pairing_name]
    g1, g2 = P.g1, P.g2
    t = trapdoor
    a, b, c, _ = og.qap_at_tau(fname, r1cs, t.tau)
    ginv = f.inv(t.gamma)
    ic = [g1.mul_affine(g1.gen, (t.beta * a[j] + t.alpha * b[j] + c[j]) % p * ginv % p) for j in range(r1cs.n_instance)]
    to_pt = lambda g, limbs: g.affine_from_mont_limbs([int(v) for v in np.asarray(limbs, dtype=np.uint64).reshape(-1)])
    proof = (to_pt(g1, proof_a), to_pt(g2, proof_b), to_pt(g1, proof_c))
    if not (g1.is_on_curve(proof[0]) and g2.is_on_curve(proof[1]) and g1.is_on_curve(proof[2])):
        return False
    return opair.groth16_verify(pairing_name, g1.mul_affine(g1.gen, t.alpha), g2.mul_affine(g2.gen, t.beta),
                                g2.mul_affine(g2.gen, t.gamma), g2.mul_affine(g2.gen, t.delta), ic, list(public_inputs), proof)
