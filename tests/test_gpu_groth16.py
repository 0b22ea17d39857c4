"""-m gpu tests of the Groth16 prover row through the C ABI.

The reference never proves or verifies in its own tests, so parity is defined
against the oracle's restatement of ark_groth16::create_proof over a KNOWN trapdoor: the device's
proof points must equal [A']G1, [B']G2, [C']G1 bit-for-bit, where (A', B', C') are the discrete logs
a correct prover produces, and those must satisfy the verification equation in the exponent."""
import random

import numpy as np
import pytest

import openzl_b200 as ozl
from openzl_b200.circuits import PoseidonChain
from openzl_b200.groth16 import Groth16, Trapdoor, fixed_base_mul, fr_spmv, ints_to_limbs, limbs_to_ints
from oracle import cbind, curves, fields
from oracle import groth16 as og

pytestmark = pytest.mark.gpu
P = fields.BN254_FR.p


def _trapdoor(seed):
    rnd = random.Random(seed)
    return Trapdoor(*[rnd.randrange(2, P) for _ in range(5)])


def test_fixed_base_mul_matches_oracle(ctx):
    rnd = random.Random(1)
    ks = [0, 1, 2, P - 1] + [rnd.randrange(P) for _ in range(60)]
    for name in ("bn254_g1", "bn254_g2", "bls12_381_g1"):
        pts, inf = fixed_base_mul(ctx, ozl.CURVE_IDS[name], ints_to_limbs(ks))
        for i, k in enumerate(ks):
            exp, exp_inf = cbind.to_affine(name, cbind.gen_mul(name, k))
            assert bool((inf[i >> 3] >> (i & 7)) & 1) == exp_inf
            if not exp_inf:
                assert (pts[i] == exp).all()


def test_spmv_matches_python(ctx):
    ch = PoseidonChain(2)
    r1 = ch.r1cs()
    z = ch.assignment(11, 22)
    coef_m = ints_to_limbs(r1.coef_table, P, mont=True)
    z_m = ints_to_limbs(z, P, mont=True)
    for M in (r1.A, r1.B, r1.C):
        y = limbs_to_ints(fr_spmv(ctx, ozl.BN254_FR, M, coef_m, z_m), P, mont=True)
        assert y == r1.matvec(M, z)
        Mt = M.transpose(r1.n_vars)
        x = [random.Random(5).randrange(P) for _ in range(M.n_rows)]
        yt = limbs_to_ints(fr_spmv(ctx, ozl.BN254_FR, Mt, coef_m, ints_to_limbs(x, P, mont=True)), P, mont=True)
        exp = [0] * r1.n_vars
        for r in range(M.n_rows):
            for k in range(int(M.row_ptr[r]), int(M.row_ptr[r + 1])):
                exp[int(M.col_idx[k])] = (exp[int(M.col_idx[k])] + x[r] * r1.coef_table[int(M.coef_idx[k])]) % P
        assert yt == exp


@pytest.mark.parametrize("links", [1, 3])
def test_prove_matches_oracle_and_verifies(ctx, links):
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    z = ch.assignment(123456789, 987654321)
    assert r1.is_satisfied(z)
    td = _trapdoor(links)
    pk, vk = Groth16.compile(ctx, "bn254", r1, td, keep_queries=True)
    try:
        # setup scalars computed on the device (ifft route) == oracle (closed-form Lagrange route)
        a, b, c, n = og.qap_at_tau("bn254_fr", r1, td.tau)
        assert pk.queries["a"] == a and pk.queries["b"] == b and pk.queries["c"] == c and pk.domain_size == n
        rnd = random.Random(99)
        r, s = rnd.randrange(P), rnd.randrange(P)
        z_m = ints_to_limbs(z, P, mont=True)
        proof, h_m = Groth16.prove_with_randomness(pk, z_m, r, s, want_h=True)
        # witness map parity (7 NTTs + pointwise) against the oracle's ark-ordered restatement
        h = og.witness_map("bn254_fr", r1, z)
        assert limbs_to_ints(h_m, P, mont=True) == h
        assert h[n - 1] == 0          # deg h <= n - 2: truncating to the n-1 h_query bases loses nothing
        A, B, C = og.prove_exponents("bn254_fr", r1, z, td, r, s, h=h)
        assert og.verify_exponents("bn254_fr", r1, td, [z[1]], A, B, C)
        assert not og.verify_exponents("bn254_fr", r1, td, [(z[1] + 1) % P], A, B, C)
        for name, k, got in (("bn254_g1", A, proof.a), ("bn254_g2", B, proof.b), ("bn254_g1", C, proof.c)):
            exp, exp_inf = cbind.to_affine(name, cbind.gen_mul(name, k))
            assert not exp_inf and (got == exp).all(), name
        # ProofSystem::verify (groth16.rs:460-466) with real pairings on the device's proof points
        if links == 1:
            from tests.util import verify_proof_with_pairings
            assert verify_proof_with_pairings("bn254", "bn254_fr", r1, td, [z[1]], proof.a, proof.b, proof.c)
            assert not verify_proof_with_pairings("bn254", "bn254_fr", r1, td, [(z[1] + 1) % P], proof.a, proof.b, proof.c)
        # wire format (groth16.rs:98-107): 128 compressed bytes that decode back to the same three points
        from openzl_b200 import serialize as ser
        raw = proof.to_bytes("bn254")
        assert len(raw) == 128
        pa, pb, pc = ser.proof_from_bytes("bn254", raw)
        assert ser.is_on_curve(ser.BN254_G1, pa) and ser.is_on_curve(ser.BN254_G2, pb) and ser.is_on_curve(ser.BN254_G1, pc)
        assert (ser.point_to_limbs(ser.BN254_G1, pa) == proof.a).all()
        assert (ser.point_to_limbs(ser.BN254_G2, pb) == proof.b).all()
        assert (ser.point_to_limbs(ser.BN254_G1, pc) == proof.c).all()
        # r = s = 0 (the degenerate blinding ark special-cases) still matches
        proof0 = Groth16.prove_with_randomness(pk, z_m, 0, 0)
        A0, B0, C0 = og.prove_exponents("bn254_fr", r1, z, td, 0, 0, h=h)
        assert (proof0.a == cbind.to_affine("bn254_g1", cbind.gen_mul("bn254_g1", A0))[0]).all()
        assert (proof0.c == cbind.to_affine("bn254_g1", cbind.gen_mul("bn254_g1", C0))[0]).all()
    finally:
        pk.free()


def test_prove_bls12_381(ctx):
    """Same prover on the other pairing family (`Pairing` for BLS12-381, pairing.rs:9-38): Poseidon
    over BLS12-381 Fr with the width-3 / 8+55 rounds of the reference's tutorial KAT."""
    from openzl_b200.circuits import PoseidonParams
    pr = fields.BLS12_381_FR.p
    ch = PoseidonChain(1, PoseidonParams.generate(modulus=pr))
    # the chain's hash is the reference's permutation: KAT [3, 1, 2] from openzl-tutorials/src/poseidon.rs:388-401
    assert ch.params.permute([3, 1, 2])[0] == 1808609226548932412441401219270714120272118151392880709881321306315053574086
    r1 = ch.r1cs()
    z = ch.assignment(42, 43)
    assert r1.is_satisfied(z)
    rnd = random.Random(381)
    td = Trapdoor(*[rnd.randrange(2, pr) for _ in range(5)])
    pk, vk = Groth16.compile(ctx, "bls12_381", r1, td)
    try:
        r, s = rnd.randrange(pr), rnd.randrange(pr)
        proof, h_m = Groth16.prove_with_randomness(pk, ints_to_limbs(z, pr, mont=True), r, s, want_h=True)
        h = og.witness_map("bls12_381_fr", r1, z)
        assert limbs_to_ints(h_m, pr, mont=True) == h
        A, B, C = og.prove_exponents("bls12_381_fr", r1, z, td, r, s, h=h)
        assert og.verify_exponents("bls12_381_fr", r1, td, [z[1]], A, B, C)
        for name, k, got in (("bls12_381_g1", A, proof.a), ("bls12_381_g2", B, proof.b), ("bls12_381_g1", C, proof.c)):
            exp, exp_inf = cbind.to_affine(name, cbind.gen_mul(name, k))
            assert not exp_inf and (got == exp).all(), name
        # and the device's proof verifies with real pairings (oracle/pairing.py), like ProofSystem::verify
        from tests.util import verify_proof_with_pairings
        assert verify_proof_with_pairings("bls12_381", "bls12_381_fr", r1, td, [z[1]], proof.a, proof.b, proof.c)
    finally:
        pk.free()


def test_prove_rejects_bad_shapes(ctx):
    ch = PoseidonChain(1)
    r1 = ch.r1cs()
    pk, _ = Groth16.compile(ctx, "bn254", r1, _trapdoor(7))
    try:
        with pytest.raises(ozl.OzlError):
            Groth16.prove_with_randomness(pk, np.zeros((3, 4), dtype=np.uint64), 1, 1)
    finally:
        pk.free()


def test_prove_2p16_constraints_matches_oracle_and_verifies(ctx):
    """188 Poseidon links = 65,424 constraints (domain 2^16): the proof of the device prover equals
    [A']G1, [B']G2, [C']G1 from the oracle's restatement bit-for-bit, h equals the oracle's witness map
    computed with the C++ oracle NTT, and the product's host `Groth16.verify` accepts it."""
    links = 188
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    assert r1.n_constraints == 65424
    z = ch.assignment(2026, 1017)
    td = _trapdoor(16)
    pk, vk = Groth16.compile(ctx, "bn254", r1, td, keep_queries=True, precompute=32)
    try:
        n = pk.domain_size
        assert n == 1 << 16
        rnd = random.Random(4)
        r, s = rnd.randrange(P), rnd.randrange(P)
        z_m = ints_to_limbs(z, P, mont=True)
        proof, h_m = Groth16.prove_with_randomness(pk, z_m, r, s, want_h=True)
        # witness map through the oracle's C++ NTT, ark's step order (oracle/groth16.py:witness_map at speed)
        nc, ni = r1.n_constraints, r1.n_instance
        rows = [r1.matvec(M, z) + [0] * (n - nc) for M in (r1.A, r1.B, r1.C)]
        for j in range(ni):
            rows[0][nc + j] = z[j]
        ev = []
        for v in rows:
            v_m = ints_to_limbs(v, P, mont=True)
            v_m = cbind.ntt("bn254_fr", v_m, inverse=True)
            ev.append(limbs_to_ints(cbind.ntt("bn254_fr", v_m, coset=True), P, mont=True))
        zinv = pow((pow(5, n, P) - 1) % P, -1, P)                  # 1 / Z(g) on the coset, g = 5
        q = [(a * b - c) * zinv % P for a, b, c in zip(*ev)]
        h = limbs_to_ints(cbind.ntt("bn254_fr", ints_to_limbs(q, P, mont=True), inverse=True, coset=True), P, mont=True)
        assert limbs_to_ints(h_m, P, mont=True) == h and h[n - 1] == 0
        # proof exponents from the setup scalars the device computed (checked against the oracle at small sizes)
        qa, qb, qc = pk.queries["a"], pk.queries["b"], pk.queries["c"]
        m = r1.n_vars
        dinv = pow(td.delta, -1, P)
        zt = (pow(td.tau, n, P) - 1) % P
        za = sum(z[j] * qa[j] for j in range(m)) % P
        zb = sum(z[j] * qb[j] for j in range(m)) % P
        A = (td.alpha + za + r * td.delta) % P
        B = (td.beta + zb + s * td.delta) % P
        l_acc = sum(z[j] * (td.beta * qa[j] + td.alpha * qb[j] + qc[j]) for j in range(ni, m)) % P * dinv % P
        ht, tp = 0, 1
        for i in range(n - 1):
            ht = (ht + h[i] * tp) % P
            tp = tp * td.tau % P
        C = (l_acc + ht * zt % P * dinv + s * A + r * B - r * s % P * td.delta) % P
        for name, k, got in (("bn254_g1", A, proof.a), ("bn254_g2", B, proof.b), ("bn254_g1", C, proof.c)):
            exp, exp_inf = cbind.to_affine(name, cbind.gen_mul(name, k))
            assert not exp_inf and (got == exp).all(), name
        assert Groth16.verify(vk, [z[1]], proof)
        assert not Groth16.verify(vk, [(z[1] + 1) % P], proof)
    finally:
        pk.free()


def test_prove_baseline_size_2p20_constraints_verifies(ctx):
    """BASELINE config 4 itself: 3013 Poseidon links = 1,048,524 constraints (domain 2^20).  The proof of the device
    prover passes the pairing equation in the product's host verifier, a wrong public input is rejected, the proof
    is reproducible (two calls, same randomness, same bytes -- the prover's four streams race on nothing), and A and
    B equal [alpha + <z, a> + r delta]G1 / [beta + <z, b> + s delta]G2 computed from the setup scalars."""
    links = 3013
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    assert r1.n_constraints == 1048524
    z = ch.assignment(99, 1234)
    td = _trapdoor(20)
    pk, vk = Groth16.compile(ctx, "bn254", r1, td, keep_queries=True, precompute=32)
    try:
        assert pk.domain_size == 1 << 20
        rnd = random.Random(20)
        r, s = rnd.randrange(P), rnd.randrange(P)
        z_m = ints_to_limbs(z, P, mont=True)
        proof = Groth16.prove_with_randomness(pk, z_m, r, s)
        again = Groth16.prove_with_randomness(pk, z_m, r, s)
        assert (proof.a == again.a).all() and (proof.b == again.b).all() and (proof.c == again.c).all()
        qa, qb = pk.queries["a"], pk.queries["b"]
        m = r1.n_vars
        A = (td.alpha + sum(z[j] * qa[j] for j in range(m)) + r * td.delta) % P
        B = (td.beta + sum(z[j] * qb[j] for j in range(m)) + s * td.delta) % P
        for name, k, got in (("bn254_g1", A, proof.a), ("bn254_g2", B, proof.b)):
            exp, exp_inf = cbind.to_affine(name, cbind.gen_mul(name, k))
            assert not exp_inf and (got == exp).all(), name
        assert Groth16.verify(vk, [z[1]], proof)
        assert not Groth16.verify(vk, [(z[1] + 1) % P], proof)
    finally:
        pk.free()


def test_proving_key_bytes_to_device_round_trip(ctx):
    """Row f-3 end to end (`ProvingContext::{encode, decode}`, groth16.rs:142-179): the device key is
    encoded as ark's unchecked uncompressed ProvingKey bytes, decoded, re-uploaded through
    `ozl_msm_bases_upload` (with the infinity bitsets of the b queries) and proves to the SAME proof."""
    ch = PoseidonChain(2)
    r1 = ch.r1cs()
    z = ch.assignment(77, 88)
    td = _trapdoor(33)
    pk, vk = Groth16.compile(ctx, "bn254", r1, td)
    pk2 = None
    try:
        raw = pk.encode()
        from openzl_b200 import serialize as ser
        g1, g2 = ser.PAIRING_GROUPS["bn254"]
        m, n = r1.n_vars, pk.domain_size
        s1, s2 = g1.uncompressed_size, g2.uncompressed_size
        vk_len = s1 + 3 * s2 + 8 + r1.n_instance * s1
        assert len(raw) == vk_len + 2 * s1 + 2 * (8 + m * s1) + (8 + m * s2) + (8 + (n - 1) * s1) + (8 + (m - r1.n_instance) * s1)
        pk2, vk2 = Groth16.proving_context_from_bytes(ctx, "bn254", raw, r1, precompute=4)
        assert pk2.encode() == raw
        rnd = random.Random(8)
        r, s = rnd.randrange(P), rnd.randrange(P)
        z_m = ints_to_limbs(z, P, mont=True)
        p1 = Groth16.prove_with_randomness(pk, z_m, r, s)
        p2 = Groth16.prove_with_randomness(pk2, z_m, r, s)
        assert (p1.a == p2.a).all() and (p1.b == p2.b).all() and (p1.c == p2.c).all()
        assert p1.to_bytes("bn254") == p2.to_bytes("bn254")
        assert Groth16.verify(vk2, [z[1]], p2) and Groth16.verify(vk, [z[1]], p2)
        assert (vk2.gamma_abc_g1 == vk.gamma_abc_g1).all()
        with pytest.raises(ozl.OzlError):          # a key for another circuit is refused
            Groth16.proving_context_from_bytes(ctx, "bn254", raw, PoseidonChain(1).r1cs())
        # random blinding from the OS CSPRNG still verifies (ProofSystem::prove)
        assert Groth16.verify(vk, [z[1]], Groth16.prove(pk, z_m))
    finally:
        pk.free()
        if pk2 is not None:
            pk2.free()


def test_pk_belongs_to_its_context_and_is_released_with_it():
    """The pk registry lives in the context: a handle is unknown to another context, and destroying the
    context releases proving keys the caller forgot."""
    c1, c2 = ozl.Context(0), ozl.Context(0)
    try:
        r1 = PoseidonChain(1).r1cs()
        pk, _ = Groth16.compile(c1, "bn254", r1, _trapdoor(5))
        size = np.zeros(1, dtype=np.uint32)
        assert c1._lib.ozl_groth16_domain_size(c1._h, pk.handle, size.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint32))) == 0
        assert c2._lib.ozl_groth16_domain_size(c2._h, pk.handle, size.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint32))) == 5
        assert c2._lib.ozl_groth16_pk_destroy(c2._h, pk.handle) == 5
    finally:
        c1.close()          # pk not freed explicitly: ozl_ctx_destroy must release it without error
        c2.close()
