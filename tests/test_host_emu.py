"""Limb-level check of the DEVICE arithmetic headers (openzl_b200/csrc/{fp,ec}.cuh) on the CPU:
the headers are compiled with g++ and the PTX carry-chain primitives emulated
(tests/host_emu/emu.cpp, -DOZL_HOST_EMU), then compared with the big-int oracle.  This catches
limb-indexing / carry mistakes without a GPU; the same code paths are re-checked on the GPU by
the -m gpu tests.  The emulation is test-only and is never linked into libozl_b200.so.
"""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import curves, fields

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "emu.cpp")
SO = os.path.join(HERE, "host_emu", "libemu.so")


@pytest.fixture(scope="module")
def emu():
    hdrs = [os.path.join(HERE, "..", "openzl_b200", "csrc", h) for h in ("fp.cuh", "ec.cuh", "ptx.cuh", "params_gen.cuh")]
    newest = max(os.path.getmtime(p) for p in hdrs + [SRC])
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-DOZL_HOST_EMU", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def l32(v, n):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def f32(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


FIELD_IDS = [(0, fields.BLS12_381_FQ), (1, fields.BLS12_381_FR), (2, fields.BN254_FQ), (3, fields.BN254_FR)]


def edge_values(f):
    n = f.limbs64 * 2
    vals = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.R2, (f.p - 1) // 2, (f.p + 1) // 2]
    vals += [((1 << (32 * k)) - 1) % f.p for k in range(1, n + 1)]
    vals += [(1 << (32 * k)) % f.p for k in range(1, n)]
    vals += [int("ffffffff00000000" * f.limbs64, 16) % f.p, int("00000000ffffffff" * f.limbs64, 16) % f.p]
    return vals


@pytest.mark.parametrize("fid,f", FIELD_IDS)
def test_field_ops_match_oracle(emu, fid, f):
    n = f.limbs64 * 2
    rnd = random.Random(fid)
    vals = edge_values(f)
    pairs = [(a, b) for a in vals for b in vals] + [(rnd.randrange(f.p), rnd.randrange(f.p)) for _ in range(3000)]
    for a, b in pairs:
        for op, exp in ((0, f.mont_mul(a, b)), (1, (a + b) % f.p), (2, (a - b) % f.p), (3, (-a) % f.p),
                        (5, f.mont_mul(a, a)), (6, (2 * a) % f.p), (7, f.mont_mul(a, a)), (8, f.mont_mul(a, b)),
                        (9, (f.mont_mul(a, b) - f.mont_mul(b, (a + b) % f.p)) % f.p),
                        (10, (f.mont_mul(a, a) + f.mont_mul(b, b)) % f.p)):
            out = np.zeros(n, dtype=np.uint32)
            emu.emu_fp_op(fid, op, P(l32(a, n)), P(l32(b, n)), P(out))
            assert f32(out) == exp, (f.name, op, hex(a), hex(b))
    for a in [1, 2, f.p - 1, rnd.randrange(1, f.p)]:
        out = np.zeros(n, dtype=np.uint32)
        emu.emu_fp_op(fid, 4, P(l32(f.to_mont(a), n)), P(l32(0, n)), P(out))
        assert f.from_mont(f32(out)) == f.inv(a)


@pytest.mark.parametrize("fid,f", [(4, fields.BLS12_381_FQ), (5, fields.BN254_FQ)])
def test_fp2_ops_match_oracle(emu, fid, f):
    F2 = curves.Fp2Ops(f)
    n = f.limbs64 * 2
    rnd = random.Random(fid)

    def enc(a):
        return np.concatenate([l32(f.to_mont(a[0]), n), l32(f.to_mont(a[1]), n)])

    def dec(o):
        return (f.from_mont(f32(o[:n])), f.from_mont(f32(o[n:])))

    for _ in range(200):
        a = (rnd.randrange(f.p), rnd.randrange(f.p))
        b = (rnd.randrange(f.p), rnd.randrange(f.p))
        for op, exp in ((0, F2.mul(a, b)), (1, F2.add(a, b)), (2, F2.sub(a, b)), (3, F2.neg(a)), (5, F2.mul(a, a)), (4, F2.inv(a))):
            out = np.zeros(2 * n, dtype=np.uint32)
            emu.emu_fp_op(fid, op, P(enc(a)), P(enc(b)), P(out))
            assert dec(out) == exp, (op,)


@pytest.mark.parametrize("fid,f", FIELD_IDS)
def test_wide_product_and_cios_reduction(emu, fid, f):
    """The building blocks of the lazily reduced Fq2 arithmetic: `Fp::mul_wide` (unreduced product of ARBITRARY
    N-limb integers through two register-aligned accumulators) equals the integer product, `Fp::redc_cios`
    (stand-alone Montgomery reduction, CIOS rows with the high limbs injected one per row) equals t / R mod p for
    any t < p R, and their composition is the Montgomery product."""
    n = f.limbs64 * 2
    rnd = random.Random(100 + fid)
    top = (1 << (32 * n)) - 1
    R = 1 << (32 * n)
    ints = [0, 1, top, top - 1, f.p, f.p - 1, 2 * f.p - 2, (1 << (32 * n - 1)), int("ffffffff00000000" * f.limbs64, 16),
            int("00000000ffffffff" * f.limbs64, 16)] + [rnd.randrange(R) for _ in range(40)]
    for a in ints:
        for b in ints:
            out = np.zeros(2 * n, dtype=np.uint32)
            emu.emu_wide_op(fid, 0, P(l32(a, n)), P(l32(b, n)), P(out))
            assert f32(out) == a * b, (f.name, hex(a), hex(b))
    rinv = pow(R, -1, f.p)
    ts = [0, 1, f.p * R - 1, f.p * R - f.p, f.p * f.p, 2 * f.p * f.p, R, R - 1, (f.p - 1) * (f.p - 1)] + [rnd.randrange(f.p * R) for _ in range(3000)]
    ts = [t for t in ts if t < f.p * R]
    for t in ts:
        out = np.zeros(n, dtype=np.uint32)
        emu.emu_wide_op(fid, 1, P(l32(t, 2 * n)), P(l32(0, n)), P(out))
        assert f32(out) == t * rinv % f.p, (f.name, hex(t))
    for a, b in [(a, b) for a in edge_values(f) for b in edge_values(f)] + [(rnd.randrange(f.p), rnd.randrange(f.p)) for _ in range(500)]:
        out = np.zeros(n, dtype=np.uint32)
        emu.emu_fp_op(fid, 11, P(l32(a, n)), P(l32(b, n)), P(out))
        assert f32(out) == f.mont_mul(a, b)


@pytest.mark.parametrize("fid,f", [(4, fields.BLS12_381_FQ), (5, fields.BN254_FQ)])
def test_fp2_lazy_products_match_oracle(emu, fid, f):
    """`Fp2::mul_lazy` (three unreduced products, two reductions) and `Fp2::mul_sub2_lazy` (a b - c d, six products, two
    reductions) against the oracle's Fq2, on random operands and on every combination of extreme components."""
    F2 = curves.Fp2Ops(f)
    n = f.limbs64 * 2
    rnd = random.Random(200 + fid)

    def enc(a):
        return np.concatenate([l32(f.to_mont(a[0]), n), l32(f.to_mont(a[1]), n)])

    def dec(o):
        return (f.from_mont(f32(o[:n])), f.from_mont(f32(o[n:])))

    # components whose MONTGOMERY form is extreme (0, 1, p - 1, p - 2, R mod p) as well as random ones
    ext = [f.from_mont(v) for v in (0, 1, f.p - 1, f.p - 2, f.R, (f.p - 1) // 2)]
    cases = [((a0, a1), (b0, b1)) for a0 in ext for a1 in ext for b0 in ext[:4] for b1 in ext[:4]]
    cases += [((rnd.randrange(f.p), rnd.randrange(f.p)), (rnd.randrange(f.p), rnd.randrange(f.p))) for _ in range(1500)]
    for a, b in cases:
        out = np.zeros(2 * n, dtype=np.uint32)
        emu.emu_fp_op(fid, 11, P(enc(a)), P(enc(b)), P(out))
        assert dec(out) == F2.mul(a, b), ("mul_lazy", a, b)
        out = np.zeros(2 * n, dtype=np.uint32)
        emu.emu_fp_op(fid, 12, P(enc(a)), P(enc(b)), P(out))
        assert dec(out) == F2.sub(F2.mul(a, b), F2.mul(b, F2.add(a, b))), ("mul_sub2_lazy", a, b)


CURVE_IDS = [(0, "bls12_381_g1"), (1, "bls12_381_g2"), (2, "bn254_g1"), (3, "bn254_g2")]


@pytest.mark.parametrize("cid,name", CURVE_IDS)
def test_xyzz_group_ops_match_oracle(emu, cid, name):
    c = curves.CURVES[name]
    f = c.base
    n32 = f.limbs64 * 2 * c.F.degree
    rnd = random.Random(cid)

    def enc_f(e):
        return np.concatenate([l32(f.to_mont(x), f.limbs64 * 2) for x in c.F.coords(e)])

    def enc_aff(Pt):
        return np.concatenate([enc_f(Pt[0]), enc_f(Pt[1])])

    def enc_xyzz_from_aff(Pt, z=None):
        # non-trivial representative: (x z^2, y z^3, z^2, z^3)
        F = c.F
        if Pt is None:
            return np.concatenate([enc_f(F.zero), enc_f(F.one), enc_f(F.zero), enc_f(F.zero)])
        z = z if z is not None else F.one
        zz = F.sqr(z)
        zzz = F.mul(zz, z)
        return np.concatenate([enc_f(F.mul(Pt[0], zz)), enc_f(F.mul(Pt[1], zzz)), enc_f(zz), enc_f(zzz)])

    def to_affine(xyzz):
        out = np.zeros(2 * n32, dtype=np.uint32)
        emu.emu_ec_op(cid, 4, P(xyzz), P(xyzz), 0, P(out))
        if not out.any():
            return None
        per = f.limbs64 * 2
        vals = [f.from_mont(f32(out[i * per:(i + 1) * per])) for i in range(2 * c.F.degree)]
        d = c.F.degree
        return (c.F.from_coords(vals[:d]), c.F.from_coords(vals[d:]))

    def rand_z():
        return c.F.from_coords([rnd.randrange(1, f.p) for _ in range(c.F.degree)])

    Pt = c.mul_affine(c.gen, rnd.randrange(1, c.fr.p))
    Qt = c.mul_affine(c.gen, rnd.randrange(1, c.fr.p))
    cases = [(Pt, Qt), (Pt, Pt), (Pt, c.neg(Pt)), (None, Qt)]
    for A, B in cases:
        acc = enc_xyzz_from_aff(A, rand_z())
        out = np.zeros(4 * n32, dtype=np.uint32)
        emu.emu_ec_op(cid, 0, P(acc), P(enc_aff(B)), 0, P(out))          # mixed add
        assert to_affine(out) == c.add_affine(A, B)
        outc = np.zeros(4 * n32, dtype=np.uint32)
        emu.emu_ec_op(cid, 6, P(acc), P(enc_aff(B)), 0, P(outc))         # cold-path mixed add (paired products)
        assert to_affine(outc) == c.add_affine(A, B)
        out2 = np.zeros(4 * n32, dtype=np.uint32)
        emu.emu_ec_op(cid, 1, P(acc), P(enc_xyzz_from_aff(B, rand_z())), 0, P(out2))   # full add
        assert to_affine(out2) == c.add_affine(A, B)
    # add with identity on the right, doubling, small scalar multiple, Jacobian export
    acc = enc_xyzz_from_aff(Pt, rand_z())
    out = np.zeros(4 * n32, dtype=np.uint32)
    emu.emu_ec_op(cid, 1, P(acc), P(enc_xyzz_from_aff(None)), 0, P(out))
    assert to_affine(out) == Pt
    emu.emu_ec_op(cid, 2, P(acc), P(acc), 0, P(out))
    assert to_affine(out) == c.add_affine(Pt, Pt)
    for k in (0, 1, 2, 3, 77, 65537, 0x7FFFF):
        emu.emu_ec_op(cid, 3, P(acc), P(acc), k, P(out))
        assert to_affine(out) == c.mul_affine(Pt, k)
    jac = np.zeros(3 * n32, dtype=np.uint32)
    emu.emu_ec_op(cid, 5, P(acc), P(acc), 0, P(jac))
    per = f.limbs64 * 2
    vals = [f.from_mont(f32(jac[i * per:(i + 1) * per])) for i in range(3 * c.F.degree)]
    d = c.F.degree
    J = tuple(c.F.from_coords(vals[i * d:(i + 1) * d]) for i in range(3))
    assert c.to_affine(J) == Pt


def test_fp64_montgomery_product_matches_oracle(emu):
    """The double-precision Montgomery product (openzl_b200/csrc/fp64mul.cuh, 8 x 48-bit limbs) equals the
    integer one on edge values and random inputs; the g++ build emulates the FP64 primitives exactly with
    128-bit integers and aborts on any step that would round, so passing also proves the exactness argument."""
    f = fields.BLS12_381_FQ
    n = 12
    # the 48-bit constants of the header
    limbs = [(f.p >> (48 * k)) & ((1 << 48) - 1) for k in range(8)]
    assert limbs == [281474976688811, 194974335351294, 270634993844222, 113459389855408, 83034393350847, 73992301405303,
                     253550359455670, 28591897852287]
    assert (-pow(f.p, -1, 1 << 48)) % (1 << 48) == 281462091612157
    rnd = random.Random(48)
    vals = edge_values(f)
    pairs = [(a, b) for a in vals for b in vals] + [(rnd.randrange(f.p), rnd.randrange(f.p)) for _ in range(20000)]
    for a, b in pairs:
        out = np.zeros(n, dtype=np.uint32)
        emu.emu_mul_fp64(P(l32(a, n)), P(l32(b, n)), P(out))
        assert f32(out) == f.mont_mul(a, b), (hex(a), hex(b))
