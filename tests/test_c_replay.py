"""The Rust shim's FFI call sequence replayed from plain C with PAGEABLE host buffers
(examples/ozl_groth16_replay.c): compile = five `ozl_msm_bases_upload` + precompute +
`ozl_groth16_pk_create`, prove = `ozl_groth16_prove`, plus one `ozl_msm` with pageable scalars.  No Rust
toolchain exists in the image, so this is how the path `plugins/b200/src/groth16.rs` takes is exercised
without Python, torch or pinned memory between the caller and libozl_b200.so.  The proof must equal the
one the in-process (ctypes) path produces from the same key, witness and blinding scalars."""
import os
import random
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "openzl_b200")


def _build(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "ozl_groth16_replay")
    subprocess.check_call([cc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ozl_groth16_replay.c"), "-L", LIBDIR, "-lozl_b200",
                           f"-Wl,-rpath,{LIBDIR}", "-o", exe])
    return exe


def test_replay_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe = _build(tmp_path)
    r = subprocess.run([exe, "missing.blob", str(tmp_path / "out")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


@pytest.mark.gpu
@pytest.mark.parametrize("links,precompute", [(2, 4), (12, 32)])
def test_replay_matches_in_process_path(tmp_path, ctx, links, precompute):
    import openzl_b200 as ozl
    from openzl_b200.circuits import PoseidonChain
    from openzl_b200.context import Bases
    from openzl_b200.groth16 import Groth16, PAIRINGS, Trapdoor, ints_to_limbs
    P = PAIRINGS["bn254"]["r"]
    ch = PoseidonChain(links)
    r1 = ch.r1cs()
    z = ch.assignment(31337, 271828)
    rnd = random.Random(links)
    td = Trapdoor(*[rnd.randrange(2, P) for _ in range(5)])
    pk, vk = Groth16.compile(ctx, "bn254", r1, td, precompute=precompute)
    try:
        r, s = rnd.randrange(P), rnd.randrange(P)
        z_m, z_c = ints_to_limbs(z, P, mont=True), ints_to_limbs(z)
        proof = Groth16.prove_with_randomness(pk, z_m, r, s)
        blob = [np.array([0x4f5a4c5245504c59, 0, r1.n_constraints, r1.n_instance, r1.n_vars, len(r1.coef_table), pk.domain_size,
                          precompute], dtype=np.uint64).tobytes()]
        for M in (r1.A, r1.B, r1.C):
            blob += [np.ascontiguousarray(M.row_ptr, dtype=np.uint32).tobytes(), np.ascontiguousarray(M.col_idx, dtype=np.uint32).tobytes(),
                     np.ascontiguousarray(M.coef_idx, dtype=np.uint32).tobytes()]
        blob.append(ints_to_limbs(r1.coef_table, P, mont=True).tobytes())
        a_bases = None
        for name, curve in (("a", ozl.BN254_G1), ("b1", ozl.BN254_G1), ("b2", ozl.BN254_G2), ("h", ozl.BN254_G1), ("l", ozl.BN254_G1)):
            h, cnt = pk.query_handles[name]
            arr = Bases(ctx, h, curve, cnt).download()
            if name == "a":
                a_bases = arr
            inf = np.packbits(~arr.any(axis=1), bitorder="little")
            blob += [np.array([cnt], dtype=np.uint64).tobytes(), arr.tobytes(), _pad8(inf.tobytes()[: (cnt + 7) // 8])]
        blob += [vk.alpha_g1.tobytes(), pk.beta_g1.tobytes(), pk.delta_g1.tobytes(), vk.beta_g2.tobytes(), vk.delta_g2.tobytes()]
        blob += [z_m.tobytes(), z_c.tobytes(), ints_to_limbs([r]).tobytes(), ints_to_limbs([s]).tobytes()]
        path = tmp_path / "key.blob"
        path.write_bytes(b"".join(blob))
        exe = _build(tmp_path)
        out = tmp_path / "proof.out"
        res = subprocess.run([exe, str(path), str(out)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        got = np.frombuffer(out.read_bytes(), dtype=np.uint64)
        assert (got[0:8] == proof.a).all() and (got[8:24] == proof.b).all() and (got[24:32] == proof.c).all()
        # the pageable-scalar MSM of the replay == the in-process MSM over the same bases
        hb = ctx.upload_bases(ozl.BN254_G1, a_bases, np.packbits(~a_bases.any(axis=1), bitorder="little"))
        try:
            exp, _ = ctx.jacobian_to_affine(ozl.BN254_G1, hb.msm(z_c))
        finally:
            hb.free()
        gotm, _ = ctx.jacobian_to_affine(ozl.BN254_G1, got[32:44])
        assert (gotm == exp).all()
        assert Groth16.verify(vk, [z[1]], proof)
    finally:
        pk.free()
