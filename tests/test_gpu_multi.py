"""-m gpu tests of the native multi-GPU path: ozl_comm_* + ozl_msm_sharded --
shard MSM, ONE ncclAllGather of Jacobian partials on the context's stream, device-side sum.
World size 1 always runs (the NCCL path on a single GPU); world size 2 when the box has two GPUs.
Every rank must hold the same combined point, equal to the oracle's MSM over ALL points."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, name, q):
    import torch.distributed as dist
    import openzl_b200 as ozl
    from openzl_b200.multi_gpu import Comm, shard_range
    from oracle import cbind, curves
    from tests.util import random_scalars
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)    # rendezvous only; the data path is the library's NCCL
    ctx = ozl.Context(rank)
    comm = Comm.from_torch(ctx)
    scalars = random_scalars(n, curves.CURVES[name].fr.p, seed=99)
    lo, hi = shard_range(n, rank, world)
    bases = ctx.upload_bases(ozl.CURVE_IDS[name], cbind.bases_seq(name, 1 + lo, hi - lo))
    jac = comm.msm_sharded(bases, scalars[lo:hi])
    aff, inf = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], jac)
    # device-pointer variant, twice in a row on the stream (the gather buffer is reused)
    d_s = torch.from_numpy(scalars[lo:hi].view(np.int64).copy()).cuda(rank)
    d_o = torch.zeros(3 * bases.coord_limbs, dtype=torch.int64, device=f"cuda:{rank}")
    for _ in range(2):
        comm.msm_sharded_device(bases, d_s.data_ptr(), hi - lo, d_o.data_ptr())
    ctx.synchronize()
    # Jacobian representatives depend on the (atomic) order of points inside a bucket: compare affine forms
    aff2, inf2 = ctx.jacobian_to_affine(ozl.CURVE_IDS[name], d_o.cpu().numpy().view(np.uint64))
    same = bool(inf2 == inf and (aff2 == aff).all())
    q.put((rank, aff.tolist(), bool(inf), same))
    dist.barrier()
    comm.close()
    bases.free()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("name", ["bls12_381_g1", "bn254_g2"])
def test_msm_sharded_nccl(world, name):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import cbind, curves
    from tests.util import random_scalars
    n = 3001 if name.endswith("g1") else 601
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, n, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    scalars = random_scalars(n, curves.CURVES[name].fr.p, seed=99)
    exp, exp_inf = cbind.to_affine(name, cbind.msm(name, cbind.bases_seq(name, 1, n), scalars, threads=8))
    for rank, aff, inf, same in got:
        assert inf == exp_inf and same
        assert aff == exp.tolist()


def _worker_large(rank, world, port, n, q):
    import torch.distributed as dist
    import openzl_b200 as ozl
    from openzl_b200.multi_gpu import Comm, shard_range
    from oracle import curves
    from tests.util import random_scalars
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ozl.Context(rank)
    comm = Comm.from_torch(ctx)
    scalars = random_scalars(n, curves.CURVES["bls12_381_g1"].fr.p, seed=2024)
    lo, hi = shard_range(n, rank, world)
    bases = ctx.generate_bases(ozl.BLS12_381_G1, 1 + lo, hi - lo).precompute(32)
    jac = comm.msm_sharded(bases, scalars[lo:hi])          # pageable host scalars: batched H2D under the accumulation
    aff, inf = ctx.jacobian_to_affine(ozl.BLS12_381_G1, jac)
    q.put((rank, aff.tolist(), bool(inf)))
    dist.barrier()
    comm.close()
    bases.free()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_msm_sharded_large_host_batches(world):
    """`ozl_msm_sharded` at 2^22 points per rank: every shard's scalars arrive in point-range batches
    (pageable memory, 7 batches) while earlier batches are accumulated, then one all-gather + sum; the
    combined point equals [sum s_i (i + 1)] G on every rank."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import cbind, curves
    from tests.util import random_scalars
    n = world << 22
    port = _free_port()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker_large, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    scalars = random_scalars(n, curves.CURVES["bls12_381_g1"].fr.p, seed=2024)
    k = cbind.dot_mod_r("bls12_381_fr", scalars, np.arange(1, n + 1, dtype=np.uint64))
    exp, _ = cbind.to_affine("bls12_381_g1", cbind.gen_mul("bls12_381_g1", k))
    for rank, aff, inf in got:
        assert not inf and aff == exp.tolist()
