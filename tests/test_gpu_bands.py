"""Band split on real GPUs (needs >= 2 B200: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_bands.py -m gpu`):
2 ranks over NCCL, libpdegpu's point kernel per band, must equal the single-GPU sweep bit for bit."""
import os
import socket

import numpy as np
import pytest

from pdegpu import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


KEYS = ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")


def _worker(rank, world, port, q, nr, nc, T, iters, transport="nccl"):
    import torch
    import torch.distributed as dist
    from pdegpu import bands, lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = synth.flow_system(77, nr, nc, late=True)
    plan = bands.BandPlan(nr, nc, rank, world, sweeps_per_exchange=T)
    # [ncols, nrows] = column-major image
    f = {k: torch.from_numpy(np.ascontiguousarray(plan.take_local(np.ascontiguousarray(s[k].T)))).cuda(rank) for k in KEYS}
    ctx = lib.Context(rank)
    band = bands.GpuBand(ctx, plan, lib.FLOW_LLIN4, f, transport=transport)
    if transport == "p2p":
        band.connect_p2p()
    band.relax(iters, 1.0)
    ctx.sync()
    q.put((rank, plan.j0, plan.j1, f["dU"][plan.own].cpu().numpy(), f["dV"][plan.own].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _single_gpu(nr, nc, iters):
    import torch
    from pdegpu import lib
    s = synth.flow_system(77, nr, nc, late=True)
    t = {k: torch.from_numpy(np.ascontiguousarray(s[k].T)).cuda(0) for k in KEYS}
    ctx = lib.Context(0)
    sysd = lib.make_system(lib.FLOW_LLIN4, nr, nc, x=(t["dU"].data_ptr(), t["dV"].data_ptr()), x0=(t["U"].data_ptr(), t["V"].data_ptr()),
                           m=t["M"].data_ptr(), c=(t["Cu"].data_ptr(), t["Cv"].data_ptr()), d=(t["Du"].data_ptr(), t["Dv"].data_ptr()),
                           w=[t[k].data_ptr() for k in ("wW", "wN", "wE", "wS")])
    ctx.relax(sysd, iters, 1.0, 1)
    ctx.sync()
    return t["dU"].cpu().numpy(), t["dV"].cpu().numpy()


@pytest.mark.parametrize("world,T,iters", [(2, 1, 4), (3, 2, 5), (4, 4, 9)])
def test_bands_in_one_process_through_the_c_abi_exchange(built, world, T, iters):
    """pdegpu_band_* with pdegpu_band_connect_local: `world` bands of one image, each with its own context (stream) on
    device 0, halo columns pushed into the neighbours' mailboxes by kernels that wait on device-side flags. Needs ONE
    GPU. All bands enqueue a step before the host synchronises with any of them (include/pdegpu.h)."""
    import torch
    from pdegpu import bands, lib
    nr, nc = 256, 384
    s = synth.flow_system(77, nr, nc, late=True)
    ctxs, bs = [], []
    for r in range(world):
        plan = bands.BandPlan(nr, nc, r, world, sweeps_per_exchange=T)
        f = {k: torch.from_numpy(np.ascontiguousarray(plan.take_local(np.ascontiguousarray(s[k].T)))).cuda(0) for k in KEYS}
        c = lib.Context(0)
        ctxs.append(c)
        bs.append(bands.GpuBand(c, plan, lib.FLOW_LLIN4, f, transport="p2p"))
    torch.cuda.synchronize()
    for r, b in enumerate(bs):
        b.connect_local(bs[r - 1] if r > 0 else None, bs[r + 1] if r < world - 1 else None)
    done = 0
    while done < iters:
        n = min(T, iters - done)
        for b in bs:                                           # in ONE process: the exchange of every band is enqueued before
            b.exchange_p2p()                                   # anything that may synchronise with the host (the first relax
        for b in bs:                                           # call of a context allocates its scratch): include/pdegpu.h
            b.ctx.relax(b.sys, n, 1.0, 1)
        done += n
    for c in ctxs:
        c.sync()
    rU, rV = _single_gpu(nr, nc, iters)
    for b in bs:
        p = b.plan
        assert np.array_equal(b.f["dU"][p.own].cpu().numpy(), rU[p.j0:p.j1]) and np.array_equal(b.f["dV"][p.own].cpu().numpy(), rV[p.j0:p.j1]), \
            f"band {p.rank} of {world} differs"
    assert bs[0].xchg.bytes_sent == ((iters + T - 1) // T) * 2 * 2 * T * nr * 4      # steps x unknowns x H columns x nrows x 4 B, one neighbour


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
@pytest.mark.parametrize("T,iters", [(1, 4), (2, 5)])
def test_two_gpu_bands_equal_one_gpu(built, T, iters, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pdegpu import lib
    nr, nc = 256, 384
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q, nr, nc, T, iters, transport)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    # single GPU
    s = synth.flow_system(77, nr, nc, late=True)
    t = {k: torch.from_numpy(np.ascontiguousarray(s[k].T)).cuda(0) for k in KEYS}
    ctx = lib.Context(0)
    sysd = lib.make_system(lib.FLOW_LLIN4, nr, nc, x=(t["dU"].data_ptr(), t["dV"].data_ptr()), x0=(t["U"].data_ptr(), t["V"].data_ptr()),
                           m=t["M"].data_ptr(), c=(t["Cu"].data_ptr(), t["Cv"].data_ptr()), d=(t["Du"].data_ptr(), t["Dv"].data_ptr()),
                           w=[t[k].data_ptr() for k in ("wW", "wN", "wE", "wS")])
    ctx.relax(sysd, iters, 1.0, 1)
    ctx.sync()
    rU, rV = t["dU"].cpu().numpy(), t["dV"].cpu().numpy()
    for rank, j0, j1, dU, dV in got:
        assert np.array_equal(dU, rU[j0:j1]) and np.array_equal(dV, rV[j0:j1]), f"rank {rank} band differs"
