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


def _worker(rank, world, port, q, nr, nc, T, iters):
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
    band = bands.GpuBand(ctx, plan, lib.FLOW_LLIN4, f)
    band.relax(iters, 1.0)
    ctx.sync()
    q.put((rank, plan.j0, plan.j1, f["dU"][plan.own].cpu().numpy(), f["dV"][plan.own].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("T,iters", [(1, 4), (2, 5)])
def test_two_gpu_bands_equal_one_gpu(built, T, iters):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pdegpu import lib
    nr, nc = 256, 384
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q, nr, nc, T, iters)) for r in range(2)]
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
