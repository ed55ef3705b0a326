"""Band decomposition of one large image: host logic on CPU with gloo, world_size 2 and 3.
The sweep is a plain numpy red-black point SOR of the PDE4 family with the reference's border fill --
a stand-in with the same data dependencies as libpdegpu's point kernel -- so the test pins what matters
here: with H = 2T halo columns exchanged every T sweeps, the owned columns of every band equal the
single-domain result bit for bit."""
import os
import socket

import numpy as np
import pytest

from pdegpu import bands


def test_band_columns_even_and_complete():
    for ncols in (8, 37, 640, 16384):
        for w in (1, 2, 3, 4, 8):
            if ncols < 2 * w:
                continue
            b = bands.band_columns(ncols, w)
            assert b[0][0] == 0 and b[-1][1] == ncols and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert all(a % 2 == 0 and c > a for a, c in b)


def rb_sweeps(x, tr, rhs, w, n, omega=1.5):
    """n red-black SOR sweeps + border fill, arrays [ncols, nrows] (x[j, i]); interior only."""
    wW, wN, wE, wS = w
    nc, nr = x.shape
    jj, ii = np.meshgrid(np.arange(nc), np.arange(nr), indexing="ij")
    for _ in range(n):
        for colour in (0, 1):
            m = ((ii + jj) & 1) == colour
            m[0, :] = m[-1, :] = False
            m[:, 0] = m[:, -1] = False
            nb = np.zeros_like(x)
            nb[1:-1, 1:-1] = (x[2:, 1:-1] * wE[1:-1, 1:-1] + x[:-2, 1:-1] * wW[1:-1, 1:-1]
                              + x[1:-1, 2:] * wS[1:-1, 1:-1] + x[1:-1, :-2] * wN[1:-1, 1:-1])
            new = (np.float32(1 - omega) * x + np.float32(omega) * (rhs + nb) / tr).astype(np.float32)
            x[m] = new[m]
        x[:, 0] = x[:, 1]; x[:, -1] = x[:, -2]          # rows first (i = 0 / nr-1) ...
        x[0, :] = x[1, :]; x[-1, :] = x[-2, :]          # ... then columns (j = 0 / nc-1)
    return x


def problem(nr, nc, seed=0):
    rng = np.random.default_rng(seed)
    f = lambda: rng.random((nc, nr)).astype(np.float32)
    w = [f() * 0.5 + 0.1 for _ in range(4)]
    tr = (sum(w) + 0.5 + f()).astype(np.float32)
    return f(), tr, f(), w


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, nr, nc, T, iters, parity_shift):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, tr, rhs, w = problem(nr, nc)
    plan = bands.BandPlan(nr, nc, rank, world, sweeps_per_exchange=T)
    xl = torch.from_numpy(np.ascontiguousarray(plan.take_local(x)))
    loc = [np.ascontiguousarray(plan.take_local(a)) for a in (tr, rhs, *w)]
    assert plan.a0 % 2 == 0                              # colours agree in band and image coordinates

    def sweep(n):
        rb_sweeps(xl.numpy(), loc[0], loc[1], loc[2:], n)

    sent = bands.relax_bands(plan, [xl], sweep, iters)
    q.put((rank, plan.j0, plan.j1, xl.numpy()[plan.own].copy(), sent))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T,iters", [(2, 1, 5), (2, 2, 5), (3, 1, 4), (3, 3, 7)])
def test_bands_equal_single_domain(world, T, iters):
    import torch.multiprocessing as mp
    nr, nc = 21, 50
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, nr, nc, T, iters, 0)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    x, tr, rhs, w = problem(nr, nc)
    ref = rb_sweeps(x.copy(), tr, rhs, w, iters)
    for rank, j0, j1, own, sent in got:
        assert np.array_equal(own, ref[j0:j1]), f"rank {rank}: band [{j0},{j1}) differs from the single-domain sweep"
        ncuts = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
        nex = -(-iters // T)
        assert sent == ncuts * nex * 2 * T * nr * 4
