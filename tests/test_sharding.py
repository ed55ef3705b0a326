"""Multi-rank host logic on CPU (gloo, world_size 2): batch sharding has no data-path collective and must
return, per item, exactly what a single rank computes ('distributed')."""
import os
import socket

import numpy as np
import pytest

from pdegpu import shard, synth


def test_partition_is_balanced_and_complete():
    for n in (0, 1, 7, 8, 512, 513):
        for w in (1, 2, 3, 4, 8):
            p = shard.partition(n, w)
            assert len(p) == w and p[0][0] == 0 and p[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(p, p[1:]))
            sizes = [b - a for a, b in p]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _solve(seed):
    # the compute leg of the CPU test is the oracle (allowed in tests); on the GPU box it is libpdegpu
    from oracle.oracle import OracleBackend
    s = synth.flow_system(seed, 19, 23, late=True)
    out = OracleBackend().call("Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, 2, 1.9, 2), 2)
    return out[0].tobytes() + out[1].tobytes()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = list(range(300, 307))                      # 7 items over 2 ranks: ragged split
    res = shard.run_sharded(seeds, _solve)
    t = shard.max_over_ranks(1.0 + rank)
    q.put((rank, res, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_matches_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [_solve(s) for s in range(300, 307)]
    for rank, res, t in got:
        assert res == want, f"rank {rank}: sharded results differ from the single-rank run"
        assert t == 2.0                                   # max over ranks
