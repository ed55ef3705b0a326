"""The batched, pipelined host-pointer entry points (pdegpu_oflow_sor_{llin4,elin4}_2d_batch): bitwise equal to one
gateway call per system (Oflow_sor_llin4_2d.c:355-361 solves one system per call), for ragged chunkings, both solvers,
and sweep by sweep against the oracle in the reference's line order."""
import numpy as np
import pytest

from pdegpu import synth
from util import assert_bitwise

pytestmark = pytest.mark.gpu

LATE = ("U", "V", "dU", "dV", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")
EARLY = ("U", "V", "M", "Cu", "Cv", "Du", "Dv", "wW", "wN", "wE", "wS")


def systems(late, shape, batch, seed=100):
    ss = [synth.flow_system(seed + b, *shape, late=late) for b in range(batch)]
    keys = LATE if late else EARLY
    return ss, {k: np.stack([s[k] for s in ss]) for k in keys}


@pytest.mark.parametrize("late", [True, False])
@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("shape,batch,chunk,lanes", [((37, 53), 7, 2, 3), ((120, 160), 5, 1, 2), ((64, 96), 9, 4, 3),
                                                     ((480, 640), 6, 0, 0)])
def test_batch_call_equals_single_calls(gpu, monkeypatch, late, solver, shape, batch, chunk, lanes):
    from pdegpu import lib, mex
    ss, fields = systems(late, shape, batch)
    fn = "Oflow_sor_llin4_2d" if late else "Oflow_sor_elin4_2d"
    singles = [mex.call(fn, synth.mex_args(fn, s, 4, 1.9, solver), 2) for s in ss]
    if chunk:
        monkeypatch.setenv("PDEGPU_HOST_CHUNK", str(chunk))
        monkeypatch.setenv("PDEGPU_HOST_LANES", str(lanes))
    ctx = lib.Context(0)
    try:
        o0, o1 = ctx.oflow_sor_batch(fields, late, 4, 1.9, solver)
    finally:
        ctx.close()
    for b in range(batch):
        assert_bitwise(o0[b], singles[b][0], f"system {b} unknown 0")
        assert_bitwise(o1[b], singles[b][1], f"system {b} unknown 1")


def test_batch_call_in_reference_order_matches_oracle_sweep_by_sweep(gpu, oracle):
    from pdegpu import lib
    ss, fields = systems(True, (45, 61), 5, seed=7)
    ctx = lib.Context(0)
    try:
        ctx.set_sweep_order(lib.ORDER_REFERENCE)
        for it in (1, 2, 4):
            o0, o1 = ctx.oflow_sor_batch(fields, True, it, 1.9, 2)
            for b, s in enumerate(ss):
                r = oracle.call("Oflow_sor_llin4_2d", synth.mex_args("Oflow_sor_llin4_2d", s, it, 1.9, 2), 2)
                rng = float(np.nanmax(np.abs(r[0]))) + 1e-12
                assert np.nanmax(np.abs(o0[b] - r[0])) <= 1e-5 * max(rng, 1.0)
                assert np.nanmax(np.abs(o1[b] - r[1])) <= 1e-5 * max(rng, 1.0)
    finally:
        ctx.close()


def test_batch_call_rejects_bad_arguments(gpu):
    from pdegpu import lib
    _, fields = systems(True, (16, 16), 2)
    ctx = lib.Context(0)
    try:
        with pytest.raises(lib.PdegpuError):
            ctx.oflow_sor_batch(fields, True, 4, 1.9, 3)          # no such solver
        o0, _ = ctx.oflow_sor_batch(fields, True, 0, 1.9, 2)      # iter 0: zeros, like the gateway
        assert not o0.any()
    finally:
        ctx.close()
