"""Pins the oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). Runs anywhere: needs neither /root/reference nor oracle/_ref."""
import glob
import os

import numpy as np
import pytest

from util import assert_bitwise, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# (urban3_pair / yosemite are input fixtures of tests/test_gpu_configs_fixtures.py, not reference outputs)
FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
               if os.path.basename(f) not in ("warp.npz", "urban3_pair.npz", "yosemite.npz"))


def load(path):
    z = np.load(path)
    fn, nlhs = str(z["fn"]), int(z["nlhs"])
    nin = len([k for k in z.files if k.startswith("in")])
    args = [np.asfortranarray(z[f"in{k}"]) for k in range(nin)]
    outs = [np.asfortranarray(z[f"out{k}"]) for k in range(nlhs)]
    return fn, args, nlhs, outs


def test_golden_present():
    assert len(FILES) >= 20


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_matches_golden(oracle, path):
    fn, args, nlhs, want = load(path)
    got = oracle.call(fn, args, nlhs)
    for k, (g, w) in enumerate(zip(got, want)):
        if os.path.basename(path) == "llin8_s2.npz":
            assert rel_err(g, w) < 2e-6          # natural summation order in the 8-neighbour line solver
        else:
            assert_bitwise(g, w, f"{fn} out{k}")


def test_oracle_warp_golden(oracle):
    z = np.load(os.path.join(GOLD, "warp.npz"))
    assert_bitwise(oracle.bilin(z["I"], z["X"], z["Y"], float("nan")), np.asfortranarray(z["out_nan"]), "warp nan")
    assert_bitwise(oracle.bilin(z["I"], z["X"], z["Y"], 0.0), np.asfortranarray(z["out_zero"]), "warp zero")
