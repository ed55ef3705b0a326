"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pdegpu.h declares, the 13 gateways exist, and the product fails LOUDLY without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pdegpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdegpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_13_gateways():
    syms = declared_symbols()
    for s in ("pdegpu_oflow_sor_elin4_2d", "pdegpu_oflow_sor_llin4_2d", "pdegpu_oflow_sor_llin8_2d",
              "pdegpu_oflow_lhs_elin4_2d", "pdegpu_oflow_lhs_llin4_2d", "pdegpu_disp_sor_llin4_2d",
              "pdegpu_disp_sor_llin_sym4_2d", "pdegpu_pdesolver4", "pdegpu_pdesolver8", "pdegpu_bilin_interp_2d",
              "pdegpu_fst_derivatives5", "pdegpu_snd_derivatives5", "pdegpu_ddiff_weights"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_gateway_library_exports_13_entry_points(built):
    from pdegpu import mex
    L = mex.library()
    assert L.mwsize_bytes == 8          # gateways read mwSize as mwSize
    for n in mex.NAMES:
        assert L.has("mex_" + n), n


def test_no_silent_fallback_without_gpu(built):
    """Without a visible GPU the library must refuse, not compute on the CPU."""
    from pdegpu import lib
    if lib.dll().pdegpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(lib.PdegpuError) as e:
        lib.Context(0)
    assert e.value.code == lib.ERR_NODEVICE
    from pdegpu import mex
    z = np.zeros((8, 8), np.float32, order="F")
    with pytest.raises(mex.MexError):
        mex.FstDerivatives5(z, z)


def test_gateways_reject_non_single_and_bad_counts(built):
    """Argument checks come before any GPU work, so they run everywhere (reference behaviour:
    mexErrMsgTxt on wrong nrhs / non-single input / too few outputs, Oflow_sor_elin4_2d.c:108-297)."""
    from pdegpu import mex, synth
    s = synth.flow_system(1, 9, 11)
    good = synth.mex_args("Oflow_sor_elin4_2d", s, 4, 1.9, 2)
    with pytest.raises(mex.MexError, match="wrong number of input"):
        mex.call("Oflow_sor_elin4_2d", good[:-1], 2)
    bad = list(good)
    bad[3] = bad[3].astype(np.float64)
    with pytest.raises(mex.MexError, match="'Cu' must be a noncomplex single"):
        mex.call("Oflow_sor_elin4_2d", bad, 2)
    bad = list(good)
    bad[11] = 4.0                        # a Matlab double scalar
    with pytest.raises(mex.MexError, match="'iter'"):
        mex.call("Oflow_sor_elin4_2d", bad, 2)
    with pytest.raises(mex.MexError, match="insufficient number of outputs"):
        mex.call("Oflow_sor_elin4_2d", good, 1)
    bad = list(good)
    bad[13] = synth.f32([[7]])
    with pytest.raises(mex.MexError, match="no such solver"):
        mex.call("Oflow_sor_elin4_2d", bad, 2)
    bad = list(good)
    bad[5] = bad[5][:4, :4]              # shape mismatch: the reference would read out of bounds
    with pytest.raises(mex.MexError, match="fewer elements"):
        mex.call("Oflow_sor_elin4_2d", bad, 2)
    for fn, n in (("Oflow_sor_llin4_2d", 16), ("Oflow_sor_llin8_2d", 20), ("Oflow_lhs_elin4_2d", 9),
                  ("Oflow_lhs_llin4_2d", 11), ("Disp_sor_llin4_2d", 11), ("Disp_sor_llin_sym4_2d", 19),
                  ("PDEsolver4", 10), ("PDEsolver8", 14), ("BilinInterp_2d", 3), ("FstDerivatives5", 2),
                  ("SndDerivatives5", 2), ("DdiffWeights", 2)):
        with pytest.raises(mex.MexError):
            mex.call(fn, [synth.f32(np.zeros((6, 6)))] * (n + 1), 5)


def test_batch_cli_builds_and_rejects_bad_usage(built):
    """pdegpu_flow_batch: plain C on the C ABI; without arguments it prints its usage and exits 2
    (no GPU is touched before the arguments are valid)."""
    import subprocess
    exe = built.build_cli()
    assert exe and os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: pdegpu_flow_batch" in r.stderr
