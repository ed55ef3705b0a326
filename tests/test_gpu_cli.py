"""pdegpu_flow_batch: the batch command line front end, plain C on the C ABI, against the Python binding of
the same library calls: same pairs, same bits."""
import os
import subprocess

import numpy as np
import pytest

from pdegpu import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("driver,channels", [("fmg", 1), ("llin", 3)])
def test_batch_cli_equals_the_library_call(built, tmp_path, driver, channels):
    from pdegpu import lib
    exe = built.build_cli()
    nr, nc, npairs = 64, 80, 3
    pairs = [synth.image_pair(60 + k, nr, nc, nframes=channels, scale=255.0, max_flow=0.8 if driver == "fmg" else 2.0) for k in range(npairs)]
    lines = []
    for k, p in enumerate(pairs):
        for name, img in (("a", p[0]), ("b", p[1])):
            np.asarray(img, dtype=np.float32).reshape(nr, nc, channels).reshape(-1, order="F").tofile(tmp_path / f"{name}{k}.raw")
        lines.append(f"{tmp_path / f'a{k}.raw'} {tmp_path / f'b{k}.raw'} {tmp_path / f'out{k}'}")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    env = dict(os.environ, PDEGPU_ORDER="fast")
    r = subprocess.run([exe, "--driver", driver, "--batch", "2", str(nr), str(nc), str(channels), str(tmp_path / "list.txt")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert f"{npairs} pairs" in r.stderr
    c = lib.Context(0)
    c.set_sweep_order(lib.ORDER_FAST)
    fn = c.flow_fmg if driver == "fmg" else c.flow_llin
    for k, p in enumerate(pairs):
        U, V = fn(p[0].reshape(nr, nc, channels), p[1].reshape(nr, nc, channels))
        Uc = np.fromfile(tmp_path / f"out{k}_U.raw", dtype=np.float32).reshape(nr, nc, order="F")
        Vc = np.fromfile(tmp_path / f"out{k}_V.raw", dtype=np.float32).reshape(nr, nc, order="F")
        assert np.array_equal(U, Uc) and np.array_equal(V, Vc), f"pair {k}"
    c.close()
