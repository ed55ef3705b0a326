"""Host-side replay of the generation-3 line kernel's hand-off protocol (tools/tline_schedule_sim.py): ring of lines,
coefficient slabs, barrier parities, block write-out -- no deadlock, every line relaxed once with the right neighbour
states, for the geometries the library picks and for adversarial ones."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("tline_sim", os.path.join(ROOT, "tools", "tline_schedule_sim.py"))
sim = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sim)


@pytest.mark.parametrize("BL,R,D,K", [(8, 24, 4, 8), (8, 24, 4, 4), (4, 16, 3, 6), (4, 12, 2, 3), (4, 8, 2, 3), (8, 16, 4, 5)])
@pytest.mark.parametrize("nlines,batch,grid", [(640, 1, 148), (37, 3, 5), (9, 2, 3), (2, 4, 2), (101, 2, 7), (480, 2, 148)])
def test_protocol_replay(BL, R, D, K, nlines, batch, grid):
    for seed, ncw in ((1, 14), (2, 3), (3, 1)):
        assert sim.simulate(nlines, batch, grid, BL, R, D, K, ncw, seed=seed)
        assert sim.simulate(nlines, batch, grid, BL, R, D, K, ncw, seed=seed, skip_border=True)
