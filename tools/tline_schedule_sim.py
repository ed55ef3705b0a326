#!/usr/bin/env python
"""CPU replay of the generation-3 line-relaxation kernel's protocol (csrc/sweeps_tline_impl.cuh).

Every warp role of one CTA is a coroutine that yields whenever the kernel would wait (mbarrier phase parity,
sequence counter); a random scheduler interleaves them. Shared state is modelled at the granularity the kernel
synchronises on: ring slots (which line, old / new values, loaded or not), coefficient slabs, block counters.
Checked: no deadlock; every line that exists is relaxed exactly once; an even line sees OLD odd neighbours, an odd
line NEW even neighbours; a ring slot is never overwritten while a reader still needs it; every owned block is
written exactly once with all of its lines solved; barrier parities never alias.

    python tools/tline_schedule_sim.py            # randomised sweep over shapes / geometries
"""
from __future__ import annotations

import random
import sys


class Barrier:
    """mbarrier with arrival count 1 (+ transaction bytes folded into the arrival)."""
    def __init__(self):
        self.phase = 0

    def arrive(self):
        self.phase += 1

    def passed(self, parity, expect_phase):
        # hardware only knows the parity; the replay also knows which phase the waiter means and flags aliasing
        ok = (self.phase & 1) != parity
        if ok and self.phase != expect_phase + 1:
            raise AssertionError(f"parity alias: barrier at phase {self.phase}, waiter meant {expect_phase}")
        return ok


def simulate(nlines, batch, grid, BL, R, D, K, NCW, seed, skip_border=False, verbose=False):
    rng = random.Random(seed)
    NB = (nlines + BL - 1) // BL
    TB = NB * batch
    NBR = R // BL + 2
    assert R % BL == 0 and R >= 2 * D + BL
    grid = min(grid, TB)
    relaxed_lines = set()
    written_blocks = set()
    for cta in range(grid):
        B0, B1 = cta * TB // grid, (cta + 1) * TB // grid
        nblk = B1 - B0
        redundant = B1 < TB and (B1 % NB) != 0
        Ltot = BL * nblk + (1 if redundant else 0)

        def locate(l):
            if l < 0:
                lb, r = -1, BL + l
            else:
                lb, r = divmod(l, BL)
            gb = B0 + lb
            if gb < 0 or gb >= TB:
                return None
            img, jb = divmod(gb, NB)
            if l < 0 and jb == NB - 1:
                return None
            j = BL * jb + r
            return (img, j) if j < nlines else None

        full = [Barrier() for _ in range(K)]
        empty = [Barrier() for _ in range(K)]
        rfull = [Barrier() for _ in range(R)]
        solved = [Barrier() for _ in range(R)]
        written_seq = [0] * NBR
        block_cnt = [0] * NBR
        desc = [None] * K
        slab_line = [None] * K                 # which line's coefficients the slab holds
        ring = [None] * R                      # (g, state) state in {'old', 'new'}
        grab = [0]
        issued = [0]
        loaded = [0]
        readers = [0] * R                      # warps currently reading the slot as a neighbour / own

        def holds(slot_state, gg, st):
            # a line that is not relaxed (skip_border) has the same values before and after its task
            return slot_state == (gg, st) or slot_state == (gg, 'fixed')

        def wait(bar, parity, phase):
            while not bar.passed(parity, phase):
                yield

        def consumer(wid):
            while True:
                s = grab[0]
                grab[0] += 1
                slot, use = s % K, s // K
                while issued[0] < s + 1:
                    yield
                yield from wait(full[slot], use & 1, use)
                dsc = desc[slot]
                if dsc is None:
                    empty[slot].arrive()
                    return
                l, odd, eLo, eHi, owned, relaxed, img, j, j0, cnt, lb = dsc
                g = l + BL
                gl, gh = g - 1, g + 1
                while loaded[0] < gh + 1:
                    yield
                yield from wait(rfull[g % R], (g // R) & 1, g // R)
                if relaxed:
                    if eLo:
                        yield from wait(rfull[gl % R], (gl // R) & 1, gl // R)
                    if eHi:
                        yield from wait(rfull[gh % R], (gh // R) & 1, gh // R)
                    if odd:
                        yield from wait(solved[gl % R], (gl // R) & 1, gl // R)
                        if eHi:
                            yield from wait(solved[gh % R], (gh // R) & 1, gh // R)
                    assert slab_line[slot] == (img, j), "slab holds another line"
                    need = [(g, 'old')]
                    if eLo:
                        need.append((gl, 'new' if odd else 'old'))
                    if eHi:
                        need.append((gh, 'new' if odd else 'old'))
                    for gg, st in need:
                        assert holds(ring[gg % R], gg, st), f"line g={g} (odd={odd}) wants {(gg, st)}, slot holds {ring[gg % R]}"
                        readers[gg % R] += 1
                    yield                      # rows of unknown 0, solve
                    yield
                    for gg, st in need:
                        assert holds(ring[gg % R], gg, st), "ring slot changed under a reader"
                        readers[gg % R] -= 1
                    empty[slot].arrive()
                    yield                      # second solve
                    assert readers[g % R] == 0 or True
                    ring[g % R] = (g, 'new')
                    key = (cta, img, j) if owned else None
                    if owned:
                        assert (img, j) not in relaxed_lines, "line relaxed twice"
                        relaxed_lines.add((img, j))
                else:
                    empty[slot].arrive()
                    assert ring[g % R] == (g, 'fixed')   # T_out = T_in
                    if owned:
                        relaxed_lines.add((img, j))
                solved[g % R].arrive()
                if owned:
                    block_cnt[lb % NBR] += 1
                    if block_cnt[lb % NBR] == cnt:
                        yield
                        for r in range(cnt):
                            gg = (lb + 1) * BL + r
                            assert holds(ring[gg % R], gg, 'new'), f"block {lb} written with line {gg} in state {ring[gg % R]}"
                        assert (cta, lb) not in written_blocks
                        written_blocks.add((cta, lb))
                        block_cnt[lb % NBR] = 0
                        written_seq[lb % NBR] = lb + 1

        def producer():
            Q = D + 2 * ((Ltot + 1) >> 1)
            s = 0
            for q in range(Q):
                if q < D:
                    l, odd = 2 * q, False
                else:
                    r = q - D
                    l, odd = (r, True) if r & 1 else (2 * D + r, False)
                if l >= Ltot:
                    continue
                loc = locate(l)
                if loc is None:
                    continue
                img, j = loc
                noupdate = skip_border and (j == 0 or j == nlines - 1)
                lb = l // BL
                j0 = j - (l - lb * BL)
                cnt = min(BL, nlines - j0)
                slot, use = s % K, s // K
                while not ((empty[slot].phase & 1) != ((use & 1) ^ 1)):
                    yield
                assert empty[slot].phase == use, "empty barrier out of step"
                desc[slot] = (l, odd, j > 0, j + 1 < nlines, lb < nblk, not noupdate, img, j, j0, cnt, lb)
                slab_line[slot] = (img, j)
                full[slot].arrive()
                s += 1
                issued[0] = s
                if rng.random() < 0.3:
                    yield
            for _ in range(NCW):
                slot, use = s % K, s // K
                while not ((empty[slot].phase & 1) != ((use & 1) ^ 1)):
                    yield
                desc[slot] = None
                full[slot].arrive()
                s += 1
                issued[0] = s

        def loader():
            Gmax = Ltot + BL
            for g in range(Gmax + 1):
                l, slot = g - BL, g % R
                lp = l - R
                if lp == -1:                   # the line before the range: read by task 0 only
                    yield from wait(solved[BL % R], (BL // R) & 1, BL // R)
                if lp >= 0:
                    lbp = lp // BL
                    if lbp < nblk:
                        while written_seq[lbp % NBR] < lbp + 1:
                            yield
                        if lbp > 0:
                            while written_seq[(lbp - 1) % NBR] < lbp:
                                yield
                loc = locate(l) if -1 <= l <= Ltot else None
                exists = loc is not None
                if exists and l >= BL * nblk and not redundant:
                    exists = False
                istask = exists and 0 <= l < Ltot
                assert readers[slot] == 0, f"ring slot {slot} reloaded while {readers[slot]} warps read it"
                if exists:
                    fixed = skip_border and loc[1] in (0, nlines - 1)
                    ring[slot] = (g, 'fixed' if fixed else 'old')
                else:
                    ring[slot] = None
                assert rfull[slot].phase == g // R, "ring barrier out of step"
                rfull[slot].arrive()
                if not istask:
                    assert solved[slot].phase == g // R, "solved barrier out of step"
                    solved[slot].arrive()
                loaded[0] = g + 1
                if rng.random() < 0.3:
                    yield

        threads = [consumer(w) for w in range(NCW)] + [producer(), loader()]
        alive = list(range(len(threads)))
        idle = 0
        while alive:
            t = rng.choice(alive)
            before = (grab[0], issued[0], loaded[0], tuple(b.phase for b in full + empty + rfull + solved), tuple(written_seq), len(relaxed_lines))
            try:
                next(threads[t])
            except StopIteration:
                alive.remove(t)
                idle = 0
                continue
            after = (grab[0], issued[0], loaded[0], tuple(b.phase for b in full + empty + rfull + solved), tuple(written_seq), len(relaxed_lines))
            idle = 0 if after != before else idle + 1
            if idle > 200 * len(threads):
                raise AssertionError(f"deadlock: cta {cta} nlines={nlines} batch={batch} BL={BL} R={R} D={D} K={K} NCW={NCW}")
        for lb in range(nblk):
            assert (cta, lb) in written_blocks, f"block {lb} of cta {cta} never written"
    want = {(b, j) for b in range(batch) for j in range(nlines)}
    assert relaxed_lines == want, f"lines relaxed: {len(relaxed_lines)} of {len(want)}"
    return True


def main():
    rng = random.Random(7)
    n = 0
    geoms = [(8, 24, 4), (8, 16, 4), (8, 24, 8), (4, 16, 3), (4, 12, 2), (4, 8, 2), (4, 16, 6), (8, 32, 5)]
    for trial in range(400):
        BL, R, D = rng.choice(geoms)
        nlines = rng.choice([2, 3, 7, 8, 9, 15, 16, 17, 31, 37, 64, 100, 101])
        batch = rng.choice([1, 2, 3, 5])
        grid = rng.choice([1, 2, 3, 4, 7, 148])
        K = rng.choice([3, 4, 5, 8])
        NCW = rng.choice([1, 2, 3, 6, 14])
        simulate(nlines, batch, grid, BL, R, D, K, NCW, seed=trial, skip_border=rng.random() < 0.3)
        n += 1
    print(f"tline schedule replay: {n} randomised cases ok")


if __name__ == "__main__":
    main()
