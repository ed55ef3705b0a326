"""CPU model of the task schedule of alr_window_kernel (csrc/sweeps_window.cu): every line of a CTA's
range is relaxed exactly once, after the lines it depends on, no wait can last forever, and every block
is written exactly once. Randomised warp speeds. Run: python tools/window_schedule_sim.py"""
import random
import sys


def simulate(nlines, batch, grid, NW, R, D, seed):
    rnd = random.Random(seed)
    NB = (nlines + 7) // 8
    TB = NB * batch
    grid = min(grid, TB)
    out_written = {}
    for cta in range(grid):
        B0, B1 = cta * TB // grid, (cta + 1) * TB // grid
        nblk = B1 - B0
        red = B1 < TB and (B1 % NB) != 0
        Ltot = 8 * nblk + (1 if red else 0)
        Q = D + 2 * ((Ltot + 1) // 2)
        NBR = R // 8
        solved_seq = [0] * R
        written_seq = [0] * NBR
        cnt = [0] * NBR
        solved_lines = {}
        # per-warp state: q index, phase
        wq = list(range(NW))
        phase = [0] * NW     # 0 = need slot, 1 = need neighbours, 2 = solving (timer), 3 = finish
        timer = [0] * NW
        steps = 0
        def decode(q):
            if q < D:
                return 2 * q, False
            r = q - D
            return (r, True) if r & 1 else (2 * D + r, False)
        active = NW
        done_w = [False] * NW
        while not all(done_w):
            steps += 1
            if steps > 200000:
                return "deadlock cta %d" % cta
            progressed = False
            order = list(range(NW))
            rnd.shuffle(order)
            for w in order:
                if done_w[w]:
                    continue
                q = wq[w]
                if q >= Q:
                    done_w[w] = True
                    progressed = True
                    continue
                l, odd = decode(q)
                lb = l >> 3
                gb = B0 + lb
                img, jb = divmod(gb, NB)
                j = 8 * jb + (l & 7)
                if l >= Ltot or j >= nlines:
                    wq[w] += NW
                    phase[w] = 0
                    progressed = True
                    continue
                owned = lb < nblk
                if phase[w] == 0:
                    ok = True
                    if l >= R:
                        lbp = (l - R) >> 3
                        ok = written_seq[lbp % NBR] >= lbp + 1 and (lbp == 0 or written_seq[(lbp - 1) % NBR] >= lbp)
                    if ok:
                        phase[w] = 1
                        progressed = True
                elif phase[w] == 1:
                    ok = True
                    if odd:
                        ok = solved_seq[(l - 1) % R] >= l
                        if j + 1 < nlines:
                            ok = ok and solved_seq[(l + 1) % R] >= l + 2
                        if ok:
                            assert solved_lines.get(l - 1) and (j + 1 >= nlines or solved_lines.get(l + 1)), "odd before even"
                    else:
                        # even lines read OLD odd neighbours from global: nothing to wait for, but they must not be in
                        # X_out order problems (out of place) -- nothing to check
                        pass
                    if ok:
                        phase[w] = 2
                        timer[w] = rnd.randint(1, 6)
                        progressed = True
                elif phase[w] == 2:
                    timer[w] -= 1
                    progressed = True
                    if timer[w] <= 0:
                        # slot must still be ours: nobody overwrote it
                        assert l not in solved_lines, "line solved twice"
                        solved_lines[l] = True
                        solved_seq[l % R] = l + 1
                        if owned:
                            cnt[lb % NBR] += 1
                            j0 = 8 * jb
                            c = min(8, nlines - j0)
                            if cnt[lb % NBR] == c:
                                # write out: all lines of the block must still be in the ring
                                for k in range(c):
                                    assert solved_seq[(8 * lb + k) % R] == 8 * lb + k + 1, "ring slot overwritten before write-out"
                                    key = (img, j0 + k)
                                    assert key not in out_written, "written twice"
                                    out_written[key] = cta
                                cnt[lb % NBR] = 0
                                written_seq[lb % NBR] = lb + 1
                        wq[w] += NW
                        phase[w] = 0
            if not progressed:
                return "deadlock cta %d (no progress)" % cta
    if len(out_written) != batch * nlines:
        return "coverage %d of %d" % (len(out_written), batch * nlines)
    return None


if __name__ == "__main__":
    bad = 0
    cases = 0
    for nlines in (8, 9, 15, 16, 17, 31, 64, 67, 128, 203, 640):
        for batch in (1, 2, 3, 7):
            for grid in (1, 2, 5, 148):
                for (NW, R, D) in [(nw, (2 * ((nw + 2) // 2) + nw + 9 + 7) & ~7, (nw + 2) // 2) for nw in (8, 7, 6, 5, 4, 3)]:   # win_geometry()
                    cases += 1
                    r = simulate(nlines, batch, grid, NW, R, D, cases)
                    if r:
                        bad += 1
                        print("FAIL", nlines, batch, grid, NW, R, D, r)
    print("%d cases, %d failures" % (cases, bad))
    sys.exit(1 if bad else 0)
