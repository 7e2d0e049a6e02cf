"""Symbolic model of sor_wave_kernel's data flow (of_dis_b200/csrc/sor_wave_kernel.cuh): every value is
a tag (I, j, sweep); the model replays the producer's loads, the stage ring, the double-buffered board,
the cluster halo ring and the in-place (du,dv) writes super-step by super-step and asserts that every
block reads exactly the operands the lexicographic scan gives it (top/left of this sweep, own/right/
bottom of the previous one) and that no bulk copy reads a location that is written while the copy may
still be in flight.  RT = rows per lane (tile of 4 columns x RT rows per thread and super-step).
CPU-only; python tools/sor_schedule_model.py [W4 h HPAD RT K]."""
import sys

PF = 4


def run(W4, h, HPAD, RT, K, verbose=False):
    HB = HPAD * RT
    nb = (h + HB - 1) // HB
    NR = PF + 1 if K == 1 else PF + 2 * K - 1   # sor_stages
    R = (h + RT - 1) // RT
    S = W4 + R + 2 * K - 2
    glob = {(I, j): -1 for I in range(W4) for j in range(h)}  # sweep whose value is stored; -1 = before this solve

    class CTA:
        pass

    ctas = []
    for c in range(nb):
        t = CTA()
        t.c, t.j0, t.r0 = c, c * HB, c * HPAD
        t.hloc = min(HB, h - t.j0)
        t.nl = (t.hloc + RT - 1) // RT
        t.S_loc = W4 + t.nl + 2 * K - 2
        t.dmax = W4 + t.nl - 1
        t.stage = [None] * NR
        t.ist = 0
        t.board = [dict(), dict()]          # parity -> {(k, board row): tag}
        t.halo = [dict(), dict(), dict()]   # slot -> {(dir, k): tag}
        t.thr = {(k, rl): dict(left=[None] * RT, own=[None] * RT, st=0, stp=0, last=None)
                 for k in range(K) for rl in range(t.nl)}
        ctas.append(t)
    cur, prev, hcur, hprev = 0, 1, 0, 2
    checked = 0
    pending = []
    for T in range(-PF, S):
        stores, hstores, gwrites = [], [], []
        for t in ctas:
            tl = T - t.r0
            n = tl + PF
            if 0 <= n < t.S_loc:  # producer: the occupied lane rows of diagonal n (records and du,dv) + the halo block
                ih = min(max(n - HPAD, 0), W4 - 1)
                snap = {}
                if n <= t.dmax:
                    for rl in range(max(0, n - (W4 - 1)), min(t.nl - 1, n) + 1):
                        I = n - rl
                        for s in range(RT):
                            j = t.j0 + rl * RT + s
                            if j < h:
                                snap[("dud", rl, s)] = ((I, j), glob[(I, j)])
                if t.c + 1 < nb:
                    snap["halo"] = ((ih, t.j0 + HB), glob[(ih, t.j0 + HB)])
                t.stage[t.ist] = dict(n=n, rec_d=n, halo_I=ih, snap=snap, issued=T)
                pending.append((t, t.ist, t.r0 + n - 2))  # waited for two super-steps before sweep 0's tile on it
                t.ist = (t.ist + 1) % NR
            for (k, rl), th in t.thr.items():
                n = tl - 2 * k
                I = n - rl
                wlo = rl & ~31
                whi = min(wlo + 31, t.nl - 1)
                active = tl >= 0 and wlo <= n + 1 and whi > n - W4
                in_range = 0 <= I < W4
                if active:
                    st = t.stage[th["st"]]
                    if in_range:
                        assert st is not None and st["n"] == n and st["rec_d"] == I + rl and st["issued"] < T
                        assert ("dud", rl, 0) in st["snap"], "lane row inside the trimmed copy"
                    own = list(th["own"])
                    rf = [None] * RT
                    nxt = [None] * RT
                    if k == 0:
                        if in_range:
                            for s in range(RT):
                                j = t.j0 + rl * RT + s
                                if j >= h:
                                    continue
                                own[s] = st["snap"][("dud", rl, s)]
                                assert own[s] == ((I, j), -1), ("own", own[s])
                            sb = t.stage[(th["st"] + 1) % NR]  # diagonal n+1: right and bottom neighbours
                            jb = t.j0 + rl * RT + RT  # row below the tile
                            if I + 1 < W4 or jb < h:
                                assert sb is not None and sb["n"] == n + 1 and t.r0 + n + 1 - 2 <= T - 1, "load n+1 landed"
                            for s in range(RT):
                                j = t.j0 + rl * RT + s
                                if j < h and I + 1 < W4:
                                    assert sb["snap"][("dud", rl, s)] == ((I + 1, j), -1)
                            if jb < h:
                                if rl + 1 < HPAD:
                                    bot = sb["snap"][("dud", rl + 1, 0)]
                                else:
                                    assert sb["halo_I"] == I
                                    bot = sb["snap"]["halo"]
                                assert bot == ((I, jb), -1), ("bot k0", bot, I, jb)
                    else:
                        for s in range(RT):
                            nxt[s] = t.board[prev].get((k - 1, rl * RT + s + 1))
                        if t.c + 1 < nb and rl == t.nl - 1:
                            botX = t.halo[hprev].get((1, k - 1))
                        else:
                            botX = t.board[prev].get((k - 1, rl * RT + RT + 1))
                        if in_range:
                            for s in range(RT):
                                j = t.j0 + rl * RT + s
                                if j >= h:
                                    continue
                                assert own[s] == (I, j, k - 1), ("own k>0", own[s], (I, j, k - 1), T)
                                if I + 1 < W4:
                                    assert nxt[s] == (I + 1, j, k - 1), ("right", nxt[s], (I + 1, j, k - 1))
                                if j + 1 < h:
                                    b = own[s + 1] if s + 1 < RT else botX
                                    assert b == (I, j + 1, k - 1), ("bot", b, (I, j + 1, k - 1), T, t.c, k, rl, s)
                    topX = t.halo[hprev].get((0, k)) if (t.c > 0 and rl == 0) else t.board[prev].get((k, rl * RT))
                    new = [None] * RT
                    for s in range(RT):
                        j = t.j0 + rl * RT + s
                        if in_range and j < h:
                            if j > 0:
                                top = topX if s == 0 else new[s - 1]
                                assert top == (I, j - 1, k), ("top", top, (I, j - 1, k), T, t.c, k, rl, s)
                            if I > 0:
                                assert th["left"][s] == (I - 1, j, k), ("left", th["left"][s], (I - 1, j, k))
                            checked += 1
                            new[s] = (I, j, k)
                            if k == K - 1:
                                gwrites.append(((I, j), k))
                        else:
                            new[s] = ("junk", T, t.c, k, rl, s)
                        th["left"][s] = new[s]
                        stores.append((t.c, cur, (k, rl * RT + s + 1), new[s]))
                    th["last"] = new
                    if k > 0:
                        th["own"] = nxt
                if nb > 1 and th["last"] is not None:  # unconditional send of the latest tile row
                    if rl == 0 and t.c > 0:
                        hstores.append((t.c - 1, hcur, (1, k), th["last"][0]))
                    elif rl == t.nl - 1 and t.c + 1 < nb:
                        hstores.append((t.c + 1, hcur, (0, k), th["last"][RT - 1]))
                if n >= 0:
                    th["stp"] = th["st"]
                    th["st"] = (th["st"] + 1) % NR
        for ci, par, key, tag in stores:
            ctas[ci].board[par][key] = tag
        for ci, slot, key, tag in hstores:
            ctas[ci].halo[slot][key] = tag
        for key, k in gwrites:
            for t, si, wait_T in pending:
                if wait_T >= T:
                    for what, (blkk, _) in t.stage[si]["snap"].items():
                        assert blkk != key, ("in-place write races a bulk copy", key, T, t.c)
            assert glob[key] == -1
            glob[key] = k
        pending = [(t, si, wt) for (t, si, wt) in pending if wt >= T]
        cur, prev = prev, cur
        hprev, hcur = hcur, (hcur + 1) % 3
    assert checked == W4 * h * K, (checked, W4 * h * K)
    assert all(v == K - 1 for v in glob.values())
    if verbose:
        print("ok: W4=%d h=%d HPAD=%d RT=%d K=%d bands=%d steps=%d, %d block updates checked" % (W4, h, HPAD, RT, K, nb, S, checked))


if __name__ == "__main__":
    if len(sys.argv) == 6:
        run(*map(int, sys.argv[1:]), verbose=True)
    else:
        for W4, h, HPAD, RT, K in [(5, 20, 32, 1, 3), (9, 70, 32, 1, 2), (12, 100, 32, 1, 3), (7, 33, 32, 1, 5),
                                   (5, 20, 32, 2, 3), (9, 70, 32, 2, 2), (12, 133, 32, 2, 3), (7, 65, 32, 2, 5),
                                   (3, 64, 32, 2, 1), (16, 56, 32, 2, 3), (6, 257, 32, 4, 3), (4, 130, 32, 4, 2),
                                   (20, 300, 64, 2, 3), (2, 96, 32, 4, 1), (1, 65, 32, 2, 3), (8, 17, 32, 4, 3)]:
            run(W4, h, HPAD, RT, K, verbose=True)
