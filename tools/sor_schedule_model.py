"""Symbolic model of sor_wave_kernel's data flow (of_dis_b200/csrc/sor_wave_kernel.cuh): every value is
a tag (I, j, sweep); the model replays the producer's loads, the stage ring, the double-buffered board,
the cluster halo stores and the in-place (du,dv) writes super-step by super-step and asserts that every
block reads exactly the operands the lexicographic scan gives it (top/left of this sweep, own/right/
bottom of the previous one) and that no bulk copy reads a location that is written while the copy may
still be in flight.  CPU-only; python tools/sor_schedule_model.py [W4 h HPAD K]."""
import itertools
import sys

PF = 3


def run(W4, h, HPAD, K, verbose=False):
    nb = (h + HPAD - 1) // HPAD
    NR = 2 * K + PF
    hb = HPAD + 2
    S = W4 + h + 2 * K - 2
    # global (du,dv): tag of the sweep whose value is stored, per block (I, j); -1 = before this solve
    glob = {(I, j): -1 for I in range(W4) for j in range(h)}
    written_at = {}
    # per CTA state
    class CTA:
        pass
    ctas = []
    for c in range(nb):
        t = CTA()
        t.c, t.j0 = c, c * HPAD
        t.hloc = min(HPAD, h - t.j0)
        t.S_loc = W4 + t.hloc + 2 * K - 2
        t.dmax = W4 + t.hloc - 1
        t.stage = [None] * NR      # each: dict(rec_d, dud_d, halo_I, issued_T, snapshot)
        t.ist = 0
        t.board = [dict(), dict()]  # parity -> {(k, row_index): tag}
        t.halo = [dict(), dict(), dict()]  # slot -> {(dir, k): tag}; dir 0 from the band above, 1 from below
        t.thr = {}
        for k in range(K):
            for jraw in range(HPAD):
                if jraw < t.hloc:
                    t.thr[(k, jraw)] = dict(left=None, own=None, st=0, stp=0)
        ctas.append(t)
    cur, prev = 0, 1
    hcur, hprev = 0, 2
    checked = 0
    pending = []  # (cta, stage index, wait_T) loads not yet waited for
    for T in range(-PF, S):
        stores = []  # (cta index, parity, key, tag) applied at the barrier
        hstores = []  # halo ring stores (cta index, slot, (dir, k), tag)
        gwrites = []
        for t in ctas:
            tl = T - t.j0
            # ---- producer
            n = tl + PF
            if 0 <= n < t.S_loc:
                d = min(n, t.dmax)
                d1 = min(n + 1, t.dmax)
                ih = min(max(n - (HPAD - 1), 0), W4 - 1)
                snap = {}
                for jl in range(HPAD):           # (du,dv) diagonal d1 of this band
                    I = d1 - jl
                    if 0 <= I < W4 and jl < t.hloc:
                        snap[("dud", jl)] = ((I, t.j0 + jl), glob[(I, t.j0 + jl)])
                if t.c + 1 < nb:
                    snap["halo"] = ((ih, t.j0 + HPAD), glob[(ih, t.j0 + HPAD)])
                t.stage[t.ist] = dict(n=n, rec_d=d, dud_d=d1, halo_I=ih, snap=snap, issued=T)
                pending.append((t, t.ist, t.j0 + n - 1))  # waited for before the barrier ending local step n-1
                t.ist = (t.ist + 1) % NR
            # ---- compute threads
            for (k, jl), th in t.thr.items():
                n = tl - 2 * k
                I = n - jl
                j = t.j0 + jl
                warp_lo = jl & ~31
                warp_hi = min(warp_lo + 31, t.hloc - 1)
                active = tl >= 0 and warp_lo <= n + 1 and warp_hi > n - W4
                if active:
                    st = t.stage[th["st"]]
                    in_range = 0 <= I < W4
                    if in_range:  # records of the block: stage of load n, still resident for every sweep
                        assert st is not None and st["n"] == n, ("stage of load n", T, t.c, k, jl, st and st["n"], n)
                        assert st["rec_d"] == min(n, t.dmax) == I + jl
                        assert st["issued"] < T, "load must have been issued (and waited for) before use"
                    if k == 0:
                        if in_range:
                            # own: diagonal n staged with load n-1 (or global for n == 0)
                            if n >= 1:
                                sp = t.stage[th["stp"]]
                                assert sp["n"] == n - 1 and sp["dud_d"] == n, ("own stage", T, t.c, jl)
                                own = sp["snap"][("dud", jl)]
                            else:
                                own = ((I, j), glob[(I, j)])
                            assert own == ((I, j), -1), ("own", own, I, j)
                            if I + 1 < W4:
                                assert st["dud_d"] == n + 1
                                rf = st["snap"][("dud", jl)]
                                assert rf == ((I + 1, j), -1), ("rf", rf)
                            if j + 1 < h:
                                if jl + 1 < HPAD:
                                    bot = st["snap"][("dud", jl + 1)]
                                else:
                                    assert st["halo_I"] == I, ("halo block", st["halo_I"], I)
                                    bot = st["snap"]["halo"]
                                assert bot == ((I, j + 1), -1), ("bot k0", bot, I, j)
                    else:
                        nxt = t.board[prev].get((k - 1, jl + 1))
                        if in_range:
                            assert th["own"] == (I, j, k - 1), ("own k>0", th["own"], (I, j, k - 1), T)
                            if I + 1 < W4:
                                assert nxt == (I + 1, j, k - 1), ("right", nxt, (I + 1, j, k - 1), T)
                            if j + 1 < h:
                                if nb > 1 and t.c + 1 < nb and jl == t.hloc - 1:
                                    bot = t.halo[hprev].get((1, k - 1))
                                else:
                                    bot = t.board[prev].get((k - 1, jl + 2))
                                assert bot == (I, j + 1, k - 1), ("bot", bot, (I, j + 1, k - 1), T, t.c, k, jl)
                        th["own"] = nxt
                    if in_range:
                        if j > 0:
                            top = t.halo[hprev].get((0, k)) if (nb > 1 and t.c > 0 and jl == 0) else t.board[prev].get((k, jl))
                            assert top == (I, j - 1, k), ("top", top, (I, j - 1, k), T, t.c, k, jl)
                        if I > 0:
                            assert th["left"] == (I - 1, j, k), ("left", th["left"], (I - 1, j, k))
                        checked += 1
                    tag = (I, j, k) if in_range else ("junk", T, t.c, k, jl)
                    th["left"] = tag
                    stores.append((t.c, cur, (k, jl + 1), tag))
                    th["last"] = tag
                # unconditional send of the latest block to the neighbouring band's halo ring
                if nb > 1 and "last" in th:
                    if jl == 0 and t.c > 0:
                        hstores.append((t.c - 1, hcur, (1, k), th["last"]))
                    elif jl == t.hloc - 1 and t.c + 1 < nb:
                        hstores.append((t.c + 1, hcur, (0, k), th["last"]))
                if active:
                    if k == K - 1 and in_range:
                        gwrites.append(((I, j), k))
                if n >= 0:
                    th["stp"] = th["st"]
                    th["st"] = (th["st"] + 1) % NR
        # barrier: stores become visible, global writes land
        for ci, par, key, tag in stores:
            ctas[ci].board[par][key] = tag
        for ci, slot, key, tag in hstores:
            ctas[ci].halo[slot][key] = tag
        for key, k in gwrites:
            # no bulk copy that may still be in flight may have this block as its source
            for t, si, wait_T in pending:
                stg = t.stage[si]
                if wait_T >= T:
                    for what, (blk, _) in stg["snap"].items():
                        assert blk != key, ("in-place write races a bulk copy", key, T, t.c, stg["n"])
            assert glob[key] == -1
            glob[key] = k
        pending = [(t, si, w) for (t, si, w) in pending if w >= T]
        cur, prev = prev, cur
        hprev, hcur = hcur, (hcur + 1) % 3
    assert checked == W4 * h * K, (checked, W4 * h * K)
    assert all(v == K - 1 for v in glob.values())
    if verbose:
        print("ok: W4=%d h=%d HPAD=%d K=%d bands=%d, %d block updates checked" % (W4, h, HPAD, K, nb, checked))


if __name__ == "__main__":
    if len(sys.argv) == 5:
        run(*map(int, sys.argv[1:]), verbose=True)
    else:
        for W4, h, HPAD, K in [(5, 20, 32, 3), (9, 70, 32, 2), (3, 64, 32, 1), (12, 100, 32, 3), (7, 33, 32, 5),
                               (20, 130, 64, 3), (4, 200, 64, 1), (16, 56, 64, 3), (2, 96, 32, 2), (1, 65, 32, 3)]:
            run(W4, h, HPAD, K, verbose=True)
