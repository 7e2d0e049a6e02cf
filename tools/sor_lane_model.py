"""Executable model of sor_lane_kernel (of_dis_b200/csrc/sor_lane_kernel.cuh), CPU only.

Replays the kernel's schedule -- warps (band, sweep), lanes = rows, one pixel per step, result rings, the
lane-skewed record layout, the asynchronous record prefetch, the progress counters and their waits -- with the
warps interleaved at random, and checks

  * that no warp ever reads a ring slot / prefetched record that does not hold the entry it expects
    (every slot carries the id of the entry written last),
  * that the protocol never deadlocks,
  * that the result equals a plain raster-scan SOR (solver.c:77-421 order) bit for bit in float32.

    python tools/sor_lane_model.py [seeds]

The constants mirror the kernel's; change them together.
"""
from __future__ import annotations

import sys

import numpy as np

C, P, R, D, DS, DP = 8, 4, 32, 6, 8, 8
f32 = np.float32


def raster_sor(rec, du, dv, K, omega):
    """rec[h,w,8] = a11 a12 a22 b1 b2 sh sv sv_top; K lexicographic sweeps, float32, the kernel's operand order."""
    h, w, _ = rec.shape
    du, dv = du.copy(), dv.copy()
    om = f32(omega)
    for _ in range(K):
        for j in range(h):
            for i in range(w):
                a11, a12, a22, b1, b2, hh, vv, vt = rec[j, i]
                du_r = du[j, i + 1] if i + 1 < w else f32(0)
                dv_r = dv[j, i + 1] if i + 1 < w else f32(0)
                su, sv = hh * du_r, hh * dv_r
                if j > 0:
                    su, sv = su + vt * du[j - 1, i], sv + vt * dv[j - 1, i]
                if j < h - 1:
                    su, sv = su + vv * du[j + 1, i], sv + vv * dv[j + 1, i]
                s1, s2 = su + b1, sv + b2
                if i > 0:
                    hl = rec[j, i - 1, 5]
                    s1, s2 = hl * du[j, i - 1] + s1, hl * dv[j, i - 1] + s2
                ou, ov = du[j, i], dv[j, i]
                du[j, i] = ou + om * (a11 * s1 + a12 * s2 - ou)
                dv[j, i] = ov + om * (a12 * s1 + a22 * s2 - ov)
    return du, dv


class Model:
    def __init__(self, w, h, K, rng, land_late):
        self.w, self.h, self.K, self.rng, self.land_late = w, h, K, rng, land_late
        self.nb = (h + 31) // 32
        self.nw = self.nb * K
        self.W2 = (w + 1) // 2
        self.ND = self.W2 + 48
        self.TLp = (self.W2 + min(h, 32) - 1 + C - 1) // C * C
        nb, ND = self.nb, self.ND
        # global memory, lane-skewed (VarRefPlanes lane mode): per (band, t, lane) two pixels
        self.rec_g = np.full((nb, ND, 32, 2, 8), np.nan, f32)
        self.dudv_g = np.full((nb, ND, 32, 4), np.nan, f32)
        # shared memory; *_id arrays hold the entry number stored last
        self.prog = np.zeros(32, np.int64)
        self.ring = np.full((self.nw, R, 32, 4), np.nan, f32)
        self.ring_id = np.full((self.nw, R), -10**9, np.int64)
        self.recs = np.full((self.nw, DS, 32, 2, 8), np.nan, f32)
        self.recs_id = np.full((self.nw, DS), -10**9, np.int64)
        self.prev = np.full((nb, DP, 33, 4), np.nan, f32)
        self.prev_id = np.full((nb, DP, 33), -10**9, np.int64)
        self.reads_checked = 0

    def load(self, rec, du, dv):
        for j in range(self.h):
            b, l = divmod(j, 32)
            for i in range(self.w):
                self.rec_g[b, i // 2 + l, l, i & 1] = rec[j, i]
                self.dudv_g[b, i // 2 + l, l, 2 * (i & 1):2 * (i & 1) + 2] = (du[j, i], dv[j, i])

    def result(self):
        du = np.zeros((self.h, self.w), f32)
        dv = np.zeros((self.h, self.w), f32)
        for j in range(self.h):
            b, l = divmod(j, 32)
            for i in range(self.w):
                du[j, i], dv[j, i] = self.dudv_g[b, i // 2 + l, l, 2 * (i & 1):2 * (i & 1) + 2]
        return du, dv

    def warp(self, wi, omega):
        """Generator: one warp of the kernel; yields at every point where another warp may run."""
        nb, K, w, h, ND, TLp, W2 = self.nb, self.K, self.w, self.h, self.ND, self.TLp, self.W2
        k, b = divmod(wi, nb)
        lanes = np.arange(32)
        j = 32 * b + lanes
        row_ok = j < h
        first_row, last_row = j == 0, j >= h - 1
        has_above, has_below = b > 0, b + 1 < nb
        k0, klast = k == 0, k == K - 1
        om = f32(omega)
        NONE = -(1 << 30)
        off = np.full(32, NONE, np.int64)

        def dep(bb, kk, o):
            if 0 <= bb < nb and 0 <= kk < K:
                x = kk * nb + bb
                off[x] = max(off[x], o)

        dep(b, k - 1, 3)
        dep(b + 1, k - 1, -29)
        dep(b - 1, k, 33)
        dep(b, k + 1, -R + 1)
        dep(b - 1, k + 1, -R + 32)
        dep(b + 1, k, -R - 30)

        pending = []  # cp.async groups: lists of closures

        def issue(tp, grp):
            def land(tp=tp):  # global memory is read when the copy lands (latest) or at issue (earliest)
                self.recs[wi, tp % DS] = self.rec_g[b, tp]
                self.recs_id[wi, tp % DS] = tp
            grp.append(land)
            if k0:
                te = tp + 1

                def land2(te=te):
                    self.prev[b, te % DP][:32] = self.dudv_g[b, te]
                    self.prev_id[b, te % DP][:32] = te
                grp.append(land2)
                if has_below and te >= 32:
                    def land3(te=te):
                        self.prev[b, te % DP][32] = self.dudv_g[b + 1, te - 32, 0]
                        self.prev_id[b, te % DP][32] = te
                    grp.append(land3)

        def commit(grp):
            if self.land_late:
                pending.append(grp)
            else:
                for f in grp:
                    f()
                pending.append([])

        def wait(n):
            while len(pending) > n:
                for f in pending.pop(0):
                    f()

        g0 = []
        if k0:
            def land0():
                self.prev[b, 0][:32] = self.dudv_g[b, 0]
                self.prev_id[b, 0][:32] = 0
            g0.append(land0)
        for tp in range(D):
            grp = g0 if tp == 0 else []
            issue(tp, grp)
            commit(grp)

        watched = off != NONE

        def blocked(t):
            lim = np.where(self.prog >= TLp, 1 << 40, self.prog - off)[watched]
            return lim.size > 0 and t > lim.min()

        def ensure(t):
            spins = 0
            while blocked(t):
                spins += 1
                if spins > 200000:
                    raise RuntimeError("deadlock: warp %d (b=%d,k=%d) at t=%d prog=%s" % (wi, b, k, t, self.prog[:self.nw]))
                yield

        pr = (k - 1) * nb + b if k > 0 else 0

        def load_operands(t1):
            """operands of step t1 (records, next/bottom previous-sweep blocks, halo row), with id checks"""
            I1 = t1 - lanes
            act1 = row_ok & (I1 >= 0) & (I1 < W2)
            has_r1 = 2 * I1 + 2 < w  # the block's second pixel has a right neighbour in the next block
            rec = self.recs[wi, t1 % DS].copy()
            if act1.any():
                assert self.recs_id[wi, t1 % DS] == t1, "record ring: stale slot"
            if k0:
                slot = (t1 + 1) % DP
                nx = self.prev[b, slot][:32].copy()
                bt = self.prev[b, slot][1:33].copy()
                assert (self.prev_id[b, slot][:32][act1 & has_r1] == t1 + 1).all(), "prev ring: next block stale"
                assert (self.prev_id[b, slot][1:33][act1 & ~last_row] == t1 + 1).all(), "prev ring: bottom block stale"
            else:
                slot = (t1 + 1) % R
                nx = self.ring[pr, slot].copy()
                if (act1 & (has_r1 | ~last_row))[:31].any() or (act1 & has_r1)[31]:
                    assert self.ring_id[pr, slot] == t1 + 1, "ring (b,k-1): entry t+1 stale (%d)" % self.ring_id[pr, slot]
                bt = np.empty((32, 4), f32)
                bt[:31] = nx[1:]
                if has_below:
                    bt[31] = self.ring[pr + 1, slot, 0]
                    if act1[31] and not last_row[31]:
                        assert self.ring_id[pr + 1, slot] == t1 - 31, "ring (b+1,k-1): entry t-31 stale"
                else:
                    bt[31] = nx[31]
            th = np.zeros(4, f32)
            if has_above:
                slot_t = (t1 + 31) % R
                th = self.ring[wi - 1, slot_t, 31].copy()
                if act1[0]:
                    assert self.ring_id[wi - 1, slot_t] == t1 + 31, "ring (b-1,k): entry t+31 stale"
            return rec, nx, bt, th

        def pixel(r8, ou, ov, ru, rv, tu, tv, bu, bv, lu, lv, hl, has_l, has_r):
            with np.errstate(all="ignore"):
                a11, a12, a22, b1, b2, hh, vv, vt = (r8[:, q] for q in range(8))
                du_r = np.where(has_r, ru, f32(0))
                dv_r = np.where(has_r, rv, f32(0))
                t1u, t1v = hh * du_r, hh * dv_r
                t2u, t2v = t1u + vt * tu, t1v + vt * tv
                bsu, bsv = np.where(first_row, t1u, t2u), np.where(first_row, t1v, t2v)
                t3u, t3v = bsu + vv * bu, bsv + vv * bv
                s1 = np.where(last_row, bsu, t3u) + b1
                s2 = np.where(last_row, bsv, t3v) + b2
                B1w, B2w = hl * lu + s1, hl * lv + s2
                B1, B2 = np.where(has_l, B1w, s1), np.where(has_l, B2w, s2)
                du = ou + om * (a11 * B1 + a12 * B2 - ou)
                dv = ov + om * (a12 * B1 + a22 * B2 - ov)
            return du.astype(f32), dv.astype(f32)

        yield from ensure(-1)
        wait(D - 2)
        if k0:
            cur = self.prev[b, 0][:32].copy()
            assert self.prev_id[b, 0][0] == 0
        else:
            cur = self.ring[pr, 0].copy()
            assert self.ring_id[pr, 0] == 0
        rec, nxt, bot, th = load_operands(0)
        res = np.zeros((32, 4), f32)
        hl = np.zeros(32, f32)
        for t0 in range(0, TLp, C):
            for s in range(C):
                t = t0 + s
                if s % P == 0 and t > 0:
                    self.prog[wi] = t
                I = t - lanes
                if s % 2 == 0:
                    yield from ensure(t + 1)  # this step and the next
                top = np.vstack((res[:1], res[:31]))
                wait(D - 2)
                yield
                rec1, nxt1, bot1, th1 = load_operands(t + 1)
                grp = []
                issue(t + D, grp)
                commit(grp)
                act = row_ok & (I >= 0) & (I < W2)
                self.reads_checked += int(act.sum()) + int((act & (2 * I + 1 < w)).sum())
                if has_above:
                    top[0] = th
                i0 = 2 * I
                du0, dv0 = pixel(rec[:, 0], cur[:, 0], cur[:, 1], cur[:, 2], cur[:, 3], top[:, 0], top[:, 1], bot[:, 0], bot[:, 1],
                                 res[:, 2], res[:, 3], hl, I > 0, i0 + 1 < w)
                du1, dv1 = pixel(rec[:, 1], cur[:, 2], cur[:, 3], nxt[:, 0], nxt[:, 1], top[:, 2], top[:, 3], bot[:, 2], bot[:, 3],
                                 du0, dv0, rec[:, 0, 5], np.ones(32, bool), i0 + 2 < w)
                hl = rec[:, 1, 5].copy()
                res = np.stack((du0, dv0, du1, dv1), axis=1)
                yield  # the loads above are long done when the store is issued: let other warps run in between
                self.ring[wi, t % R] = res
                self.ring_id[wi, t % R] = t
                if klast:
                    self.dudv_g[b, t] = res  # unpredicated: lanes without a block write positions nobody reads
                rec, cur, nxt, bot, th = rec1, nxt, nxt1, bot1, th1
        self.prog[wi] = TLp

    def run(self, omega):
        gens = [self.warp(wi, omega) for wi in range(self.nw)]
        alive = list(range(self.nw))
        # random interleaving with bursts: a warp runs 1..12 scheduling points in a row; some warps are starved
        # (picked 30x less often) so that producers run far ahead of consumers and the other way round
        weight = np.where(self.rng.random(self.nw) < 0.3, 1.0 / 30, 1.0)
        while alive:
            pw = weight[alive] / weight[alive].sum()
            wi = alive[self.rng.choice(len(alive), p=pw)]
            for _ in range(int(self.rng.integers(1, 13))):
                try:
                    next(gens[wi])
                except StopIteration:
                    alive.remove(wi)
                    break


def one_case(w, h, K, seed, land_late):
    rng = np.random.default_rng(seed)
    rec = rng.uniform(0.05, 0.4, (h, w, 8)).astype(f32)
    rec[..., 3:5] = rng.uniform(-1, 1, (h, w, 2)).astype(f32)
    du0 = rng.uniform(-1, 1, (h, w)).astype(f32)
    dv0 = rng.uniform(-1, 1, (h, w)).astype(f32)
    for j in range(h):  # sv_top(i,j) = sv(i,j-1)
        rec[j, :, 7] = rec[j - 1, :, 6] if j > 0 else 0
    m = Model(w, h, K, rng, land_late)
    m.load(rec, du0, dv0)
    m.run(1.6)
    du, dv = m.result()
    eu, ev = raster_sor(rec, du0, dv0, K, 1.6)
    ok = np.array_equal(du.view(np.uint32), eu.view(np.uint32)) and np.array_equal(dv.view(np.uint32), ev.view(np.uint32))
    return ok, m.reads_checked


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cases = [(20, 14, 3), (33, 28, 3), (40, 56, 3), (17, 70, 2), (64, 33, 1), (9, 100, 3), (50, 64, 4), (37, 32, 3), (5, 40, 3), (3, 64, 3), (2, 33, 2), (130, 20, 3), (1, 96, 3)]
    bad = 0
    for (w, h, K) in cases:
        for seed in range(seeds):
            for late in (True, False):
                ok, n = one_case(w, h, K, seed, late)
                print("w=%d h=%d K=%d seed=%d land_%s: %s (%d pixel reads checked)" % (w, h, K, seed, "late" if late else "early", "bitwise equal" if ok else "MISMATCH", n))
                bad += not ok
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
