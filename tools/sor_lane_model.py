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

C, R, D, DS, DP = 8, 32, 6, 8, 8
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
        self.ND = w + 32
        self.TLp = (w + 31 + C - 1) // C * C
        nb, ND = self.nb, self.ND
        # global memory, lane-skewed (VarRefPlanes lane mode)
        self.rec_g = np.full((nb, ND, 32, 8), np.nan, f32)
        self.dudv_g = np.full((nb, ND, 32, 2), np.nan, f32)
        # shared memory; *_id arrays hold the entry number stored last (-1: never written)
        self.prog = np.zeros(32, np.int64)
        self.ring = np.full((self.nw, R, 32, 2), np.nan, f32)
        self.ring_id = np.full((self.nw, R), -10**9, np.int64)
        self.recs = np.full((self.nw, DS, 32, 8), np.nan, f32)
        self.recs_id = np.full((self.nw, DS, 32), -10**9, np.int64)
        self.prev = np.full((nb, DP, 33, 2), np.nan, f32)
        self.prev_id = np.full((nb, DP, 33), -10**9, np.int64)
        self.reads_checked = 0

    def load(self, rec, du, dv):
        for j in range(self.h):
            b, l = divmod(j, 32)
            for i in range(self.w):
                self.rec_g[b, i + l, l] = rec[j, i]
                self.dudv_g[b, i + l, l] = (du[j, i], dv[j, i])

    def result(self):
        du = np.zeros((self.h, self.w), f32)
        dv = np.zeros((self.h, self.w), f32)
        for j in range(self.h):
            b, l = divmod(j, 32)
            for i in range(self.w):
                du[j, i], dv[j, i] = self.dudv_g[b, i + l, l]
        return du, dv

    def warp(self, wi, omega):
        """Generator: one warp of the kernel; yields at every point where another warp may run."""
        nb, K, w, h, ND, TLp = self.nb, self.K, self.w, self.h, self.ND, self.TLp
        k, b = divmod(wi, nb)
        lanes = np.arange(32)
        j = 32 * b + lanes
        row_ok = j < h
        first_row, last_row = j == 0, j >= h - 1
        w_eff = np.where(row_ok, w, 0)
        has_above, has_below = b > 0, b + 1 < nb
        k0, klast = k == 0, k == K - 1
        om = f32(omega)
        NONE = -(1 << 30)
        off = np.full(32, NONE, np.int64)

        def dep(bb, kk, o):
            if 0 <= bb < nb and 0 <= kk < K:
                x = kk * nb + bb
                off[x] = max(off[x], o)

        dep(b, k - 1, C + 1)
        dep(b + 1, k - 1, C - 31)
        dep(b - 1, k, C + 31)
        dep(b, k + 1, C - 1 - R)
        dep(b - 1, k + 1, C - 1 - R + 32)
        dep(b + 1, k, C - 1 - R - 30)

        pending = []  # cp.async groups: lists of closures

        def issue(tp, grp):
            ok = row_ok & (tp - lanes >= 0) & (tp - lanes < w_eff)
            if ok.any():
                def land(ok=ok, tp=tp):  # global memory is read when the copy lands (latest) or at issue (earliest)
                    src = self.rec_g[b, tp]
                    self.recs[wi, tp % DS][ok] = src[ok]
                    self.recs_id[wi, tp % DS][ok] = tp
                grp.append(land)
            if k0:
                te = tp + 1
                ok2 = row_ok & (te - lanes >= 0) & (te - lanes < w_eff)
                if ok2.any():
                    def land2(ok2=ok2, te=te):
                        src2 = self.dudv_g[b, te]
                        self.prev[b, te % DP][:32][ok2] = src2[ok2]
                        self.prev_id[b, te % DP][:32][ok2] = te
                    grp.append(land2)
                if has_below and 0 <= te - 32 < w:
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
        if k0 and row_ok[0]:
            def land0():
                self.prev[b, 0][0] = self.dudv_g[b, 0, 0]
                self.prev_id[b, 0][0] = 0
            g0.append(land0)
        for tp in range(D):
            grp = g0 if tp == 0 else []
            issue(tp, grp)
            commit(grp)

        du_l = np.zeros(32, f32)
        dv_l = np.zeros(32, f32)
        hl = np.zeros(32, f32)
        nxt = np.zeros((32, 2), f32)
        pr = (k - 1) * nb + b if k > 0 else 0
        for t0 in range(0, TLp, C):
            if t0 > 0:
                self.prog[wi] = t0
            need = np.where(off == NONE, 0, np.clip(t0 + off, 0, TLp))
            spins = 0
            while not (self.prog >= need).all():
                spins += 1
                if spins > 200000:
                    raise RuntimeError("deadlock: warp %d (b=%d,k=%d) at t0=%d need=%s prog=%s" % (wi, b, k, t0, need[:self.nw], self.prog[:self.nw]))
                yield
            if t0 == 0:
                if k0:
                    wait(D - 1)
                    nxt = self.prev[b, 0][:32].copy()
                    assert self.prev_id[b, 0][0] == 0 or not row_ok[0]
                else:
                    nxt = self.ring[pr, 0].copy()
                    assert self.ring_id[pr, 0] == 0
            for s in range(C):
                t = t0 + s
                i = t - lanes
                grp = []
                issue(t + D, grp)
                commit(grp)
                wait(D)
                yield
                act = row_ok & (i >= 0) & (i < w)
                # records
                rec = self.recs[wi, t % DS].copy()
                assert (self.recs_id[wi, t % DS][act] == t).all(), "record ring: stale slot"
                own = nxt
                has_l, has_r = i > 0, i + 1 < w
                if k0:
                    slot = (t + 1) % DP
                    nxt = self.prev[b, slot][:32].copy()
                    bot = self.prev[b, slot][1:33].copy()
                    chk = act & has_r
                    assert (self.prev_id[b, slot][:32][chk] == t + 1).all(), "prev ring: right neighbour stale"
                    chk = act & ~last_row
                    assert (self.prev_id[b, slot][1:33][chk] == t + 1).all(), "prev ring: bottom neighbour stale"
                else:
                    slot = (t + 1) % R
                    nxt = self.ring[pr, slot].copy()
                    if (act & (has_r | ~last_row))[:31].any() or (act & has_r)[31]:
                        assert self.ring_id[pr, slot] == t + 1, "ring (b,k-1): entry t+1 stale (%d)" % self.ring_id[pr, slot]
                    bot = np.empty((32, 2), f32)
                    bot[:31] = nxt[1:]
                    if has_below:
                        bot[31] = self.ring[pr + 1, slot, 0]
                        if act[31] and not last_row[31]:
                            assert self.ring_id[pr + 1, slot] == t - 31, "ring (b+1,k-1): entry t-31 stale"
                    else:
                        bot[31] = nxt[31]
                self.reads_checked += int(act.sum())
                top_u = np.concatenate(([du_l[0]], du_l[:31]))
                top_v = np.concatenate(([dv_l[0]], dv_l[:31]))
                if has_above and t < w:
                    slot_t = (t + 31) % R
                    top_u[0], top_v[0] = self.ring[wi - 1, slot_t, 31]
                    if act[0]:
                        assert self.ring_id[wi - 1, slot_t] == t + 31, "ring (b-1,k): entry t+31 stale"
                with np.errstate(all="ignore"):
                    a11, a12, a22, b1, b2, hh, vv, vt = (rec[:, q] for q in range(8))
                    du_r = np.where(has_r, nxt[:, 0], f32(0))
                    dv_r = np.where(has_r, nxt[:, 1], f32(0))
                    t1u, t1v = hh * du_r, hh * dv_r
                    t2u, t2v = t1u + vt * top_u, t1v + vt * top_v
                    bsu, bsv = np.where(first_row, t1u, t2u), np.where(first_row, t1v, t2v)
                    t3u, t3v = bsu + vv * bot[:, 0], bsv + vv * bot[:, 1]
                    s1 = np.where(last_row, bsu, t3u) + b1
                    s2 = np.where(last_row, bsv, t3v) + b2
                    B1w, B2w = hl * du_l + s1, hl * dv_l + s2
                    B1, B2 = np.where(has_l, B1w, s1), np.where(has_l, B2w, s2)
                    du = own[:, 0] + om * (a11 * B1 + a12 * B2 - own[:, 0])
                    dv = own[:, 1] + om * (a12 * B1 + a22 * B2 - own[:, 1])
                hl = hh.copy()
                du_l, dv_l = du.astype(f32), dv.astype(f32)
                self.ring[wi, t % R, :, 0] = du_l
                self.ring[wi, t % R, :, 1] = dv_l
                self.ring_id[wi, t % R] = t
                if klast:
                    for l in np.nonzero(act)[0]:
                        self.dudv_g[b, t, l] = (du_l[l], dv_l[l])
        self.prog[wi] = TLp

    def run(self, omega):
        gens = [self.warp(wi, omega) for wi in range(self.nw)]
        alive = list(range(self.nw))
        while alive:
            # random interleaving with bursts: a warp runs 1..12 scheduling points in a row
            wi = alive[self.rng.integers(len(alive))]
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
    cases = [(20, 14, 3), (33, 28, 3), (40, 56, 3), (17, 70, 2), (64, 33, 1), (9, 100, 3), (50, 64, 4), (37, 32, 3), (5, 40, 3)]
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
