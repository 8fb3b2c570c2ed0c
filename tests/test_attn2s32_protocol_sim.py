"""CPU: discrete-event simulation of the DEFAULT long-sequence attention kernel's synchronisation protocol
(csrc/attn_2s32.cu: one TMA producer, two MMA-issuing threads — one per softmax stream —, 2 x 4 softmax warps; load
units {K_u, V^T_(u-1)} in a ring of KST stages with two-commit `u_empty`; per stream two 32-column S halves used as a
double buffer, P written back over S, `pv_done` one phase per sub-tile, single-phase `o_done`, `x_full` for the merge).

Every role below follows the CUDA control flow statement by statement (which barrier, which parity, where the commits
are); tcgen05 is modelled as the hardware defines it: MMAs issued by one thread execute in issue order, a commit arrives
on its mbarrier when everything that thread issued before it has retired; TMA loads land after a random delay.  Under
random latencies and for full, ragged and minimal key counts the simulation asserts
  * no role deadlocks and every barrier wait that passes is for the phase the code means (no parity aliasing),
  * an MMA only ever reads a K / V^T stage that holds the unit it expects, fully landed, and TMA never overwrites a
    stage before every MMA that reads its old contents has retired,
  * the softmax reads S(h) — not a stale or a newer half —, P·V(h) reads P(h) from all four warps, the next S into the
    same half starts only after that P·V retired, and the lazy O rescale never overlaps a P·V in flight,
  * the epilogue reads O, and the other stream's {m, l}, only after the writers are done.
It checks the PROTOCOL (what a kernel edit made without a GPU is most likely to break), not the arithmetic."""
import heapq
import random

import pytest

KST = 3
BKV = 128


class Barrier:
    def __init__(self, count, name):
        self.count, self.pending, self.phase, self.name = count, count, 0, name   # `phase`: index of the one in progress
        self.tx = 0                                                               # outstanding TMA transactions
        self.waiters = []                                                         # (parity, resume) of blocked roles

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than the barrier was initialised for"
        self._maybe_flip()

    def expect_tx(self, n):          # mbarrier.arrive.expect_tx: one arrival + n transactions still to land
        self.tx += n
        self.pending -= 1
        assert self.pending >= 0, self.name
        self._maybe_flip()

    def complete_tx(self):
        self.tx -= 1
        assert self.tx >= 0, self.name
        self._maybe_flip()

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count
            blocked, self.waiters = self.waiters, []
            for parity, resume in blocked:
                if self.done(parity):
                    resume()
                else:
                    self.waiters.append((parity, resume))

    def done(self, parity):          # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


def n_sub(lkv, g):                   # attn_2s32.cu: sub-tiles (32 keys) of stream g
    rem = lkv - g * 64
    if rem <= 0:
        return 0
    full, tail = divmod(rem, BKV)
    return 2 * full + (2 if tail > 32 else (1 if tail > 0 else 0))


class Sim:
    def __init__(self, lkv, seed, profile):
        self.lkv, self.rng, self.profile = lkv, random.Random(seed), profile
        self.n0 = (lkv + BKV - 1) // BKV
        self.nh = [n_sub(lkv, 0), n_sub(lkv, 1)]
        self.now, self.q, self.seq = 0.0, [], 0
        B = Barrier
        self.q_bar = B(1, "q_bar")
        self.u_full = [B(1, f"u_full{s}") for s in range(KST)]
        self.u_empty = [B(2, f"u_empty{s}") for s in range(KST)]
        self.s_full = [[B(1, f"s_full{g}{b}") for b in range(2)] for g in range(2)]
        self.p_full = [[B(4, f"p_full{g}{b}") for b in range(2)] for g in range(2)]
        self.pv_done = [B(1, f"pv_done{g}") for g in range(2)]
        self.o_done = [B(1, f"o_done{g}") for g in range(2)]
        self.x_full = B(4, "x_full")
        # ---- data model ----
        self.q_landed = False
        self.k_stage = [None] * KST          # [unit, landed]
        self.v_stage = [None] * KST
        self.k_reads_left = {}               # unit -> MMA reads of K_unit still to retire
        self.v_reads_left = {}               # unit -> MMA reads of V^T_(unit-1) still to retire
        for u in range(self.n0 + 1):
            self.k_reads_left[u] = sum(self._subs_in_tile(g, u) for g in range(2)) if u < self.n0 else 0
            self.v_reads_left[u] = sum(self._subs_in_tile(g, u - 1) for g in range(2)) if u > 0 else 0
        self.s_half = [[[("free", -1)] * 4 for _ in range(2)] for _ in range(2)]   # [g][b][quarter] -> (kind, h)
        self.pv_inflight = [0, 0]
        self.pv_retired = [0, 0]
        self.o_rmw = [0, 0]                  # softmax warps currently rescaling O_g
        self.x_written = 0
        self.finished = set()
        # in-order tensor pipe per issuing thread: time at which the last issued op retires
        self.pipe_free = [0.0, 0.0]
        self.waits = 0

    def _subs_in_tile(self, g, t):
        return max(0, min(2, self.nh[g] - 2 * t))

    # ---------------- event loop ----------------
    def lat(self, kind):
        lo, hi = self.profile[kind]
        return self.rng.uniform(lo, hi)

    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.q, (t, self.seq, fn))

    def spawn(self, name, gen):
        def step():
            try:
                req = next(gen)
            except StopIteration:
                self.finished.add(name)
                return
            if req[0] == "delay":
                self.at(self.now + req[1], step)
            else:                                     # ("wait", barrier, parity): mbar_wait
                _, bar, parity = req
                self.waits += 1

                def resume():
                    self.at(self.now + self.lat("wake"), step)
                if bar.done(parity):
                    resume()
                else:
                    bar.waiters.append((parity, resume))
        self.at(0.0, step)

    def run(self, names):
        while self.q:
            t, _, fn = heapq.heappop(self.q)
            self.now = t
            fn()
        # nothing left to happen: a role that has not finished is blocked on a barrier for ever
        assert self.finished == set(names), f"deadlock: roles that never finished: {sorted(set(names) - self.finished)}"

    # ---------------- tcgen05 model ----------------
    def mma(self, g, dur, on_start, on_retire):
        """Issue an MMA from stream g's issuing thread now: starts when the thread's earlier MMAs have retired."""
        start = max(self.now, self.pipe_free[g])
        end = start + dur
        self.pipe_free[g] = end
        self.at(start, on_start)
        self.at(end, on_retire)

    def commit(self, g, bar):
        """tcgen05.commit: arrives when everything this thread issued so far has retired."""
        self.at(max(self.now, self.pipe_free[g]), bar.arrive)

    # ---------------- roles ----------------
    def producer(self):
        def land_q():
            self.q_landed = True
            self.q_bar.complete_tx()
        self.q_bar.expect_tx(1)
        self.at(self.now + self.lat("tma"), land_q)
        st, ph = 0, 1
        for u in range(self.n0 + 1):
            yield ("wait", self.u_empty[st], ph)
            has_k, has_v = u < self.n0, u > 0
            # overwrite check: every MMA that reads the old contents must have retired
            for stage, left in ((self.k_stage, self.k_reads_left), (self.v_stage, self.v_reads_left)):
                old = stage[st]
                if old is not None:
                    assert old[1], f"unit {u}: stage {st} re-armed while unit {old[0]} was still landing"
                    assert left[old[0]] == 0, f"unit {u} overwrites stage {st}: unit {old[0]} still has {left[old[0]]} reads pending"
            self.u_full[st].expect_tx(int(has_k) + 2 * int(has_v))
            if has_k:
                self.k_stage[st] = [u, False]
                self.at(self.now + self.lat("tma"), self._lander(self.k_stage, st, u, 1))
            if has_v:
                self.v_stage[st] = [u, False, 0]
                for _ in range(2):                    # two 64-key chunks
                    self.at(self.now + self.lat("tma"), self._lander(self.v_stage, st, u, 2))
            yield ("delay", self.lat("issue"))
            st += 1
            if st == KST:
                st, ph = 0, ph ^ 1

    def _lander(self, stage, st, u, parts):
        def land():
            e = stage[st]
            assert e[0] == u, "a later unit was armed before this one landed"
            if parts == 2:
                e[2] += 1
                e[1] = e[2] == 2
            else:
                e[1] = True
            self.u_full[st].complete_tx()
        return land

    def issuer(self, g):
        nh = self.nh[g]
        if nh == 0:
            return

        def issue_s(h, st):
            b, t = h & 1, h >> 1

            def start():
                assert self.q_landed
                assert self.k_stage[st] == [t, True], f"S({g},{h}) reads K stage {st} = {self.k_stage[st]}, wants tile {t}"
                for qd in range(4):
                    kind, hh = self.s_half[g][b][qd]
                    assert kind in ("free", "consumed") and (kind == "free" or hh == h - 2), \
                        f"S({g},{h}) overwrites half {b} in state {kind}({hh})"

            def retire():
                self.k_reads_left[t] -= 1
                self.s_half[g][b] = [("S", h)] * 4
            self.mma(g, self.lat("mma_s"), start, retire)
            self.commit(g, self.s_full[g][b])

        yield ("wait", self.q_bar, 0)
        yield ("wait", self.u_full[0], 0)
        issue_s(0, 0)
        if nh > 1:
            issue_s(1, 0)
        self.commit(g, self.u_empty[0])
        yield ("delay", self.lat("issue"))
        st, ph = 1, 0
        for h in range(nh):
            b, t = h & 1, h >> 1
            if b == 0:
                yield ("wait", self.u_full[st], ph)
            yield ("wait", self.p_full[g][b], (h >> 1) & 1)

            def start(h=h, b=b, t=t, st=st):
                assert self.v_stage[st][:2] == [t + 1, True], f"PV({g},{h}) reads V stage {st} = {self.v_stage[st]}, wants unit {t + 1}"
                assert all(x == ("P", h) for x in self.s_half[g][b]), f"PV({g},{h}) reads {self.s_half[g][b]}"
                assert self.o_rmw[g] == 0, f"PV({g},{h}) accumulates while a softmax warp rescales O"
                self.pv_inflight[g] += 1

            def retire(h=h, b=b, t=t):
                self.v_reads_left[t + 1] -= 1
                self.pv_inflight[g] -= 1
                self.pv_retired[g] += 1
                self.s_half[g][b] = [("consumed", h)] * 4
            self.mma(g, self.lat("mma_pv"), start, retire)
            self.commit(g, self.pv_done[g])
            if h + 1 >= nh:
                self.commit(g, self.o_done[g])
            if h + 2 < nh:
                issue_s(h + 2, st)
            yield ("delay", self.lat("issue"))
            if b == 1 or h + 1 >= nh:
                self.commit(g, self.u_empty[st])
                st += 1
                if st == KST:
                    st, ph = 0, ph ^ 1

    def softmax(self, g, qd):
        nh = self.nh[g]
        for h in range(nh):
            b = h & 1
            yield ("wait", self.s_full[g][b], (h >> 1) & 1)
            assert self.s_half[g][b][qd] == ("S", h), f"softmax({g},{qd}) at {h} reads {self.s_half[g][b][qd]}"
            yield ("delay", self.lat("ld"))
            if h > 0 and self.rng.random() < self.profile["p_rescale"]:
                yield ("wait", self.pv_done[g], (h - 1) & 1)
                assert self.pv_inflight[g] == 0 and self.pv_retired[g] == h, \
                    f"rescale({g},{qd}) at {h}: {self.pv_retired[g]} P·V retired, {self.pv_inflight[g]} in flight"
                self.o_rmw[g] += 1
                yield ("delay", self.lat("rescale"))
                self.o_rmw[g] -= 1
            yield ("delay", self.lat("exp"))
            assert self.s_half[g][b][qd] == ("S", h)
            self.s_half[g][b][qd] = ("P", h)
            self.p_full[g][b].arrive()                 # one arrival per warp (lane 0)
        # ---- epilogue ----
        if g == 1:
            if nh > 0:
                yield ("wait", self.o_done[1], 0)
                assert self.pv_retired[1] == nh and self.pv_inflight[1] == 0, "stream 1 publishes {m, l} over a P still to be read"
                self.x_written += 1
                self.x_full.arrive()
        else:
            yield ("wait", self.o_done[0], 0)
            if self.nh[1] > 0:
                yield ("wait", self.o_done[1], 0)
                yield ("wait", self.x_full, 0)
                assert self.x_written == 4
            assert self.pv_retired[0] == nh and self.pv_retired[1] == self.nh[1] and self.pv_inflight == [0, 0], \
                f"epilogue reads O with P·V retired {self.pv_retired} of {self.nh}"


PROFILES = {
    # (lo, hi) cycles; roughly the measured kernel (profiles/r02_attn2s_ab.log), then adversarial skews
    "measured": dict(tma=(600, 1500), issue=(20, 60), wake=(30, 300), mma_s=(60, 120), mma_pv=(60, 120), ld=(100, 200),
                     exp=(400, 700), rescale=(200, 400), p_rescale=0.05),
    "slow_tma": dict(tma=(3000, 30000), issue=(20, 60), wake=(30, 300), mma_s=(60, 120), mma_pv=(60, 120), ld=(50, 100),
                     exp=(100, 200), rescale=(100, 200), p_rescale=0.3),
    "slow_tensor": dict(tma=(100, 300), issue=(5, 10), wake=(5, 20), mma_s=(500, 5000), mma_pv=(500, 5000), ld=(10, 50),
                        exp=(20, 100), rescale=(20, 2000), p_rescale=0.5),
    "slow_softmax": dict(tma=(100, 300), issue=(5, 10), wake=(5, 2000), mma_s=(10, 30), mma_pv=(10, 30), ld=(10, 3000),
                         exp=(100, 6000), rescale=(100, 3000), p_rescale=0.5),
    "chaos": dict(tma=(1, 20000), issue=(1, 3000), wake=(1, 3000), mma_s=(1, 4000), mma_pv=(1, 4000), ld=(1, 3000),
                  exp=(1, 5000), rescale=(1, 3000), p_rescale=0.4),
}


def simulate(lkv, seed, profile, cls=None):
    sim = (cls or Sim)(lkv, seed, PROFILES[profile])
    names = ["tma", "mma0", "mma1"] + [f"sm{g}{q}" for g in range(2) for q in range(4)]
    sim.spawn("tma", sim.producer())
    for g in range(2):
        sim.spawn(f"mma{g}", sim.issuer(g))
        for q in range(4):
            sim.spawn(f"sm{g}{q}", sim.softmax(g, q))
    sim.run(names)
    assert all(v == 0 for v in sim.k_reads_left.values()) and all(v == 0 for v in sim.v_reads_left.values())
    return sim


@pytest.mark.parametrize("profile", sorted(PROFILES))
@pytest.mark.parametrize("lkv", [1, 31, 33, 64, 65, 97, 128, 129, 161, 192, 193, 225, 257, 384, 385, 1024, 1055, 2304])
def test_attn2s32_barrier_protocol(lkv, profile):
    for seed in range(3):
        sim = simulate(lkv, 1000 * seed + lkv, profile)
        assert sim.pv_retired == [n_sub(lkv, 0), n_sub(lkv, 1)]


def test_n_sub_covers_every_key_exactly_once():
    for lkv in list(range(1, 700)) + [1024, 2304, 9216]:
        covered = set()
        for g in range(2):
            for h in range(n_sub(lkv, g)):
                first = (h >> 1) * BKV + g * 64 + (h & 1) * 32
                assert first < lkv                                   # the kernel's nvalid >= 1
                covered.update(range(first, min(first + 32, lkv)))
        assert covered == set(range(lkv)), lkv


class OneCommitSim(Sim):
    """`u_empty` released by stream 0 alone: stream 1's reads of the stage are not waited for."""
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.u_empty = [Barrier(1, f"u_empty{s}") for s in range(KST)]

    def commit(self, g, bar):
        if g == 1 and any(bar is x for x in self.u_empty):
            return
        super().commit(g, bar)


class EarlyRescaleSim(Sim):
    """The lazy rescale without its wait for P·V(h-1)."""
    def softmax(self, g, qd):
        for req in super().softmax(g, qd):
            if req[0] == "wait" and req[1] is self.pv_done[g]:
                continue
            yield req


class AliasedParitySim(Sim):
    """`pv_done`-style multi-phase barrier in place of the single-phase `o_done`: the epilogue of stream 0 waiting for
    stream 1's LAST P·V through a barrier that flips once per sub-tile passes on an earlier phase of the same parity."""
    def __init__(self, lkv, seed, profile):
        super().__init__(lkv, seed, dict(profile, p_rescale=0.0))      # this variant is about the epilogue only

    def commit(self, g, bar):
        if any(bar is x for x in self.pv_done):
            super().commit(g, self.o_done[g])
        elif any(bar is x for x in self.o_done):
            return
        else:
            super().commit(g, bar)

    def softmax(self, g, qd):
        for req in super().softmax(g, qd):
            if req[0] == "wait" and any(req[1] is x for x in self.o_done):
                n = self.nh[self.o_done.index(req[1])]
                yield ("wait", req[1], (n - 1) & 1)
            else:
                yield req


@pytest.mark.parametrize("cls,profile", [(OneCommitSim, "chaos"), (EarlyRescaleSim, "slow_tensor"),
                                         (AliasedParitySim, "chaos")])
def test_the_simulation_detects_a_broken_protocol(cls, profile):
    """The checker must be able to fail: each variant above removes one ingredient of the protocol."""
    with pytest.raises(AssertionError):
        for seed in range(30):
            simulate(2304, seed, profile, cls)
