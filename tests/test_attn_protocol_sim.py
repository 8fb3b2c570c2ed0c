"""CPU: discrete-event simulation of the attention kernel's mbarrier protocol (csrc/attn_tc.cu: TMA producer, MMA
issuer, four softmax warps; K / V^T ring of KST stages, single S, single P) under random latencies — for the
default single ring and for the split K / V^T rings (MDK_ATTN_SPLITKV, written without a GPU).  The three warp roles
below mirror the CUDA control flow statement by statement (barrier, parity, commit); the simulation asserts that no
role ever deadlocks, that every MMA consumes the tile it is meant to, and that TMA never overwrites a stage that a
pending MMA still reads.  It checks the PROTOCOL (the class of bug a blind kernel edit is most likely to contain),
not the arithmetic."""
import heapq
import random

import pytest


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0        # `phase` = index of the phase in progress

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def done(self, parity):           # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, n_tiles, kst, split, seed):
        self.n, self.kst, self.split = n_tiles, kst, split
        self.rng = random.Random(seed)
        self.now, self.q, self.seq = 0.0, [], 0
        B = Barrier
        self.kv_full = [B(1) for _ in range(kst)]
        self.kv_empty = [B(1) for _ in range(kst)]
        self.v_full = [B(1) for _ in range(kst)]
        self.v_empty = [B(1) for _ in range(kst)]
        self.s_full, self.s_free, self.p_full, self.pv_done = B(1), B(4), B(4), B(1)
        # data model: which tile each buffer holds, and which reads are still in flight
        self.k_tile, self.v_tile = [None] * kst, [None] * kst
        self.k_busy, self.v_busy = [0] * kst, [0] * kst            # MMAs issued but not retired reading the stage
        self.s_tile, self.p_tile, self.p_busy = None, None, 0
        self.done_roles, self.qk_done, self.pv_retired = set(), [], []

    # -- event machinery ---------------------------------------------------------------------
    def at(self, delay, fn):
        self.seq += 1
        heapq.heappush(self.q, (self.now + delay, self.seq, fn))

    def lat(self, lo, hi):
        return self.rng.uniform(lo, hi)

    def spawn(self, name, gen):
        def step():
            try:
                cond = next(gen)
            except StopIteration:
                self.done_roles.add(name)
                return
            self.blocked[name] = cond
        self.blocked = getattr(self, "blocked", {})
        self.runnable = getattr(self, "runnable", {})
        self.runnable[name] = step
        step()

    def run(self, roles):
        for name, gen in roles.items():
            self.spawn(name, gen)
        guard = 0
        while len(self.done_roles) < len(roles):
            guard += 1
            assert guard < 2_000_000, "simulation did not terminate"
            progressed = False
            for name in list(self.blocked):
                cond = self.blocked[name]
                if cond():
                    del self.blocked[name]
                    self.runnable[name]()
                    progressed = True
            if progressed:
                continue
            assert self.q, f"DEADLOCK at t={self.now:.0f}: waiting roles {sorted(self.blocked)}"
            self.now, _, fn = heapq.heappop(self.q)
            fn()

    # -- asynchronous hardware ---------------------------------------------------------------
    def tma_load(self, kind, stage, tile, bar):
        busy = self.k_busy if kind == "k" else self.v_busy
        assert busy[stage] == 0, f"TMA overwrites {kind} stage {stage} (tile {tile}) while an MMA still reads it"

        def land():
            (self.k_tile if kind == "k" else self.v_tile)[stage] = tile
            bar.arrive()                                          # complete_tx
        self.at(self.lat(200, 3000), land)

    @staticmethod
    def _joint(bar):
        """Two loads counted on ONE full barrier (a single expect_tx for K + V^T bytes): arrives when both landed."""
        state = {"left": 2}

        class _Half:
            def arrive(_s):
                state["left"] -= 1
                if state["left"] == 0:
                    bar.arrive()
        return _Half()

    def commit(self, bars, on_retire):                            # tcgen05.commit: arrives when prior MMAs retire
        def retire():
            on_retire()
            for b in bars:
                b.arrive()
        self.at(self.lat(50, 400), retire)

    # -- the three roles, mirroring attn_tc.cu -----------------------------------------------
    def producer(self):
        n, kst = self.n, self.kst
        if self.split:
            def load_k(t):
                st = t % kst
                yield lambda: self.kv_empty[st].done(((t // kst) & 1) ^ 1)
                self.tma_load("k", st, t, self.kv_full[st])
            yield from load_k(0)
            for j in range(n):
                if j + 1 < n:
                    yield from load_k(j + 1)
                st = j % kst
                yield lambda: self.v_empty[st].done(((j // kst) & 1) ^ 1)
                self.tma_load("v", st, j, self.v_full[st])
        else:
            stage, phase = 0, 0
            for j in range(n):
                yield lambda: self.kv_empty[stage].done(phase ^ 1)
                half = self._joint(self.kv_full[stage])           # one expect_tx covering K and V^T bytes
                self.tma_load("k", stage, j, half)
                self.tma_load("v", stage, j, half)
                stage += 1
                if stage == kst:
                    stage, phase = 0, phase ^ 1

    def mma(self):
        n, kst = self.n, self.kst

        def issue_s(stage, tile):
            assert self.k_tile[stage] == tile, f"QK^T of tile {tile} reads K stage {stage} holding {self.k_tile[stage]}"
            self.k_busy[stage] += 1
            bars = [self.s_full] + ([self.kv_empty[stage]] if self.split else [])

            def retire():
                self.k_busy[stage] -= 1
                self.s_tile = tile
                self.qk_done.append(tile)
            self.commit(bars, retire)
        yield lambda: self.kv_full[0].done(0)
        issue_s(0, 0)
        stage, phase = 0, 0
        for j in range(n):
            nstage, nphase = stage + 1, phase
            if nstage == kst:
                nstage, nphase = 0, phase ^ 1
            if j + 1 < n:
                yield lambda: self.kv_full[nstage].done(nphase)
                yield lambda: self.s_free.done(j & 1)
                issue_s(nstage, j + 1)
            yield lambda: self.p_full.done(j & 1)
            if self.split:
                yield lambda: self.v_full[stage].done(phase)
            assert self.p_tile == j, f"P V of tile {j} reads P holding {self.p_tile}"
            assert self.v_tile[stage] == j, f"P V of tile {j} reads V stage {stage} holding {self.v_tile[stage]}"
            self.v_busy[stage] += 1
            self.p_busy += 1
            if not self.split:
                self.k_busy[stage] += 0                             # (K of this stage was consumed by QK^T earlier)
            st = stage

            def retire(st=st, j=j):
                self.v_busy[st] -= 1
                self.p_busy -= 1
                self.pv_retired.append(j)
            self.commit([self.v_empty[st] if self.split else self.kv_empty[st], self.pv_done], retire)
            stage, phase = nstage, nphase

    def softmax(self, w):
        for j in range(self.n):
            yield lambda: self.s_full.done(j & 1)
            assert self.s_tile == j, f"softmax warp {w} drains S holding tile {self.s_tile}, expected {j}"
            t_ld = self.now + self.lat(50, 300)
            self.at(t_ld - self.now, lambda: None)                  # a timer event so that time reaches t_ld
            yield lambda: self.now >= t_ld
            self.s_free.arrive()
            if j > 0:
                yield lambda: self.pv_done.done((j - 1) & 1)
            assert self.p_busy == 0, "P is rewritten while P V of the previous tile still reads it"
            t_exp = self.now + self.lat(300, 2500)
            self.at(t_exp - self.now, lambda: None)
            yield lambda: self.now >= t_exp
            if w == 0:
                self.p_tile = j                                     # (all four warps publish before the MMA proceeds)
            self.p_full.arrive()
        yield lambda: self.pv_done.done((self.n - 1) & 1)


def _simulate(n_tiles, kst, split, seed):
    sim = Sim(n_tiles, kst, split, seed)
    roles = {"tma": sim.producer(), "mma": sim.mma()}
    for w in range(4):
        roles[f"softmax{w}"] = sim.softmax(w)
    sim.run(roles)
    assert sim.qk_done == list(range(n_tiles)) and sim.pv_retired == list(range(n_tiles))
    return sim.now


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("kst", [2, 3])
def test_attention_barrier_protocol(split, kst):
    for seed in range(40):
        for n_tiles in (1, 2, 3, 5, 9):
            _simulate(n_tiles, kst, split, seed)


def test_split_rings_shorten_the_period_when_tma_latency_dominates():
    """With loads slower than the softmax phase the single ring pays one load per tile; split rings hide it."""
    def period(split):
        tot = 0.0
        for seed in range(10):
            sim = Sim(24, 2, split, seed)
            sim.lat = lambda lo, hi, r=sim.rng: {(200, 3000): r.uniform(2400, 2600), (300, 2500): r.uniform(1000, 1100)}.get((lo, hi), r.uniform(lo, min(hi, 200)))
            roles = {"tma": sim.producer(), "mma": sim.mma()}
            for w in range(4):
                roles[f"softmax{w}"] = sim.softmax(w)
            sim.run(roles)
            tot += sim.now / 24
        return tot / 10
    single, split = period(False), period(True)
    assert single > 2400 and split < 0.75 * single, (single, split)
