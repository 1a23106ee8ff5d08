"""A CPU model of the barrier protocol of the fused forward kernel (csrc/fused_sa.cu, TF32 path).

The kernel's roles — 8 producer warps, the 32 producer threads that own the (dp, 0) column, the MMA thread, the dp
warp, the epilogue — talk through mbarriers only: full[s] / empty[s] per ring slot, acc_full[b] / acc_empty[b] per
TMEM accumulator, dp_ready[b] / dp_free[b] per staging buffer.  A mistake there does not give wrong numbers, it gives
a kernel that never ends, so the protocol is restated here with the hardware's semantics — an mbarrier completes a
phase when its pending-arrival count reaches zero, a waiter can only ask "has the phase with parity P completed?" —
and run under thousands of random interleavings (including arbitrarily late asynchronous arrivals: cp.async
completions, tcgen05.commit) for every ring depth and chunk count the launcher can produce.  Checked: no deadlock;
the MMA never reads a slot before every writer of that chunk has finished; no writer touches a slot the MMA has not
released; the staging buffers are never overwritten before they are drained.

The last test feeds the model the protocol the kernel had for one afternoon of round 2 — the dp warp writing into
the ring and waiting on a slot's `empty` barrier by parity while visiting it only on some passes — and expects the
model to catch it (it hung on the GPU at C = 128, 4 ring slots)."""
import random

import pytest


class MBar:
    """mbarrier: `count` arrivals complete a phase; try_wait(P) is true once the phase of parity P is over"""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        assert n <= self.pending, "more arrivals than the phase expects"
        self.pending -= n
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def done(self, parity):
        return self.phase != parity


class Model:
    PW = 8                    # producer warps; 28 of a warp's lanes copy with cp.async in the dp chunk, 4 own the dp column

    def __init__(self, S, nchunks, nitems, rng, broken_dp_warp=False, mutation=None):
        self.S, self.nchunks, self.nitems, self.rng, self.broken = S, nchunks, nitems, rng, broken_dp_warp
        self.mutation = mutation              # deliberately wrong variants the model must reject (last test)
        self.kc_dp = nchunks - 1
        self.full = [MBar(256) for _ in range(S)]
        self.empty = [MBar(1) for _ in range(S)]
        self.acc_full = [MBar(1) for _ in range(2)]
        self.acc_empty = [MBar(256) for _ in range(2)]
        self.dp_ready = [MBar(32) for _ in range(2)]
        self.dp_free = [MBar(32) for _ in range(2)]
        self.slot_chunk = [None] * S          # chunk id the slot is being filled with / holds
        self.slot_parts = [0] * S             # writers finished for that chunk (8 warps [+ the dp column])
        self.slot_released = [True] * S       # the MMA has finished reading the previous contents
        self.stage = [None, None]             # item whose (dp, 0) column sits in the staging buffer
        self.stage_drained = [True, True]
        self.acc_item = [None, None]
        self.late = []                        # asynchronous completions still in flight: callables
        self.consumed = 0
        self.epilogues = 0

    # ------------------------------------------------------------------ roles (generators yield wait conditions)
    def begin_write(self, st, c):
        if self.slot_chunk[st] != c:
            assert self.slot_released[st], f"slot {st} overwritten for chunk {c} before the MMA released it"
            self.slot_chunk[st], self.slot_parts[st] = c, 0
            self.slot_released[st] = False          # belongs to the writers of chunk c from now on

    def producer_warp(self, w):
        st, pas = 0, 0
        for it in range(self.nitems):
            for kc in range(self.nchunks):
                c = it * self.nchunks + kc
                if pas > 0 and self.mutation != "producers_skip_empty":
                    yield lambda st=st, pas=pas: self.empty[st].done((pas - 1) & 1)
                self.begin_write(st, c)
                lanes = 28 if kc == self.kc_dp else 32

                def landed(st=st, c=c, lanes=lanes):                    # cp.async.mbarrier.arrive.noinc fires
                    assert self.slot_chunk[st] == c
                    self.slot_parts[st] += 1
                    self.full[st].arrive(lanes)
                self.late.append(landed)
                st += 1
                if st == self.S:
                    st, pas = 0, pas + 1

    def dp_lanes(self):
        """the 4 x 8 producer threads of the (dp, 0) column: same ring walk; they only act in the dp chunk"""
        st, pas = 0, 0
        for it in range(self.nitems):
            for kc in range(self.nchunks):
                c = it * self.nchunks + kc
                if pas > 0:
                    yield lambda st=st, pas=pas: self.empty[st].done((pas - 1) & 1)
                if kc == self.kc_dp and not self.broken:
                    sb = it & 1
                    yield lambda sb=sb, it=it: self.dp_ready[sb].done((it >> 1) & 1)
                    assert self.stage[sb] == it, "stale staging buffer"
                    self.begin_write(st, c)
                    self.slot_parts[st] += 1
                    self.stage_drained[sb] = True
                    self.dp_free[sb].arrive(32)
                    self.full[st].arrive(32)
                st += 1
                if st == self.S:
                    st, pas = 0, pas + 1

    def dp_warp(self):
        if self.broken:
            # the withdrawn protocol: write the column straight into the ring slot of the item's last chunk, waiting
            # for that slot's `empty` barrier by parity although this warp visits the slot only on some passes
            for it in range(self.nitems):
                c = it * self.nchunks + self.kc_dp
                st, pas = c % self.S, c // self.S
                if pas > 0:
                    yield lambda st=st, pas=pas: self.empty[st].done((pas - 1) & 1)
                self.begin_write(st, c)
                self.slot_parts[st] += 1
                self.full[st].arrive(32)
            return
        for it in range(self.nitems):
            sb = it & 1
            if it >= 2 and self.mutation != "dp_warp_skips_free":
                yield lambda sb=sb, it=it: self.dp_free[sb].done(((it >> 1) - 1) & 1)
            assert self.stage_drained[sb], "staging buffer overwritten before it was drained"
            self.stage[sb], self.stage_drained[sb] = it, False
            self.dp_ready[sb].arrive(32)

    def mma(self):
        st, ph = 0, 0
        for it in range(self.nitems):
            buf = it & 1
            yield lambda buf=buf, it=it: self.acc_empty[buf].done(((it >> 1) & 1) ^ 1)
            assert self.acc_item[buf] is None, "accumulator overwritten before the epilogue drained it"
            for kc in range(self.nchunks):
                c = it * self.nchunks + kc
                yield lambda st=st, ph=ph: self.full[st].done(ph)
                assert self.slot_chunk[st] == c, f"MMA expected chunk {c} in slot {st}, found {self.slot_chunk[st]}"
                assert self.slot_parts[st] == self.PW + (1 if kc == self.kc_dp else 0), "chunk read before every writer finished"
                self.consumed += 1

                def committed(st=st, buf=buf, it=it, last=(kc == self.nchunks - 1)):   # tcgen05.commit arrives later
                    self.slot_released[st] = True
                    self.empty[st].arrive()
                    if last:
                        self.acc_item[buf] = it
                        self.acc_full[buf].arrive()
                self.late.append(committed)
                st += 1
                if st == self.S:
                    st, ph = 0, ph ^ 1

    def epilogue(self):
        for it in range(self.nitems):
            buf = it & 1
            yield lambda buf=buf, it=it: self.acc_full[buf].done((it >> 1) & 1)
            assert self.acc_item[buf] == it
            self.acc_item[buf] = None
            self.epilogues += 1
            self.acc_empty[buf].arrive(256)

    # ------------------------------------------------------------------ scheduler
    def run(self, max_steps=2_000_000):
        # acc_empty starts "completed" for the first use of each buffer: the kernel waits for parity 1 there
        agents = [self.producer_warp(w) for w in range(self.PW)] + [self.dp_lanes(), self.dp_warp(), self.mma(), self.epilogue()]
        waiting = [None] * len(agents)
        alive = [True] * len(agents)
        for _ in range(max_steps):
            runnable = [i for i in range(len(agents)) if alive[i] and (waiting[i] is None or waiting[i]())]
            choices = len(runnable) + (1 if self.late else 0)
            if choices == 0:
                if not any(alive):
                    return "done"
                return "deadlock"
            pick = self.rng.randrange(choices)
            if pick == len(runnable):                      # one asynchronous completion lands (any of them: no order)
                self.late.pop(self.rng.randrange(len(self.late)))()
                continue
            i = runnable[pick]
            try:
                waiting[i] = next(agents[i])
            except StopIteration:
                alive[i], waiting[i] = False, None
        return "timeout"


def _launcher_configs():
    """(ring depth, chunks per item) as launch_fused_fwd derives them from C (TF32 path, 212 KB ring budget)"""
    out = set()
    for C in list(range(8, 257, 8)) + [264, 512, 1024, 2048]:
        nchunks = (C + 8 + 31) // 32
        for nslices in (1, 2):
            resident = nslices == 1 and nchunks * 16 + 3 * 32 <= 212
            fixed, slot = (nchunks * 16, 32) if resident else (0, 48)
            out.add((min(8, (212 - fixed) // slot), nchunks))
    return sorted(out)


@pytest.mark.parametrize("S,nchunks", _launcher_configs())
def test_fused_forward_protocol_never_deadlocks_or_reads_early(S, nchunks):
    assert S >= 2
    for seed in range(12):
        rng = random.Random(1000 * S + 10 * nchunks + seed)
        nitems = rng.choice((1, 2, 3, 5, 8))
        m = Model(S, nchunks, nitems, rng)
        assert m.run() == "done", (S, nchunks, nitems, seed)
        assert m.consumed == nitems * nchunks and m.epilogues == nitems


def test_ring_depths_and_chunk_counts_beyond_the_launcher():
    rng = random.Random(5)
    for _ in range(150):
        S, nchunks, nitems = rng.randrange(2, 9), rng.randrange(1, 40), rng.randrange(1, 7)
        m = Model(S, nchunks, nitems, rng)
        assert m.run() == "done", (S, nchunks, nitems)


def test_the_model_catches_the_withdrawn_dp_warp_protocol():
    """dp warp writing into the ring and waiting on `empty` by parity: hung on the GPU with 5 chunks per item in 4 slots"""
    failures = 0
    for seed in range(40):
        m = Model(4, 5, 6, random.Random(seed), broken_dp_warp=True)
        try:
            failures += m.run() != "done"
        except AssertionError:
            failures += 1
    assert failures >= 20, failures
    # and two further deliberately wrong variants: producers that do not wait for `empty`, a dp warp that does not
    # wait for `dp_free`
    for mutation, cfg in (("producers_skip_empty", (3, 7, 4)), ("dp_warp_skips_free", (4, 2, 8))):
        failures = 0
        for seed in range(30):
            m = Model(*cfg, random.Random(seed), mutation=mutation)
            try:
                failures += m.run() != "done"
            except AssertionError:
                failures += 1
        assert failures >= 10, (mutation, failures)
