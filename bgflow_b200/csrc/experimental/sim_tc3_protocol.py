#!/usr/bin/env python
"""Abstract model of the mbarrier protocol of experimental/bgx_coupling_tc3.cu (one CTA): producer,
MMA issuer and eight epilogue warps as coroutines over phase-counting barriers, run under a random
scheduler with random completion latencies.  Checks (a) no deadlock, (b) no barrier is waited on with
a stale parity (a waiter never falls two phases behind), (c) no accumulator half / weight slot /
operand buffer is overwritten before its consumers are done, (d) every tile's dims are evaluated
exactly once.  It validates the PROTOCOL, not the arithmetic or the PTX.

    python bgflow_b200/csrc/experimental/sim_tc3_protocol.py            # a few thousand random schedules
"""

import random
import sys


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def done(self, parity, waiter_phase):
        """mbarrier.try_wait.parity semantics: true once the phase with this parity has completed.
        `waiter_phase` (how many completions the waiter has consumed) lets the model flag aliasing."""
        assert self.phase - waiter_phase <= 1, f"{self.name}: waiter is {self.phase - waiter_phase} phases behind"
        assert parity == (waiter_phase & 1)
        return self.phase > waiter_phase


class Sim:
    def __init__(self, n_tiles, n_hidden_ktiles, nhalf, seed):
        self.rng = random.Random(seed)
        self.n_tiles, self.kt, self.nhalf = n_tiles, list(n_hidden_ktiles), nhalf
        self.L1 = len(self.kt)                       # number of hidden layers (L - 1)
        B = Barrier
        self.full = [B("full0", 1), B("full1", 1)]
        self.x_ready, self.a_ready = B("x_ready", 8), B("a_ready", 8)
        self.hfull = B("hfull", 1)
        self.afull = [B("afull0", 1), B("afull1", 1)]
        self.aempty = [B("aempty0", 8), B("aempty1", 8)]
        self.async_events = []                       # (fire_time, callable): TMA landings, MMA completions
        self.time = 0
        # resource ownership for the overwrite checks
        self.slot_busy = [False, False]              # weight slot holds data the MMAs have not consumed
        self.acc_state = ["free", "free"]            # free | mma | full (waiting for the epilogue pulls)
        self.acc_pulled = [0, 0]
        self.evaluated = [[0] * (2 * nhalf) for _ in range(n_tiles)]
        self.mma_queue_done = 0                      # in-order completion of committed MMA groups

    # ---- helpers for asynchronous completions
    def later(self, fn, lo=1, hi=40):
        self.async_events.append((self.time + self.rng.randint(lo, hi), fn))

    def wait(self, bar, counter):
        """Coroutine: block until bar's phase `counter` completed; returns nothing, caller increments."""
        while not bar.done(counter & 1, counter):
            yield

    # ---- roles
    def producer(self):
        slot, filled, released, events = 0, 0, 0, 0
        ph_h, ph_a = 0, [0, 0]
        units = self.L1 + self.nhalf

        def wait_slot():
            nonlocal released, events, ph_h
            while filled - released >= 2:
                u = events % units
                if u < self.L1:
                    if self.hfull.done(ph_h & 1, ph_h):
                        ph_h += 1
                        released += self.kt[u]
                        events += 1
                        continue
                else:
                    b = (u - self.L1) & 1
                    if self.afull[b].done(ph_a[b] & 1, ph_a[b]):
                        ph_a[b] += 1
                        released += 1
                        events += 1
                        continue
                yield

        for _ in range(self.n_tiles):
            for l in range(self.L1):
                for _t in range(self.kt[l]):
                    yield from wait_slot()
                    assert not self.slot_busy[slot], "producer overwrites a weight slot still in use"
                    self.slot_busy[slot] = True
                    s = slot
                    self.later(lambda s=s: self.full[s].arrive())
                    filled += 1
                    slot ^= 1
            for _h in range(self.nhalf):
                yield from wait_slot()
                assert not self.slot_busy[slot], "producer overwrites a weight slot still in use"
                self.slot_busy[slot] = True
                s = slot
                self.later(lambda s=s: self.full[s].arrive())
                filled += 1
                slot ^= 1

    def commit(self, fn, slots):
        """MMAs complete in issue order some time after the commit; then their slots are free."""
        def fire():
            for s in slots:
                self.slot_busy[s] = False
            fn()
        self.later(fire, 5, 60)

    def mma(self):
        slot = 0
        ph_full, ph_x, ph_a, ph_e = [0, 0], 0, 0, [0, 0]
        outstanding = [0, 0]

        def drain(b):
            nonlocal ph_e
            if outstanding[b]:
                yield from self.wait(self.aempty[b], ph_e[b])
                ph_e[b] += 1
                outstanding[b] = 0

        for _ in range(self.n_tiles):
            for l in range(self.L1):
                if l == 0:
                    yield from self.wait(self.x_ready, ph_x)
                    ph_x += 1
                    yield from drain(0)
                    yield from drain(1)
                else:
                    yield from self.wait(self.a_ready, ph_a)
                    ph_a += 1
                assert self.acc_state == ["free", "free"], f"hidden layer overwrites a busy accumulator {self.acc_state}"
                self.acc_state = ["mma", "mma"]
                used = []
                for _t in range(self.kt[l]):
                    yield from self.wait(self.full[slot], ph_full[slot])
                    ph_full[slot] += 1
                    used.append(slot)
                    slot ^= 1

                def done_hidden():
                    self.acc_state = ["hfull", "hfull"]
                    self.hfull.arrive()
                self.commit(done_hidden, used)
            yield from self.wait(self.a_ready, ph_a)
            ph_a += 1
            for h in range(self.nhalf):
                b = h & 1
                yield from drain(b)
                yield from self.wait(self.full[slot], ph_full[slot])
                ph_full[slot] += 1
                assert self.acc_state[b] == "free", f"half {h} overwrites accumulator {b} in state {self.acc_state[b]}"
                self.acc_state[b] = "mma"

                def done_half(b=b):
                    self.acc_state[b] = "full"
                    self.acc_pulled[b] = 0
                    self.afull[b].arrive()
                self.commit(done_half, [slot])
                slot ^= 1
                outstanding[b] = 1

    def epilogue(self, w):
        j = w >> 2
        ph_h, ph_f = 0, [0, 0]

        def stage_x():
            # (c_full / c_free of the tile I/O are as in the tc2 kernel and not modelled)
            for _ in range(self.rng.randint(0, 3)):
                yield
            self.x_ready.arrive()

        for it in range(self.n_tiles):
            if it == 0:
                yield from stage_x()
            for _l in range(self.L1):
                yield from self.wait(self.hfull, ph_h)
                ph_h += 1
                assert self.acc_state[0] in ("hfull", "free_pending"), self.acc_state
                for _ in range(self.rng.randint(0, 4)):
                    yield
                # the last of the eight arrivals frees the accumulator (a_ready completes)
                if self.a_ready.pending == 1:
                    self.acc_state = ["free", "free"]
                self.a_ready.arrive()
            for h in range(self.nhalf):
                b = h & 1
                yield from self.wait(self.afull[b], ph_f[b])
                ph_f[b] += 1
                assert self.acc_state[b] == "full", f"warp {w} pulls half {h} from accumulator in state {self.acc_state[b]}"
                d = 2 * h + j
                self.acc_pulled[b] += 1
                if self.acc_pulled[b] == 8:
                    self.acc_state[b] = "free"
                self.aempty[b].arrive()
                for _ in range(self.rng.randint(0, 6)):      # evaluation
                    yield
                if (w & 3) == 0:
                    self.evaluated[it][d] += 1
                if h == self.nhalf - 1 and it + 1 < self.n_tiles:
                    yield from stage_x()

    def run(self):
        roles = [self.producer(), self.mma()] + [self.epilogue(w) for w in range(8)]
        alive = list(range(len(roles)))
        idle_rounds = 0
        while alive:
            self.time += 1
            fired = [e for e in self.async_events if e[0] <= self.time]
            self.async_events = [e for e in self.async_events if e[0] > self.time]
            for _, fn in fired:
                fn()
            progressed = bool(fired)
            for i in self.rng.sample(alive, len(alive)):
                snapshot = self.state_key()
                try:
                    next(roles[i])
                except StopIteration:
                    alive.remove(i)
                    progressed = True
                    continue
                if self.state_key() != snapshot:
                    progressed = True
            if progressed or self.async_events:
                idle_rounds = 0
            else:
                idle_rounds += 1
                assert idle_rounds < 50, f"deadlock: roles {alive} blocked (0 = producer, 1 = MMA, 2.. = epilogue warps)"
        for it in range(self.n_tiles):
            assert all(c == 1 for c in self.evaluated[it]), f"tile {it}: dims evaluated {self.evaluated[it]}"
        return True

    def state_key(self):
        bars = self.full + [self.x_ready, self.a_ready, self.hfull] + self.afull + self.aempty
        return tuple((b.phase, b.pending) for b in bars) + (tuple(self.slot_busy), tuple(self.acc_state))


def main(n_runs=300):
    configs = [((1, 2), 17), ((1, 2), 1), ((1, 2), 2), ((1, 2), 3), ((1, 2), 16), ((2, 2, 2), 5), ((1,), 4)]
    total = 0
    for kt, nhalf in configs:
        for n_tiles in (1, 2, 3, 5):
            for seed in range(n_runs // 10):
                Sim(n_tiles, kt, nhalf, seed * 7919 + nhalf + 31 * n_tiles).run()
                total += 1
    print(f"tc3 protocol model: {total} random schedules, no deadlock / aliasing / overwrite")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 300))
