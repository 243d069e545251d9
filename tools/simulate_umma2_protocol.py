"""Randomised protocol simulation (pure Python, CPU) of the mbarrier / TMEM hand-offs in
prefill_umma2_kernel (nano-vllm-ascend_b200/csrc/prefill_umma2.cu).

Agents: K producer lane, V producer lane, MMA issuer, the in-order tensor pipe, two softmax
warpgroups (modelled as 128 -> 4 threads each). Barriers follow PTX mbarrier semantics
(try_wait.parity(P) is true iff the phase with parity P has completed). Every agent is a generator
that yields the condition it is blocked on; a random scheduler picks any runnable agent, TMA and MMA
completions happen after random delays. The model carries the CONTENT of every shared resource
(K/V stages, S/P/O regions of TMEM) and asserts at each use that it is what the consumer expects,
and the run must terminate (no deadlock) for every (n_tiles[0], n_tiles[1]) combination.

    python tools/simulate_umma2_protocol.py          # prints "ok" lines; raises on a violation
"""
from __future__ import annotations

import random

STAGES = 2
WG = 4          # threads per modelled softmax warpgroup (128 in the kernel)


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, n_tiles, rng):
        self.nt = n_tiles
        self.NT = max(n_tiles)
        self.rng = rng
        self.q_full = Bar(1)
        self.k_full = [Bar(1) for _ in range(STAGES)]
        self.v_full = [Bar(1) for _ in range(STAGES)]
        self.k_empty = [Bar(1) for _ in range(STAGES)]
        self.v_empty = [Bar(1) for _ in range(STAGES)]
        self.s_full = [Bar(1), Bar(1)]
        self.p_full = [Bar(WG), Bar(WG)]
        self.o_full = [Bar(1), Bar(1)]
        self.kst = [None] * STAGES          # content of K stages
        self.vst = [None] * STAGES
        self.S = [None, None]               # ('S', j) / ('P', j, n_threads_written) / None
        self.O = [-1, -1]                   # last PV accumulated into O[t]
        self.pv_issued = [-1, -1]
        self.pipe = []                      # in-order tensor pipe: callables
        self.async_ops = []                 # (fire_time, callable) TMA completions
        self.time = 0
        self.reads_left = {}                # (t, j) -> threads that still have to read S

    # ---- agents -------------------------------------------------------------------------
    def producer(self, is_k):
        full, empty, st = (self.k_full, self.k_empty, self.kst) if is_k else (self.v_full, self.v_empty, self.vst)
        if is_k:
            self.async_ops.append((self.time + self.rng.randint(1, 5), self.q_full.arrive))
        for j in range(self.NT):
            s = j % STAGES
            while not empty[s].done(((j // STAGES) & 1) ^ 1):
                yield
            assert st[s] is None or st[s][0] == "free", f"overwriting live stage {st[s]}"
            st[s] = ("loading", j)

            def land(s=s, j=j):
                st[s] = ("data", j)
                full[s].arrive()
            self.async_ops.append((self.time + self.rng.randint(1, 8), land))
            yield

    def mma(self):
        def issue_s(t, j):
            def run():
                assert self.kst[j % STAGES] == ("data", j), (self.kst, j)
                assert self.S[t] is None or self.S[t] == ("consumed", j - 1), (self.S[t], t, j)
                self.S[t] = ("S", j)
                self.reads_left[(t, j)] = WG
            self.pipe.append(run)
            self.pipe.append(self.s_full[t].arrive)           # commit

        def free_k(s):
            def run():
                self.kst[s] = ("free",)
                self.k_empty[s].arrive()
            return run

        def free_v(s):
            def run():
                self.vst[s] = ("free",)
                self.v_empty[s].arrive()
            return run

        while not self.q_full.done(0):
            yield
        if self.NT > 0:
            while not self.k_full[0].done(0):
                yield
            for t in range(2):
                if self.nt[t] > 0:
                    issue_s(t, 0)
            self.pipe.append(free_k(0))
            yield
        for j in range(self.NT):
            s = j % STAGES
            while not self.v_full[s].done((j // STAGES) & 1):
                yield
            more = j + 1 < self.NT
            if more:
                while not self.k_full[(j + 1) % STAGES].done(((j + 1) // STAGES) & 1):
                    yield
            for t in range(2):
                if j < self.nt[t]:
                    while not self.p_full[t].done(j & 1):
                        yield

                    def pv(t=t, j=j, s=s):
                        assert self.vst[s] == ("data", j), (self.vst, j)
                        assert self.S[t] == ("P", j, WG), (self.S[t], t, j)
                        assert self.O[t] == j - 1
                        self.O[t] = j
                        self.S[t] = ("consumed", j)
                    self.pv_issued[t] = j
                    self.pipe.append(pv)
                    self.pipe.append(self.o_full[t].arrive)
                    if j + 1 < self.nt[t]:
                        issue_s(t, j + 1)
                    yield
            self.pipe.append(free_v(s))
            if more:
                self.pipe.append(free_k((j + 1) % STAGES))
            yield

    def softmax(self, t, tid):
        for j in range(self.nt[t]):
            while not self.s_full[t].done(j & 1):
                yield
            # the row is read (tcgen05.ld) before anything is written
            cur = self.S[t]
            assert cur == ("S", j) or (cur[0] == "P" and cur[1] == j), (cur, t, j)
            self.reads_left[(t, j)] -= 1
            yield
            if j > 0 and self.rng.random() < 0.4:             # lazy rescale of O[t]
                while not self.o_full[t].done((j - 1) & 1):
                    yield
                assert self.O[t] == j - 1 and self.pv_issued[t] == j - 1, (self.O[t], self.pv_issued[t], j)
                yield
            # write P over S (own lane only): other threads may still be reading THEIR lanes
            cur = self.S[t]
            n = cur[2] + 1 if cur[0] == "P" else 1
            self.S[t] = ("P", j, n)
            yield
            self.p_full[t].arrive()
        if self.nt[t] > 0:                                     # epilogue
            while not self.o_full[t].done((self.nt[t] - 1) & 1):
                yield
            assert self.O[t] == self.nt[t] - 1

    # ---- scheduler ----------------------------------------------------------------------
    def run(self):
        agents = [self.producer(True), self.producer(False), self.mma()]
        agents += [self.softmax(t, i) for t in range(2) for i in range(WG)]
        alive = list(agents)
        idle = 0
        while alive:
            self.time += 1
            progressed = False
            # asynchronous completions
            for op in [o for o in self.async_ops if o[0] <= self.time]:
                self.async_ops.remove(op)
                op[1]()
                progressed = True
            if self.pipe and self.rng.random() < 0.6:          # tensor pipe retires in order
                self.pipe.pop(0)()
                progressed = True
            a = self.rng.choice(alive)
            before = self.snapshot()
            try:
                next(a)
            except StopIteration:
                alive.remove(a)
                progressed = True
            if self.snapshot() != before:
                progressed = True
            idle = 0 if progressed or self.async_ops or self.pipe else idle + 1
            assert idle < 20000, f"deadlock: n_tiles={self.nt}"
        while self.pipe:
            self.pipe.pop(0)()

    def snapshot(self):
        bars = [self.q_full, *self.k_full, *self.v_full, *self.k_empty, *self.v_empty, *self.s_full,
                *self.p_full, *self.o_full]
        return (tuple((b.phase, b.pending) for b in bars), tuple(map(str, self.S)), len(self.pipe),
                len(self.async_ops), tuple(self.pv_issued))


def main(seeds=30):
    for n0 in range(0, 7):
        for n1 in (n0, n0 + 1, 0):
            if n0 == 0 and n1 == 0:
                continue
            if n1 == 0 and n0 == 0:
                continue
            for seed in range(seeds):
                Sim((n0, n1), random.Random(seed * 977 + n0 * 13 + n1)).run()
        print(f"ok n_tiles[0]={n0}")


if __name__ == "__main__":
    main()
