"""Discrete-event model of the attention kernels' mbarrier protocol (v6 = shipped, v7 = tools/experimental candidate).

Checks, for many tile counts and random latencies, that the protocol (a) never deadlocks and (b) never lets two agents touch
the same TMEM / smem buffer at the same time in conflicting ways (S overwritten before the softmax read it, P overwritten
before its PV executed, a K/V stage refilled while an MMA still reads it, O rescaled while a PV accumulates into it).
This is the class of bug that no compiler sees and that costs GPU minutes to find (profiles/r01_attention_tc_v2_notes.txt).

    python tools/experimental/mbar_sim.py            # v6 and v7, n_tiles 1..40, 200 random schedules each
"""
import heapq
import random
import sys


class Barrier:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, 'too many arrivals'
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):            # mbarrier.try_wait.parity: true once the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.now, self.events, self.seq = 0.0, [], 0
        self.agents, self.blocked = [], {}
        self.tensor_free = 0.0           # in-order tensor pipe
        self.busy = {}                   # buffer -> list of (t0, t1, mode, who)

    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def use(self, buf, t0, t1, mode, who):
        for a0, a1, m, w in self.busy.get(buf, []):
            if a0 < t1 and t0 < a1 and ('w' in (m, mode)):
                raise AssertionError(f'hazard on {buf}: {who} [{t0:.0f},{t1:.0f}] {mode} vs {w} [{a0:.0f},{a1:.0f}] {m}')
        self.busy.setdefault(buf, []).append((t0, t1, mode, who))

    def spawn(self, gen, name):
        self.agents.append(name)
        self.step(gen, name)

    def step(self, gen, name):
        try:
            req = next(gen)
        except StopIteration:
            self.agents.remove(name)
            return
        kind = req[0]
        if kind == 'wait':
            _, bar, parity = req
            if bar.passed(parity):
                self.at(self.now + self.rng.uniform(5, 60), lambda: self.step(gen, name))
            else:
                self.blocked[name] = (gen, bar, parity)
        elif kind == 'sleep':
            self.at(self.now + req[1], lambda: self.step(gen, name))

    def poll(self):
        for name, (gen, bar, parity) in list(self.blocked.items()):
            if bar.passed(parity):
                del self.blocked[name]
                self.at(self.now + self.rng.uniform(20, 300), lambda g=gen, n=name: self.step(g, n))   # wake-up latency

    def run(self):
        while self.events:
            self.now, _, fn = heapq.heappop(self.events)
            fn()
            self.poll()
        if self.agents or self.blocked:
            raise AssertionError(f'deadlock: still alive {self.agents}, blocked {list(self.blocked)}')

    # ---- async engines
    def tma(self, bufs, bar):
        t0, t1 = self.now, self.now + self.rng.uniform(400, 2500)
        for b in bufs:
            self.use(b, t0, t1, 'w', 'tma')
        self.at(t1, bar.arrive)

    def mma(self, thread, reads, writes, dur, who):
        t0 = max(self.now, self.tensor_free)
        t1 = t0 + dur
        self.tensor_free = t1
        for b in reads:
            self.use(b, t0, t1, 'r', who)
        for b in writes:
            self.use(b, t0, t1, 'w', who)
        thread['last'] = max(thread.get('last', 0.0), t1)

    def commit(self, thread, bar):
        self.at(max(self.now, thread.get('last', 0.0)) + self.rng.uniform(10, 80), bar.arrive)


def attention(sim, n_tiles, version):
    """version 6: 64-key tiles, one S buffer per group, one MMA thread.  version 7: 32-key tiles, two S buffers per group,
    one MMA thread per group, pv_done barriers for the lazy rescale."""
    rng = sim.rng
    stages = 5 if version == 6 else 10
    q_full = Barrier(1)
    kv_full = [Barrier(1) for _ in range(stages)]
    kv_empty = [Barrier(1) for _ in range(stages)]
    nbuf = 1 if version == 6 else 2
    s_full = [[Barrier(1) for _ in range(nbuf)] for _ in range(2)]
    p_full = [[Barrier(4) for _ in range(nbuf)] for _ in range(2)]
    pv_done = [[Barrier(1) for _ in range(2)] for _ in range(2)]
    done = [Barrier(1), Barrier(1)]
    my_tiles = [max(0, (n_tiles - g + 1) // 2) if n_tiles > g else 0 for g in range(2)]

    def producer():
        sim.tma(['Q'], q_full)
        s, ph = 0, 0
        for j in range(n_tiles):
            yield ('wait', kv_empty[s], ph ^ 1)
            sim.tma([f'K{s}', f'V{s}'], kv_full[s])
            yield ('sleep', rng.uniform(10, 40))
            s += 1
            if s == stages:
                s, ph = 0, ph ^ 1

    def mma_thread(groups):
        th = {}

        def qk(g, it):
            j = g + 2 * it
            s = j % stages
            yield ('wait', kv_full[s], (j // stages) & 1)
            sim.mma(th, ['Q', f'K{s}'], [f'S{g}{it % nbuf}q{q}' for q in range(4)], 128 if version == 6 else 100, f'QK{j}')
            sim.commit(th, s_full[g][it % nbuf])
            yield ('sleep', rng.uniform(20, 120))

        yield ('wait', q_full, 0)
        for g in groups:
            for it in range(min(nbuf if version == 7 else 1, my_tiles[g])):
                yield from qk(g, it)
        order = sorted((g + 2 * it, g, it) for g in groups for it in range(my_tiles[g]))
        for j, g, it in order:
            s = j % stages
            yield ('wait', p_full[g][it % nbuf], (it // nbuf) & 1)
            sim.mma(th, [f'S{g}{it % nbuf}q{q}' for q in range(4)] + [f'V{s}'], [f'O{g}q{q}' for q in range(4)],
                    128 if version == 6 else 64, f'PV{j}')
            if version == 7:
                sim.commit(th, pv_done[g][it & 1])
            sim.commit(th, kv_empty[s])
            yield ('sleep', rng.uniform(20, 120))
            if it + nbuf < my_tiles[g]:
                yield from qk(g, it + nbuf)
        for g in groups:
            sim.commit(th, done[g])

    def softmax_warp(g, w):
        for it in range(my_tiles[g]):
            b = it % nbuf
            yield ('wait', s_full[g][b], (it // nbuf) & 1)
            t0 = sim.now
            dur = rng.uniform(300, 1500)
            rescale = it > 0 and rng.random() < 0.3
            if rescale:
                if version == 7:
                    yield ('wait', pv_done[g][(it - 1) & 1], ((it - 1) >> 1) & 1)
                sim.use(f'O{g}q{w}', sim.now, sim.now + 50, 'w', f'rescale g{g} w{w} it{it}')
                yield ('sleep', 50)
            sim.use(f'S{g}{b}q{w}', t0, sim.now + dur, 'r', f'softmax g{g} w{w} it{it}')  # reads its rows of S ...
            yield ('sleep', dur)
            sim.use(f'S{g}{b}q{w}', sim.now, sim.now + 20, 'w', f'P store g{g} w{w} it{it}')   # ... then overwrites their head with P
            yield ('sleep', 20)
            p_full[g][b].arrive()
        yield ('wait', done[0], 0)
        yield ('wait', done[1], 0)
        sim.use(f'O0q{w}', sim.now, sim.now + 30, 'r', f'epilogue g{g} w{w}')
        sim.use(f'O1q{w}', sim.now, sim.now + 30, 'r', f'epilogue g{g} w{w}')

    sim.spawn(producer(), 'producer')
    if version == 6:
        sim.spawn(mma_thread([0, 1]), 'mma')
    else:
        sim.spawn(mma_thread([0]), 'mma0')
        sim.spawn(mma_thread([1]), 'mma1')
    for g in range(2):
        for w in range(4):
            sim.spawn(softmax_warp(g, w), f'softmax{g}{w}')
    sim.run()


def main():
    for version in (6, 7):
        for n_tiles in range(1, 41):
            for seed in range(200):
                sim = Sim(seed * 1000 + n_tiles)
                try:
                    attention(sim, n_tiles, version)
                except AssertionError as e:
                    print(f'v{version} n_tiles={n_tiles} seed={seed}: {e}')
                    sys.exit(1)
        print(f'v{version}: ok (n_tiles 1..40 x 200 random schedules: no deadlock, no buffer hazard)')


if __name__ == '__main__':
    main()
