"""Discrete simulation of the mbarrier protocol of a dual-issuer variant of conv_tc_kernel / conv_lin_kernel (producer, two
MMA-issuing threads with one stage ring each, epilogue) under random interleavings: checks that it never deadlocks and that
every MMA reads the stage contents of its own (tile, tap).  The variant was built, passed the GPU suite and was measured
slower than the single-issuer kernels; the first attempt, with ONE ring shared by both issuers, hung:
a parity wait is only safe while one thread waits through every phase of a barrier in order."""
import random

class Bar:
    def __init__(self, count): self.count, self.pending, self.phase = count, count, 0
    def arrive(self):
        self.pending -= 1
        if self.pending == 0: self.pending, self.phase = self.count, self.phase + 1
    def done(self, parity): return (self.phase & 1) != parity    # try_wait.parity: phase with this parity has completed

def run(kind, tiles, S, NT, NACC, seed):
    rnd = random.Random(seed)
    full = [Bar(1) for _ in range(S)]; empty = [Bar(1) for _ in range(S)]
    tfull = [Bar(1) for _ in range(NACC)]; tempty = [Bar(1) for _ in range(NACC)]
    stage = [None] * S
    inflight = []          # pending async events: (kind, payload)
    log = []
    def producer():
        if kind == "tc":
            itp = [0, 0]
            for lt in range(tiles):
                par = lt & 1
                for t in range(NT):
                    i = itp[par]; itp[par] += 1
                    s = 2 * (i % (S // 2)) + par
                    while not empty[s].done(((i // (S // 2)) & 1) ^ 1): yield
                    inflight.append(("tma", (s, (lt, t))))
                    yield
        else:
            for it in range(tiles):
                par, i = it & 1, it >> 1
                R = S >> 1
                s = 2 * (i % R) + par
                while not empty[s].done(((i // R) & 1) ^ 1): yield
                inflight.append(("tma", (s, (it, 0))))
                yield
    def issuer(par):
        if kind == "tc":
            lt = par
            while lt < tiles:
                acc = par
                i = (lt >> 1) * NT
                while not tempty[acc].done(((lt >> 1) & 1) ^ 1): yield
                for t in range(NT):
                    s = 2 * (i % (S // 2)) + par
                    while not full[s].done((i // (S // 2)) & 1): yield
                    assert stage[s] == (lt, t), ("wrong operands", stage[s], (lt, t))
                    inflight.append(("commit_empty", s))
                    i += 1
                    yield
                inflight.append(("commit_tfull", acc))
                lt += 2
        else:
            it = par
            R = S >> 1
            while it < tiles:
                acc = it % NACC
                i = it >> 1
                s = 2 * (i % R) + par
                while not tempty[acc].done(((it // NACC) & 1) ^ 1): yield
                while not full[s].done((i // R) & 1): yield
                assert stage[s] == (it, 0), ("wrong operands", stage[s], it)
                inflight.append(("commit_empty", s)); inflight.append(("commit_tfull", acc))
                it += 2
                yield
    def epilogue():
        for lt in range(tiles):
            acc = (lt & 1) if kind == "tc" else lt % NACC
            par = ((lt >> 1) & 1) if kind == "tc" else ((lt // NACC) & 1)
            while not tfull[acc].done(par): yield
            log.append(lt)
            yield
            tempty[acc].arrive()
    threads = {"prod": producer(), "i0": issuer(0), "i1": issuer(1), "epi": epilogue()}
    steps = 0
    while threads:
        steps += 1
        if steps > 200000: return "deadlock", log
        # retire a random async event sometimes
        if inflight and rnd.random() < 0.5:
            k, payload = inflight.pop(rnd.randrange(len(inflight)) if rnd.random() < 0.3 else 0)
            if k == "tma": s, tag = payload; stage[s] = tag; full[s].arrive()
            elif k == "commit_empty": empty[payload].arrive()
            else: tfull[payload].arrive()
            continue
        name = rnd.choice(list(threads))
        try: next(threads[name])
        except StopIteration: del threads[name]
    while inflight:
        k, payload = inflight.pop(0)
    assert log == list(range(tiles)), log
    return "ok", log

bad = 0
for seed in range(300):
    for tiles in (1, 2, 3, 5, 8, 13):
        r = run("tc", tiles, 8, 9, 2, seed)
        if r[0] != "ok": bad += 1; print("tc", tiles, seed, r[0])
        r = run("tc", tiles, 8, 8, 2, seed)
        if r[0] != "ok": bad += 1; print("tc8", tiles, seed, r[0])
        for S in (2, 3, 4, 5, 6, 7, 8):
            r = run("lin", tiles, S, 1, 4, seed)
            if r[0] != "ok": bad += 1; print("lin", tiles, S, seed, r[0])
print("failures:", bad)
