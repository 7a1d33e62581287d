"""CPU test: the encoder kernel's SNR-offset search evaluates the next three LIKELY probes of the reference's search per
pass and replays the reference's state machine on what it finds (csrc/ac3_encode.cu, stage E3; reference
src/ac3enc/ac3enc.cpp:921-967).  This model pins the control logic: for any bits-left function - monotone or not - the
batched search ends in the same (csnroffst, fsnroffst, failed) as the probe-by-probe search, in fewer passes."""
import random


class Search:
    def __init__(self, cs):
        self.phase, self.cs, self.fs, self.pcs, self.pfs, self.done, self.failed = 0, cs, 0, cs, 0, 0, 0

    def copy(self):
        t = Search(0)
        t.__dict__.update(self.__dict__)
        return t


def search_step(q, left):
    """one probe result into the state machine (search_step of the kernel)"""
    ph = q.phase
    if ph == 0:
        if left >= 0:
            q.cs = q.pcs
            ph = 1
        else:
            q.pcs -= 4
            if q.pcs < 0:
                q.failed, q.cs, q.fs, q.phase, q.done = 1, 0, 0, 5, 1
            return
    elif left >= 0:
        q.cs, q.fs = q.pcs, q.pfs
    else:
        ph += 1
    while True:
        if ph == 1:
            if q.cs + 4 <= 63:
                q.pcs, q.pfs = q.cs + 4, 0
                break
            ph = 2
        elif ph == 2:
            if q.cs + 1 <= 63:
                q.pcs, q.pfs = q.cs + 1, 0
                break
            ph = 3
        elif ph == 3:
            if q.fs + 4 <= 15:
                q.pcs, q.pfs = q.cs, q.fs + 4
                break
            ph = 4
        elif ph == 4:
            if q.fs + 1 <= 15:
                q.pcs, q.pfs = q.cs, q.fs + 1
                break
            ph = 5
        else:
            q.done = 1
            break
    q.phase = ph


def probe_by_probe(f, cs0):
    q, n = Search(cs0), 0
    while not q.done:
        search_step(q, f(q.pcs, q.pfs))
        n += 1
    return (q.cs, q.fs, q.failed), n


def three_at_a_time(f, cs0):
    q, bank, passes = Search(cs0), [], 0
    while True:
        t, cand = q.copy(), []
        for _ in range(3):
            if t.done:
                break
            cand.append(t.pcs * 16 + t.pfs)
            search_step(t, 0 if (t.phase == 0 or t.phase >= 3) else -1)       # the likely outcome
        while len(cand) < 3:
            cand.append(cand[-1])
        passes += 1
        bank = (bank + [(k, f(k >> 4, k & 15)) for k in cand])[-6:]          # this pass and the previous one
        while not q.done:
            key = q.pcs * 16 + q.pfs
            hit = [left for k, left in bank if k == key]
            if not hit:
                break
            search_step(q, hit[-1])
        if q.done:
            return (q.cs, q.fs, q.failed), passes


def test_batched_search_equals_the_reference_search_for_any_bits_left_function():
    rng = random.Random(1)
    for _ in range(4000):
        thr, noise = rng.uniform(-50, 1100), rng.choice([0, 0, 0, 30, 200])
        tab = {}

        def f(cs, fs):
            if (cs, fs) not in tab:
                tab[(cs, fs)] = thr - (cs * 16 + fs) + rng.uniform(-noise, noise)
            return tab[(cs, fs)]
        cs0 = rng.randint(0, 63)
        assert probe_by_probe(f, cs0)[0] == three_at_a_time(f, cs0)[0]


def test_batched_search_takes_three_passes_from_a_warm_start():
    rng = random.Random(2)
    probes = passes = 0
    for _ in range(1000):
        thr = rng.uniform(100, 900)

        def f(cs, fs):
            return thr - (cs * 16 + fs)
        warm = probe_by_probe(f, 40)[0][0]
        a, n = probe_by_probe(f, warm)
        b, p = three_at_a_time(f, warm)
        assert a == b
        probes += n
        passes += p
    assert passes / 1000 <= 3.5 < probes / 1000
