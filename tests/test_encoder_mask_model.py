"""CPU test: the encoder kernel computes the masking curve of bands 0..21 without a serial lane (csrc/ac3_encode.cu,
e3_mask; reference src/ac3enc/ac3enc.cpp:262-330 walks the bands one by one).  The low-frequency compensation is a chain
of steps x -> max(x - a, c) that compose, so it is a scan; the fast / slow leaks are max-plus recurrences that restart at
band begin - 1, so they are prefix maxima.  This model pins that reformulation against the band-by-band walk."""
import random

BIG = 1 << 20
SGAIN, FGAIN = 0x4d8, 0x280


def band_by_band(psd, halfrate):
    sdecay, fdecay = 0x13 >> halfrate, 0x53 >> halfrate
    mk = [None] * 22

    def lc1(a, b0, b1):
        return 384 if b0 + 256 == b1 else (max(a - 64, 0) if b0 > b1 else a)
    lowcomp, fast, slow, begin = 0, 0, 0, 7
    lowcomp = lc1(lowcomp, psd[0], psd[1])
    mk[0] = psd[0] - FGAIN - lowcomp
    lowcomp = lc1(lowcomp, psd[1], psd[2])
    mk[1] = psd[1] - FGAIN - lowcomp
    for b in range(2, 7):
        lowcomp = lc1(lowcomp, psd[b], psd[b + 1])
        fast, slow = psd[b] - FGAIN, psd[b] - SGAIN
        mk[b] = fast - lowcomp
        if psd[b] <= psd[b + 1]:
            begin = b + 1
            break
    for b in range(begin, 22):
        b0, b1 = psd[b], psd[b + 1]
        if b < 7:
            lowcomp = lc1(lowcomp, b0, b1)
        elif b < 20:
            lowcomp = 320 if b0 + 256 == b1 else (max(lowcomp - 64, 0) if b0 > b1 else lowcomp)
        else:
            lowcomp = max(lowcomp - 128, 0)
        fast = max(fast - fdecay, psd[b] - FGAIN)
        slow = max(slow - sdecay, psd[b] - SGAIN)
        mk[b] = max(fast - lowcomp, slow)
    return mk, fast, slow


def lanes_as_bands(psd, halfrate):
    """what the 22 lanes of the kernel compute (Hillis-Steele scans written out)"""
    sdecay, fdecay = 0x13 >> halfrate, 0x53 >> halfrate
    n = 22
    a, c = [0] * n, [0] * n
    for b in range(n):
        p0, p1 = psd[b], psd[b + 1]
        if b < 20:
            if p0 + 256 == p1:
                a[b], c[b] = BIG, (384 if b < 7 else 320)
            elif p0 > p1:
                a[b], c[b] = 64, 0
            else:
                a[b], c[b] = 0, -BIG
        else:
            a[b], c[b] = 128, 0
    rise = [b for b in range(2, 7) if psd[b] <= psd[b + 1]]
    begin = rise[0] + 1 if rise else 7
    pf = [psd[b] - FGAIN + b * fdecay if b >= begin - 1 else -BIG for b in range(n)]
    ps = [psd[b] - SGAIN + b * sdecay if b >= begin - 1 else -BIG for b in range(n)]
    o = 1
    while o < 32:
        a2, c2, pf2, ps2 = a[:], c[:], pf[:], ps[:]
        for b in range(o, n):
            c2[b] = max(c[b - o] - a[b], c[b])
            a2[b] = min(a[b - o] + a[b], BIG)
            pf2[b] = max(pf[b], pf[b - o])
            ps2[b] = max(ps[b], ps[b - o])
        a, c, pf, ps = a2, c2, pf2, ps2
        o *= 2
    mk = []
    for b in range(n):
        lowcomp = max(-a[b], c[b])
        f, s = pf[b] - b * fdecay, ps[b] - b * sdecay
        mk.append(psd[b] - FGAIN - lowcomp if b < begin else max(f - lowcomp, s))
    return mk, pf[21] - 21 * fdecay, ps[21] - 21 * sdecay


def test_masking_curve_by_scans_equals_the_band_by_band_walk():
    rng = random.Random(3)
    for _ in range(20000):
        v, psd = rng.randint(0, 3072), []
        for _ in range(24):
            v = max(-300, min(3100, v + rng.choice([0, 0, 256, -256, 128, -128, -64, 37, -511, 300, -40])))
            psd.append(v)
        hr = rng.randint(0, 2)
        assert band_by_band(psd, hr) == lanes_as_bands(psd, hr)
