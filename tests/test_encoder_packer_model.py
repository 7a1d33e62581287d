"""CPU test: lane-level model of the encoder kernel's block-warp mantissa packer (csrc/ac3_encode.cu, stage E4;
reference src/ac3enc/ac3enc.cpp:1346-1501) checked bit for bit against the encoder oracle's frames.  A warp owns one
audio block and walks its channels in coded order; a lane owns eight consecutive bins; class counters, the multiply
test for the member that opens a group and the rings are written exactly as the kernel writes them, so the algorithm is
pinned here without a GPU (the kernel itself is compared byte for byte in tests/test_encoder_gpu.py)."""
import numpy as np
import pytest

from refbind import OracleEnc
from synth import synth_pcm

PLAIN = [0, 0, 0, 3, 0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16]
WIDTH = [0, 5, 7, 3, 7, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16]
LEVELS = [0, 3, 5, 7, 11, 15, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
QBITS = [0, 0, 0, 0, 0, 0, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16]
CLS = [0, 1, 2, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
TABA = [PLAIN[b] | ((b == 1) << 8) | ((b == 2) << 16) | ((b == 4) << 24) for b in range(16)]


def i32(x):
    x &= 0xffffffff
    return x - (1 << 32) if x & 0x80000000 else x


def quant(b, c, e):
    lv = LEVELS[b]
    if lv:
        a = -c if c < 0 else c
        m = i32(lv * i32(a << e)) >> 24
        m = (m + 1) >> 1
        return (lv >> 1) + (m if c >= 0 else -m)
    q = QBITS[b]
    ls = e + q - 24
    v = i32(c << ls) if ls >= 0 else c >> (-ls)
    v = (v + 1) >> 1
    m = 1 << (q - 1)
    if v >= m:
        v = m - 1
    return v & ((1 << q) - 1)


def G3(x):
    return ((x + 2) * 43691) >> 17


def G2(x):
    return (x + 1) >> 1


class Frame:
    def __init__(self, nbits):
        self.bits = np.zeros(nbits + 64, np.uint8)

    def put(self, pos, n, v):
        assert 0 <= v < (1 << n), (n, v)
        for k in range(n):
            bit = (v >> (n - 1 - k)) & 1
            assert not (bit and self.bits[pos + k]), "overlap"
            self.bits[pos + k] |= bit


def pack_block(fr, blk, nch_all, lfe, coef, enc, bap, exp_shift, mant_pos):
    """the warp of block blk"""
    N1 = N2 = N4 = 0
    pos0 = mant_pos
    ring_v = np.zeros((3, 256), np.int32)
    ring_p = np.zeros((3, 128), np.int32)
    for ch in range(nch_all):
        ncoef = 7 if (lfe and ch == 5) else 223
        gexp = int(exp_shift[blk][ch])
        # phase A
        acc = [0] * 32
        bb = [[0] * 8 for _ in range(32)]
        for lane in range(32):
            for k in range(8):
                i = 8 * lane + k
                b = int(bap[blk][ch][i]) if i < ncoef else 0
                bb[lane][k] = b
                acc[lane] += TABA[b]
        cnt = [a >> 8 for a in acc]
        pl = [a & 0xff for a in acc]
        icnt = np.cumsum(cnt).tolist()
        ipl = np.cumsum(pl).tolist()
        tcnt, tpl = icnt[31], ipl[31]
        assert all(((tcnt >> s) & 0xff) <= 223 for s in (0, 8, 16))
        writes_v, writes_p = [], []
        for lane in range(32):
            ec, ep = icnt[lane] - cnt[lane], ipl[lane] - pl[lane]
            X1, X2, X4 = N1 + (ec & 0xff), N2 + ((ec >> 8) & 0xff), N4 + (ec >> 16)
            Xm = (N1 % 768 + (ec & 0xff)) | (N2 % 768 + ((ec >> 8) & 0xff)) << 10 | (N4 % 768 + (ec >> 16)) << 20
            pos = pos0 + ep + 5 * (G3(X1) - G3(N1)) + 7 * (G3(X2) - G3(N2)) + 7 * (G2(X4) - G2(N4))
            for k in range(8):
                b = bb[lane][k]
                if b == 0:
                    continue
                i = 8 * lane + k
                c = int(coef[blk][ch][i])
                e = int(enc[blk][ch][i]) - gexp
                v = quant(b, c, e)
                cl = CLS[b]
                if cl:
                    # packed per-class counters relative to the channel start; the carrier test is a multiply
                    # (x % 3 == 0 <=> x * 0xAAAAAAAB mod 2^32 <= 0x55555555; x even <=> x * 2^31 mod 2^32 == 0)
                    # (absolute counters modulo 768 = 3 * 256, ten bits each: all uses of x are modulo 256, 3 or 2)
                    sh = (10 * cl + 22) & 31
                    x = (Xm >> sh) & 0x3ff
                    Xm += 1 << sh
                    M, Tt = (0x80000000, 0) if cl == 3 else (0xAAAAAAAB, 0x55555555)
                    carrier = ((x * M) & 0xffffffff) <= Tt
                    assert carrier == ((x & 1) == 0 if cl == 3 else x % 3 == 0)
                    assert x < 1024
                    writes_v.append((cl - 1, x & 255, v))
                    if carrier:
                        writes_p.append((cl - 1, (x >> 1) & 127, pos))
                        pos += WIDTH[b]
                else:
                    fr.put(pos, WIDTH[b], v)
                    pos += WIDTH[b]
        for c_, x_, v_ in writes_v:
            ring_v[c_, x_] = v_
        for c_, g_, p_ in writes_p:
            ring_p[c_, g_] = p_
        # epilogue: the groups this channel closed
        t1, t2, t4 = tcnt & 0xff, (tcnt >> 8) & 0xff, tcnt >> 16
        d1a, d1b = (N1 * 43691) >> 17, ((N1 + t1) * 43691) >> 17
        d2a, d2b = (N2 * 43691) >> 17, ((N2 + t2) * 43691) >> 17
        d4a, d4b = N4 >> 1, (N4 + t4) >> 1
        n1g, n2g, n4g = d1b - d1a, d2b - d2a, d4b - d4a
        for i in range(n1g + n2g + n4g):
            if i < n1g:
                g = d1a + i; x0 = 3 * g
                code = 9 * ring_v[0, x0 & 255] + 3 * ring_v[0, (x0 + 1) & 255] + ring_v[0, (x0 + 2) & 255]
                fr.put(int(ring_p[0, (x0 >> 1) & 127]), 5, int(code))
            elif i < n1g + n2g:
                g = d2a + i - n1g; x0 = 3 * g
                code = 25 * ring_v[1, x0 & 255] + 5 * ring_v[1, (x0 + 1) & 255] + ring_v[1, (x0 + 2) & 255]
                fr.put(int(ring_p[1, (x0 >> 1) & 127]), 7, int(code))
            else:
                g = d4a + i - n1g - n2g; x0 = 2 * g
                code = 11 * ring_v[2, x0 & 255] + ring_v[2, (x0 + 1) & 255]
                fr.put(int(ring_p[2, (x0 >> 1) & 127]), 7, int(code))
        pos0 += tpl + 5 * (G3(N1 + t1) - G3(N1)) + 7 * (G3(N2 + t2) - G3(N2)) + 7 * (G2(N4 + t4) - G2(N4))
        N1 += t1; N2 += t2; N4 += t4
    # the block's open groups
    r = N1 - 3 * ((N1 * 43691) >> 17)
    if r:
        g = (N1 * 43691) >> 17; x0 = 3 * g
        code = 9 * ring_v[0, x0 & 255] + (3 * ring_v[0, (x0 + 1) & 255] if r == 2 else 0)
        fr.put(int(ring_p[0, (x0 >> 1) & 127]), 5, int(code))
    r = N2 - 3 * ((N2 * 43691) >> 17)
    if r:
        g = (N2 * 43691) >> 17; x0 = 3 * g
        code = 25 * ring_v[1, x0 & 255] + (5 * ring_v[1, (x0 + 1) & 255] if r == 2 else 0)
        fr.put(int(ring_p[1, (x0 >> 1) & 127]), 7, int(code))
    if N4 & 1:
        g = N4 >> 1
        fr.put(int(ring_p[2, g & 127]), 7, int(11 * ring_v[2, (2 * g) & 255]))
    return pos0


def side_lengths(nch_all, lfe, acmod, strategy, bap):
    """bit position of every block's first mantissa and the bits of its mantissas (the kernel's side-information sizes)"""
    nch = nch_all - (1 if lfe else 0)
    bsi = 16 + 16 + 2 + 6 + 5 + 3 + 3 + (2 if ((acmod & 1) and acmod != 1) else 0) + (2 if acmod & 4 else 0) \
        + (2 if acmod == 2 else 0) + 1 + 5 + 4 + 1 + 3
    pos = bsi
    out = []
    for blk in range(6):
        nnew = sum(1 for ch in range(nch) if strategy[blk][ch])
        ln = 0
        mant = 0
        n1 = n2 = n4 = 0
        for ch in range(nch_all):
            st = int(strategy[blk][ch])
            is_lfe = lfe and ch == 5
            if st:
                gs = {1: 1, 2: 2, 3: 4}[st]
                ng = ((7 if is_lfe else 223) + gs * 3 - 4) // (3 * gs)
                ln += 4 + 7 * ng + (0 if is_lfe else 2)
            nc = 7 if is_lfe else 223
            bp = bap[blk][ch][:nc]
            n1 += int((bp == 1).sum()); n2 += int((bp == 2).sum()); n4 += int((bp == 4).sum())
            mant += sum(PLAIN[int(b)] for b in bp)
        mant += 5 * ((n1 + 2) // 3) + 7 * ((n2 + 2) // 3) + 7 * ((n4 + 1) // 2)
        ln += 2 * nch + 1 + (2 if blk == 0 else 1) + ((5 if blk == 0 else 1) if acmod == 2 else 0) + 2 * nch \
            + (1 if lfe else 0) + 6 * nnew + 1 + (11 if blk == 0 else 0) + 1 + ((6 + 7 * nch_all) if blk == 0 else 0) + 2
        out.append((pos + ln, mant))
        pos += ln + mant
    return out


CONFIGS = [(6, 448000, 48000), (6, 384000, 44100), (5, 320000, 48000), (4, 192000, 44100), (3, 128000, 48000),
           (2, 192000, 48000), (2, 96000, 48000), (1, 64000, 32000), (2, 128000, 22050), (6, 640000, 48000),
           (6, 32000, 48000)]


@pytest.mark.parametrize("nch,br,rate", CONFIGS)
def test_packer_model_reproduces_the_oracle_mantissa_bits(nch, br, rate):
    ora = OracleEnc()
    acmod_of = {1: 1, 2: 2, 3: 3, 4: 6, 5: 7, 6: 7}
    lfe = nch == 6
    total = 0
    for s in range(3):
        nfr = 3
        pcm = synth_pcm(7, 10 * nch + s, nch, 1536 * nfr, rate, noise=[0.02, 0.2, 0.001][s], bursts=(s == 1))
        fb = ora.init(rate, br, nch)
        for f in range(nfr):
            want = ora.frame(pcm[f * 1536:(f + 1) * 1536])
            wbits = np.unpackbits(want)
            coef, strategy, enc, bap, shift = ora.get(0), ora.get(2), ora.get(3), ora.get(4), ora.get(5)
            if ora.get(6)[3]:
                bap = np.zeros_like(bap)                              # a failed search packs no mantissas
            fr = Frame(fb * 8)
            for blk, (mp, mant) in enumerate(side_lengths(nch, lfe, acmod_of[nch], strategy, bap)):
                end = pack_block(fr, blk, nch, lfe, coef, enc, bap, shift, mp)
                assert end == mp + mant, (blk, end, mp, mant)
                lim = min(mp + mant, fb * 8 - 16)
                assert (fr.bits[mp:lim] == wbits[mp:lim]).all(), (s, f, blk)
                total += mant
    assert total > 0 or br == 32000
