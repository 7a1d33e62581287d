"""Test-only AC-3 bitstream writer.

The reference encoder never emits block switching, coupling, rematrixing,
dynamic range words, delta bit allocation, skip fields, dual mono or the 2/1,
3/1 modes, so the decoder paths for those are exercised with frames produced
here: syntactically valid frames with random-but-smooth exponents and uniformly
random mantissa codes.  Validity is established by the unmodified reference
decoder accepting every frame (tests assert that), parity is then GPU vs
reference on these frames.

The writer mirrors the decoder's state machine (exponent / bit-allocation reuse)
and calls a bit allocator (the pinned oracle restatement, or the reference) to
know how many mantissa bits each bin takes.
"""
import numpy as np

NFCHANS = [2, 1, 2, 3, 3, 4, 4, 5]
BITRATES = [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512, 576, 640]
CPL_BAND = [31, 35, 37, 39, 41, 42, 43, 44, 45, 45, 46, 46, 47, 47, 48, 48]


class Injected(Exception):
    """Raised by the writer once a deliberately invalid field has been written: the rest of the frame is padding."""

    def __init__(self, bw):
        self.bw = bw


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, n, v):
        assert 0 <= v < (1 << n), (n, v)
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def __len__(self):
        return len(self.bits)

    def tobytes(self, nbytes):
        assert len(self.bits) <= nbytes * 8, "frame overflow: %d bits > %d" % (len(self.bits), nbytes * 8)
        b = np.zeros(nbytes * 8, np.uint8)
        b[: len(self.bits)] = self.bits
        return np.packbits(b)


def frame_bytes(fscod, frmsizecod):
    kbps = BITRATES[frmsizecod >> 1]
    if fscod == 0:
        return 4 * kbps
    if fscod == 1:
        return 2 * (320 * kbps // 147 + (frmsizecod & 1))
    return 6 * kbps


def _walk_exponents(rng, n, start, smooth):
    """n exponents following `start` with steps in [-2, 2] staying inside [0, 24]."""
    out = []
    e = start
    for _ in range(n):
        if rng.rand() < smooth:
            d = 0
        else:
            d = int(rng.randint(-2, 3))
        if e + d < 0 or e + d > 24:
            d = -d
        if e + d < 0 or e + d > 24:
            d = 0
        e += d
        out.append(e)
    return out


class StreamWriter:
    """Keeps the cross-block state a decoder keeps, so that `reuse` strategies are legal."""

    def __init__(self, seed, acmod, lfeon, fscod=0, frmsizecod=36, bsid=8, alloc=None, features=None):
        self.rng = np.random.RandomState(seed)
        self.acmod, self.lfeon, self.fscod, self.frmsizecod, self.bsid = acmod, lfeon, fscod, frmsizecod, bsid
        self.nfchans = NFCHANS[acmod]
        self.alloc = alloc                      # callable like Oracle.bit_allocate
        f = dict(blksw=0.3, cpl=0.7, dynrng=0.5, deltba=0.15, skip=0.2, remat=0.7, reuse=0.5, phsflg=0.5,
                 dith=0.8, addbsi=0.2, zero_snr=0.0, badframe=0.0)
        if features:
            f.update(features)
        self.f = f
        self.exp = np.zeros((7, 256), np.int64)     # 0..4 fbw, 5 lfe, 6 cpl
        self.bap = np.zeros((7, 256), np.int64)     # liba52 numbering
        self.endmant = [0] * 5
        self.cplinu = 0
        self.chincpl = [0] * 5
        self.cplbegf = self.cplendf = 0
        self.cplleak_sent = False
        self.deltba = [None] * 7
        self.deltbae = [2] * 7
        self.phsflginu = 0
        self.ncplbnd = 0
        self.cplfleak = self.cplsleak = 0
        self.cplco_sent = [False] * 5
        self.cpl_new = False
        self.csnr = 0

    # -- helpers ---------------------------------------------------------------
    def _bits_of(self, bap, counts):
        """Mantissa bits taken by one bin with liba52 bap code, updating the group phases."""
        if bap == 0:
            return 0
        if bap == -1:
            counts[0] += 1
            return 5 if counts[0] % 3 == 1 else 0
        if bap == -2:
            counts[1] += 1
            return 7 if counts[1] % 3 == 1 else 0
        if bap == -3:
            counts[2] += 1
            return 7 if counts[2] % 2 == 1 else 0
        return int(bap)

    def _emit_mantissas(self, bw, bap_seq, counts):
        rng = self.rng
        for bap in bap_seq:
            n = self._bits_of(int(bap), counts)
            if n == 0:
                continue
            if bap == -1:
                bw.put(5, int(rng.randint(27)))
            elif bap == -2:
                bw.put(7, int(rng.randint(125)))
            elif bap == -3:
                bw.put(7, int(rng.randint(121)))
            elif bap == 3:
                bw.put(3, int(rng.randint(7)))
            elif bap == 4:
                bw.put(4, int(rng.randint(15)))
            else:
                bw.put(n, int(rng.randint(1 << n)))

    def _write_exps(self, bw, arr, strategy, ngrps, start_exp, dest, blk=-1):
        rep = 1 << (strategy - 1)
        if arr == 0 and self._hit(blk, "exp_code"):
            bw.put(7, 125 + int(self.rng.randint(3)))      # not a group code: exp_1[] adds 25 (parse.c:226-229)
            raise Injected(bw)
        if arr == 0 and self._hit(blk, "exp_range"):
            for _ in range(ngrps):
                bw.put(7, 124)                             # +2 +2 +2 per group: past 24 within three groups (:227-256)
            raise Injected(bw)
        vals = _walk_exponents(self.rng, 3 * ngrps, start_exp, 0.4)
        prev = start_exp
        for g in range(ngrps):
            d = []
            for j in range(3):
                d.append(vals[3 * g + j] - prev + 2)
                prev = vals[3 * g + j]
            bw.put(7, d[0] * 25 + d[1] * 5 + d[2])
        k = dest
        for v in vals:
            for _ in range(rep):
                if k < 256:
                    self.exp[arr, k] = v
                k += 1

    # -- one frame ----------------------------------------------------------------
    def frame(self):
        self.frame_no = getattr(self, "frame_no", -1) + 1
        for attempt in range(12):
            state = self._save()
            csnr = max(1, int(self.rng.randint(8, 40)) - 4 * attempt)
            try:
                return self._frame(csnr)
            except Injected as inj:
                # the decoder stops at the injected field (a52_block returns 1): what follows is never read
                if len(inj.bw) <= 8 * frame_bytes(self.fscod, self.frmsizecod):
                    return inj.bw.tobytes(frame_bytes(self.fscod, self.frmsizecod))
                self._restore(state)
            except AssertionError:
                self._restore(state)
        raise RuntimeError("could not fit a frame")

    def _hit(self, blk, site):
        """inject = (frame, block, site): write an invalid value at that site (one of the `return 1` sites of
        liba52's a52_block, parse.c:218-294, 600-701)."""
        inj = getattr(self, "inject", None)
        return inj is not None and inj == (self.frame_no, blk, site)

    def _save(self):
        return (self.rng.get_state(), self.exp.copy(), self.bap.copy(), list(self.endmant), self.cplinu,
                list(self.chincpl), self.cplbegf, self.cplendf, self.cplleak_sent, list(self.deltba), list(self.deltbae))

    def _restore(self, s):
        (st, self.exp, self.bap, self.endmant, self.cplinu, self.chincpl, self.cplbegf, self.cplendf,
         self.cplleak_sent, self.deltba, self.deltbae) = s
        self.rng.set_state(st)
        self.rng.rand()      # perturb so that the retry differs

    def _frame(self, csnr):
        rng, f = self.rng, self.f
        acmod, nfchans, lfeon = self.acmod, self.nfchans, self.lfeon
        nbytes = frame_bytes(self.fscod, self.frmsizecod)
        halfrate = max(0, self.bsid - 8)
        bw = BitWriter()
        bw.put(16, 0x0B77)
        bw.put(16, int(rng.randint(65536)))          # crc1: never checked by liba52
        bw.put(2, self.fscod)
        bw.put(6, self.frmsizecod)
        bw.put(5, self.bsid)
        bw.put(3, 0)
        bw.put(3, acmod)
        if (acmod & 1) and acmod != 1:
            bw.put(2, int(rng.randint(4)))
        if acmod & 4:
            bw.put(2, int(rng.randint(4)))
        if acmod == 2:
            bw.put(2, int(rng.randint(4)))           # dsurmod (2 = Dolby surround)
        bw.put(1, lfeon)
        for _ in range(2 if acmod == 0 else 1):
            bw.put(5, int(rng.randint(1, 32)))
            if rng.rand() < 0.3:
                bw.put(1, 1); bw.put(8, int(rng.randint(256)))
            else:
                bw.put(1, 0)
            if rng.rand() < 0.3:
                bw.put(1, 1); bw.put(8, int(rng.randint(256)))
            else:
                bw.put(1, 0)
            if rng.rand() < 0.3:
                bw.put(1, 1); bw.put(7, int(rng.randint(128)))
            else:
                bw.put(1, 0)
        bw.put(2, int(rng.randint(4)))
        for _ in range(2):
            if rng.rand() < 0.2:
                bw.put(1, 1); bw.put(14, int(rng.randint(1 << 14)))
            else:
                bw.put(1, 0)
        if rng.rand() < f["addbsi"]:
            n = int(rng.randint(4))
            bw.put(1, 1); bw.put(6, n)
            for _ in range(n + 1):
                bw.put(8, int(rng.randint(256)))
        else:
            bw.put(1, 0)

        self.deltbae = [2] * 7                   # per-frame reset (parse.c:173-175); lfe is always NONE
        zero_snr = rng.rand() < f["zero_snr"]
        bai = None
        chbai = [0] * 7
        for blk in range(6):
            for ch in range(nfchans):
                bw.put(1, int(rng.rand() < f["blksw"]))
            for ch in range(nfchans):
                bw.put(1, int(rng.rand() < f["dith"]))
            for _ in range(2 if acmod == 0 else 1):
                if rng.rand() < f["dynrng"]:
                    bw.put(1, 1); bw.put(8, int(rng.randint(256)))
                else:
                    bw.put(1, 0)
            # coupling strategy
            cplstre = 1 if blk == 0 else int(rng.rand() < 0.25)
            if self._hit(blk, "cpl_mono"):                     # coupling in a 1+1 or 1/0 frame (parse.c:611-613)
                assert acmod < 2
                bw.put(1, 1); bw.put(1, 1)
                for ch in range(nfchans):
                    bw.put(1, 1)
                raise Injected(bw)
            if self._hit(blk, "cpl_range"):                    # cplendf + 3 - cplbegf < 0 (parse.c:620-621)
                assert acmod >= 2
                bw.put(1, 1); bw.put(1, 1)
                for ch in range(nfchans):
                    bw.put(1, 1)
                if acmod == 2:
                    bw.put(1, 0)
                bw.put(4, 12 + int(rng.randint(4))); bw.put(4, int(rng.randint(0, 9)))
                raise Injected(bw)
            bw.put(1, cplstre)
            if cplstre:
                cplinu = int(acmod >= 2 and rng.rand() < f["cpl"])
                bw.put(1, cplinu)
                self.cplinu = cplinu
                self.chincpl = [0] * 5
                if cplinu:
                    while sum(self.chincpl) < min(2, nfchans):
                        self.chincpl = [int(rng.rand() < 0.7) if c < nfchans else 0 for c in range(5)]
                    for ch in range(nfchans):
                        bw.put(1, self.chincpl[ch])
                    if acmod == 2:
                        self.phsflginu = int(rng.rand() < f["phsflg"])
                        bw.put(1, self.phsflginu)
                    self.cplbegf = int(rng.randint(0, 12))
                    self.cplendf = int(rng.randint(max(self.cplbegf - 2, 0), 16))
                    bw.put(4, self.cplbegf); bw.put(4, self.cplendf)
                    nsub = self.cplendf + 3 - self.cplbegf
                    self.ncplbnd = nsub
                    for i in range(nsub - 1):
                        b = int(rng.rand() < 0.4)
                        bw.put(1, b)
                        self.ncplbnd -= b
                    self.cplleak_sent = False
                self.cpl_new = True
                self.cplco_sent = [False] * 5
            else:
                self.cpl_new = False
            cplinu = self.cplinu
            if cplinu:
                anyco = False
                for ch in range(nfchans):
                    if self.chincpl[ch]:
                        cplcoe = 1 if not self.cplco_sent[ch] else int(rng.rand() < 0.4)
                        bw.put(1, cplcoe)
                        if cplcoe:
                            anyco = True
                            self.cplco_sent[ch] = True
                            bw.put(2, int(rng.randint(4)))
                            for _ in range(self.ncplbnd):
                                bw.put(4, int(rng.randint(16))); bw.put(4, int(rng.randint(16)))
                if acmod == 2 and self.phsflginu and anyco:
                    for _ in range(self.ncplbnd):
                        bw.put(1, int(rng.rand() < 0.5))
            if acmod == 2:
                rematstr = 1 if blk == 0 else int(rng.rand() < 0.4)
                bw.put(1, rematstr)
                if rematstr:
                    end = (37 + 12 * self.cplbegf) if cplinu else 253
                    edges = [25, 37, 61, 253]
                    i = 0
                    while True:
                        bw.put(1, int(rng.rand() < f["remat"]))
                        if not edges[i] < end:
                            break
                        i += 1
            # exponent strategies
            force_new = (blk == 0) or self.cpl_new
            cplexpstr = 0
            if cplinu:
                cplexpstr = int(rng.randint(1, 4)) if force_new or rng.rand() > f["reuse"] else 0
                bw.put(2, cplexpstr)
            chexpstr = []
            for ch in range(nfchans):
                s = int(rng.randint(1, 4)) if force_new or rng.rand() > f["reuse"] else 0
                chexpstr.append(s)
                bw.put(2, s)
            lfeexpstr = 0
            if lfeon:
                lfeexpstr = 1 if blk == 0 or rng.rand() > f["reuse"] else 0
                bw.put(1, lfeexpstr)
            cplstrt, cplend = 37 + 12 * self.cplbegf, 73 + 12 * self.cplendf
            for ch in range(nfchans):
                if chexpstr[ch]:
                    if cplinu and self.chincpl[ch]:
                        self.endmant[ch] = cplstrt
                    else:
                        if self._hit(blk, "chbwcod"):          # chbwcod > 60 (parse.c:697-698)
                            bw.put(6, 61 + int(rng.randint(3)))
                            raise Injected(bw)
                        bwc = int(rng.randint(0, 61))
                        bw.put(6, bwc)
                        self.endmant[ch] = 73 + 3 * bwc
            if cplexpstr:
                ngrps = (cplend - cplstrt) // (3 << (cplexpstr - 1))
                absexp = int(rng.randint(2, 10))
                bw.put(4, absexp)
                self._write_exps(bw, 6, cplexpstr, ngrps, absexp << 1, cplstrt)
            for ch in range(nfchans):
                if chexpstr[ch]:
                    gsz = 3 << (chexpstr[ch] - 1)
                    ngrps = (self.endmant[ch] + gsz - 4) // gsz
                    e0 = int(rng.randint(0, 12))
                    bw.put(4, e0)
                    self.exp[ch, 0] = e0
                    self._write_exps(bw, ch, chexpstr[ch], ngrps, e0, 1, blk)
                    bw.put(2, int(rng.randint(4)))
            if lfeexpstr:
                e0 = int(rng.randint(0, 12))
                bw.put(4, e0)
                self.exp[5, 0] = e0
                self._write_exps(bw, 5, 1, 2, e0, 1)
            # bit allocation parameters
            baie = 1 if blk == 0 else int(rng.rand() < 0.15)
            bw.put(1, baie)
            if baie:
                bai = int(rng.randint(2048))
                if (bai & 7) == 7 and rng.rand() < 0.8:
                    bai &= ~1                                  # floorcod 7 is extreme; keep it rare
                bw.put(11, bai)
            # the coupling channel's snr offset only travels with snroffste: resend when coupling starts
            snre = 1 if (blk == 0 or (self.cpl_new and cplinu)) else int(rng.rand() < 0.2)
            bw.put(1, snre)
            if snre:
                c = 0 if zero_snr else csnr
                bw.put(6, c)
                self.csnr = c
                order = ([6] if cplinu else []) + list(range(nfchans)) + ([5] if lfeon else [])
                for a in order:
                    v = 0 if zero_snr else int(rng.randint(128))
                    bw.put(7, v)
                    chbai[a] = v
            if cplinu:
                leake = 1 if not self.cplleak_sent else int(rng.rand() < 0.2)
                bw.put(1, leake)
                if leake:
                    self.cplfleak, self.cplsleak = int(rng.randint(8)), int(rng.randint(8))
                    bw.put(3, self.cplfleak); bw.put(3, self.cplsleak)
                    self.cplleak_sent = True
            if self._hit(blk, "deltba_len"):                   # a delta segment running past band 50 (parse.c:287-288)
                bw.put(1, 1)
                order = ([6] if cplinu else []) + list(range(nfchans))
                for a in order:
                    bw.put(2, 1)                               # new info for every array
                bw.put(3, 1)                                   # two segments
                bw.put(5, 31); bw.put(4, 3); bw.put(3, 5)      # band 31, 3 bands: fine
                bw.put(5, 10); bw.put(4, 9); bw.put(3, 2)      # band 44 + 9 >= 50
                raise Injected(bw)
            if rng.rand() < f["deltba"]:
                bw.put(1, 1)
                order = ([6] if cplinu else []) + list(range(nfchans))
                modes = {}
                for a in order:
                    choices = [1, 2] + ([0] if self.deltba[a] is not None and self.deltbae[a] != 2 else [])
                    m = int(rng.choice(choices))
                    modes[a] = m
                    bw.put(2, m)
                for a in order:
                    self.deltbae[a] = modes[a]
                    if modes[a] == 1:
                        nseg = int(rng.randint(1, 5))
                        bw.put(3, nseg - 1)
                        d = np.zeros(50, np.int8)
                        band = 0
                        for _ in range(nseg):
                            offs = int(rng.randint(0, 8))
                            ln = int(rng.randint(0, 5))
                            if band + offs + ln >= 50:
                                offs, ln = 0, 0
                            code = int(rng.randint(8))
                            bw.put(5, offs); bw.put(4, ln); bw.put(3, code)
                            band += offs
                            delta = code - 3 if code >= 4 else code - 4
                            for _ in range(ln):
                                d[band] = delta
                                band += 1
                        self.deltba[a] = d
            else:
                bw.put(1, 0)
            if rng.rand() < f["skip"]:
                n = int(rng.randint(0, 20))
                bw.put(1, 1); bw.put(9, n)
                for _ in range(n):
                    bw.put(8, int(rng.randint(256)))
            else:
                bw.put(1, 0)

            # bit allocation exactly as the decoder will redo it
            all_zero = (self.csnr == 0 and all((chbai[a] >> 3) == 0 for a in
                                               ([6] if cplinu else []) + list(range(nfchans)) + ([5] if lfeon else [])))
            if all_zero:
                self.bap[:] = 0
            else:
                for ch in range(nfchans):
                    d = self.deltba[ch] if self.deltbae[ch] in (0, 1) else None
                    self.bap[ch] = self.alloc(self.fscod, halfrate, bai, self.csnr, chbai[ch], self.exp[ch],
                                              self.endmant[ch], d)
                if cplinu:
                    d = self.deltba[6] if self.deltbae[6] in (0, 1) else None
                    self.bap[6] = self.alloc(self.fscod, halfrate, bai, self.csnr, chbai[6], self.exp[6], cplend, d,
                                             CPL_BAND[self.cplbegf], cplstrt, (9 - self.cplfleak) << 8,
                                             (9 - self.cplsleak) << 8)
                if lfeon:
                    self.bap[5] = self.alloc(self.fscod, halfrate, bai, self.csnr, chbai[5], self.exp[5], 7, None)
            counts = [0, 0, 0]
            done_cpl = False
            for ch in range(nfchans):
                self._emit_mantissas(bw, self.bap[ch, : self.endmant[ch]], counts)
                if cplinu and self.chincpl[ch] and not done_cpl:
                    done_cpl = True
                    self._emit_mantissas(bw, self.bap[6, cplstrt:cplend], counts)
            if lfeon:
                self._emit_mantissas(bw, self.bap[5, :7], counts)
        return bw.tobytes(nbytes)


def make_stream(seed, acmod, lfeon, nframes, alloc, fscod=0, frmsizecod=36, bsid=8, features=None, inject=None):
    """Concatenated frames of one synthetic stream (uint8 array) and its frame size.
    inject = (frame, block, site): that block carries one deliberately invalid field (see StreamWriter._hit)."""
    w = StreamWriter(seed, acmod, lfeon, fscod, frmsizecod, bsid, alloc, features)
    w.inject = inject
    frames = [w.frame() for _ in range(nframes)]
    return np.concatenate(frames), frame_bytes(fscod, frmsizecod)
