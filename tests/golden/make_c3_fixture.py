#!/usr/bin/env python3
"""Generate tests/golden/c3_fixture.npz: the config-3 corpus base (BASELINE.json configs[2]).

8 unique 5.1 640 kb/s streams x 32 frames written by the test-only bitstream writer
(tests/bitstream_writer.py: block switching in half the blocks, coupling in 90 %, dynrng words in 50 %, delta
bit allocation in 10 %, random mantissas) + the energy of the UNMODIFIED reference decoder's output per stream
for the stereo and the 5.1 request.  bench.py tiles the frames into its config-3 figures and checks the energies.
Run in the build container after `make -C oracle ref`:   python tests/golden/make_c3_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refbind import RefA52, Oracle, A52_STEREO, A52_3F2R, A52_LFE, A52_ADJUST_LEVEL  # noqa
from bitstream_writer import make_stream  # noqa

NS, NF, FB = 8, 32, 2560


def main():
    ref, ora = RefA52(), Oracle()
    frames, e_st, e_51 = [], [], []
    for k in range(NS):
        es, fb = make_stream(0xC3000 + k, 7, 1, NF, ora.bit_allocate, frmsizecod=36,
                             features=dict(blksw=0.5, cpl=0.9, dynrng=0.5, deltba=0.1))
        assert fb == FB
        nf, pcm = ref.decode_stream(es, A52_STEREO | A52_ADJUST_LEVEL, 1.0, 0.0)
        assert nf == NF
        e_st.append(float((pcm.astype(np.float64) ** 2).sum()))
        nf, pcm = ref.decode_stream(es, A52_3F2R | A52_LFE, 1.0, 0.0)
        assert nf == NF
        e_51.append(float((pcm.astype(np.float64) ** 2).sum()))
        frames.append(np.asarray(es, np.uint8).reshape(NF, FB))
        print(k, e_st[-1], e_51[-1])
    np.savez_compressed(os.path.join(HERE, "c3_fixture.npz"), frames=np.stack(frames),
                        energy_stereo=np.array(e_st), energy_51=np.array(e_51))


if __name__ == "__main__":
    main()
