"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:  python oracle/make_golden.py [name ...]
The fixtures are small on purpose (they are committed); full-size parity uses the oracle
restatements, which tests/test_oracle_*.py pin against these fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refenv  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cnoise(rng, n, scale=0.5):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def gen_channelizer_c5():
    from wavecapsdr.dsp.channelizer import PolyphaseChannelizer
    from wavecapsdr.dsp.fm import quadrature_demod

    rng = np.random.default_rng(5)
    fs, bw = 125_000_000.0, 488281
    n, cut = 256 + 128 * 40 + 50, 256 + 128 * 17 + 3
    x = cnoise(rng, n)
    # an FM carrier on bin 12 so the discriminator output is not just noise
    t = np.arange(n)
    x += (0.4 * np.exp(1j * (2 * np.pi * 12 * 488281.25 / fs * t + 8.0 * np.sin(2 * np.pi * 3e3 / fs * t)))).astype(np.complex64)
    ch = PolyphaseChannelizer(fs, bw)
    f1 = np.array(ch.process(x[:cut]))
    f2 = np.array(ch.process(x[cut:]))
    rate = int(ch.channel_sample_rate)
    fm2 = np.stack([quadrature_demod(ch.extract_channel(list(f2), k), rate) for k in range(256)], axis=1)
    np.savez_compressed(os.path.join(OUT, "channelizer_c5.npz"), fs=fs, bw=bw, cut=cut, x=x, frames1=f1,
                        frames2=f2, fm2=fm2.astype(np.float32), arm_history=ch.arm_history, arms=ch.arms,
                        demod_rate=rate)


GENERATORS = {"channelizer_c5": gen_channelizer_c5}


def main(argv):
    refenv.load()
    os.makedirs(OUT, exist_ok=True)
    names = argv or list(GENERATORS)
    for name in names:
        GENERATORS[name]()
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1:])
