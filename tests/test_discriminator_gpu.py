"""GPU parity (through the C ABI): the voice-channel discriminator path — wc_fm_discriminator vs
np.diff(np.unwrap(...)) of the live reference (float32 angles: <= 6e-7 abs, two float32 ulps at pi) and wc_discdemod_* vs DiscriminatorDemodulator:
dibits and counts identical to the reference goldens and to the oracle on a 40-channel bank."""
import numpy as np
import pytest

from conftest import golden_path
from oracle import discriminator as od
from oracle.c4fm import modulate_c4fm, random_frames
from oracle.make_golden import discriminator_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(golden_path("p25_discriminator.npz"))


@pytest.mark.parametrize("case", discriminator_cases(), ids=lambda c: c[0])
def test_iq_to_dibits_matches_reference_golden(native, gold, case):
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorDemodulator
    from wavecap_sdr_b200.trunking import VoiceDiscriminator

    name, chunk = case[0], case[1]
    x = gold[name + "_x"]
    disc = VoiceDiscriminator(1)
    d = DiscriminatorDemodulator(sample_rate=48000)
    aud, ds, cnt = [], [], []
    starts = list(range(0, len(x), chunk))
    for j, s0 in enumerate(starts):
        au = disc.process(x[s0:s0 + chunk].reshape(1, -1))[0]
        if name == "disc_999" and j == len(starts) // 2:
            d.reset()
        # the demodulator is fed the reference's own float32 audio so that a 1-ulp difference of the float64
        # discriminator (cumsum rounding inside np.unwrap) cannot leak into the dibit comparison
        ref_au = gold[name + "_audio"][s0:s0 + len(au)]
        a = d.demodulate(ref_au.astype(np.float32))
        aud.append(au)
        ds.append(a)
        cnt.append(len(a))
    assert np.max(np.abs(np.concatenate(aud) - gold[name + "_audio"])) <= 6e-7
    assert np.array_equal(np.array(cnt, np.int32), gold[name + "_counts"])
    assert np.array_equal(np.concatenate(ds), gold[name + "_dibits"])
    st = gold[name + "_state"]
    s = d._bank.state(0)
    assert abs(s["input_gain"] - st[0]) <= 1e-6 * abs(st[0]) and abs(s["symbol_spread"] - st[3]) <= 1e-5
    assert d.demodulate(np.zeros(0, np.float32)).size == 0


def test_gpu_audio_end_to_end(native, gold):
    """IQ -> GPU discriminator -> GPU demodulator with nothing taken from the reference in between: dibits equal to the
    reference's wherever the float32 audio is the same, i.e. everywhere except after a (rare) 1-ulp audio difference."""
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorDemodulator
    from wavecap_sdr_b200.trunking import VoiceDiscriminator

    x = gold["disc_2400_x"]
    disc, d = VoiceDiscriminator(1), DiscriminatorDemodulator(48000)
    ds, same = [], True
    for s0 in range(0, len(x), 2400):
        au = disc.process(x[s0:s0 + 2400].reshape(1, -1))[0].astype(np.float32)
        same &= bool(np.array_equal(au, gold["disc_2400_audio"][s0:s0 + len(au)].astype(np.float32)))
        ds.append(d.demodulate(au))
    got = np.concatenate(ds)
    ref = gold["disc_2400_dibits"]
    assert len(got) == len(ref)
    if same:
        assert np.array_equal(got, ref)
    else:
        assert np.mean(got != ref) < 0.01


def test_bank_of_40_channels_vs_oracle(native):
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorBank

    C, chunk = 40, 3000
    aus = []
    for c in range(C):
        rng = np.random.default_rng(500 + c)
        x = modulate_c4fm(random_frames(rng, n_frames=5, payload=150, gap=40), 48000, snr_db=16.0 + c % 12,
                          cfo_hz=25.0 * (c - 20), timing=0.025 * c, seed=500 + c)
        au, _ = od.fm_discriminator(x, 0.0)
        scale = [1.0, 0.3, 4.0, 0.02][c % 4]   # exercises the auto gain and the spread clamps
        aus.append((au * scale).astype(np.float32))
    n = min(len(a) for a in aus)
    A = np.array([a[:n] for a in aus])
    bank = DiscriminatorBank(C, 48000)
    oracles = [od.DiscriminatorOracle(48000, portable=True) for _ in range(C)]
    clamped = 0
    for s0 in range(0, n, chunk):
        dib, soft, cnt = bank.demodulate(A[:, s0:s0 + chunk])
        for c in range(C):
            ref = oracles[c].demodulate(A[c, s0:s0 + chunk].copy())
            k = int(cnt[c])
            assert k == len(ref), (c, s0)
            assert np.array_equal(dib[c, :k], ref), (c, s0)
            clamped += oracles[c].state_dtypes()["spread"] == "float"
    for c in range(C):
        s = bank.state(c)
        assert abs(s["symbol_spread"] - float(oracles[c].spread)) <= 1e-6
        assert abs(s["symbol_clock"] - float(oracles[c].clock)) <= 1e-6
    assert clamped > 0  # the literal-1.6/2.4 branch was really exercised


@pytest.mark.gpu
def test_bank_rows_do_not_depend_on_the_bank(native):
    """A channel's dibits / soft values / loop state are the same whichever lane, CTA and bank size it runs in: 67 channels
    (three CTAs, the last with three live rows) against the same rows demodulated one at a time, over calls whose lengths
    leave partial tiles (and one call shorter than the interpolator history)."""
    from wavecap_sdr_b200.decoders.p25 import DiscriminatorBank
    C = 67
    rng = np.random.default_rng(77)
    base = []
    for c in range(5):
        x = modulate_c4fm(random_frames(rng, n_frames=4, payload=150, gap=40), 48000, snr_db=20.0, cfo_hz=60.0 * c,
                          timing=0.2 * c, seed=900 + c)
        au, _ = od.fm_discriminator(x, 0.0)
        base.append(au.astype(np.float32))
    n = min(len(a) for a in base)
    A = np.stack([np.roll(base[c % 5][:n], 7 * c) * [1.0, 0.3, 4.0][c % 3] for c in range(C)]).astype(np.float32)
    bank = DiscriminatorBank(C, 48000)
    picks = [0, 31, 32, 63, 64, 66]
    singles = {c: DiscriminatorBank(1, 48000) for c in picks}
    s0 = 0
    for ln in [5, 1000, 129, 128, 4801, n]:
        seg = A[:, s0:s0 + ln]
        if seg.shape[1] == 0:
            break
        dib, soft, cnt = bank.demodulate(seg)
        for c in picks:
            d1, s1, c1 = singles[c].demodulate(seg[c:c + 1])
            k = int(cnt[c])
            assert k == int(c1[0]), (c, s0, ln)
            assert np.array_equal(dib[c, :k], d1[0, :k]) and np.array_equal(soft[c, :k], s1[0, :k]), (c, s0, ln)
        s0 += ln
    for c in picks:
        assert bank.state(c) == singles[c].state(0), c
