"""GPU parity: csrc/channelizer.cu vs the oracle restatement of wavecapsdr/dsp/channelizer.py.

Tolerance (BASELINE.json north_star): float outputs within 1e-4 relative RMS of the reference."""
import numpy as np
import pytest

from conftest import rel_rms, wrap_rel_rms, golden_path
from oracle.channelizer import ChannelizerOracle, channelize_fm

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _iq(n, seed, scale=0.5):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def _chan(native, fs, bw, t=9):
    from wavecap_sdr_b200.dsp.channelizer import PolyphaseChannelizer

    return PolyphaseChannelizer(fs, bw, t)


def test_attrs_match_reference_design(native):
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    assert ch.channel_count == 256 and ch.taps_per_channel == 9
    assert ch.channel_sample_rate == o.channel_sample_rate
    assert ch.arms.shape == (256, 9) and ch.arms.dtype == np.float64
    assert np.abs(ch.arms - o.arms).max() < 1e-12  # in-library firwin/kaiser vs scipy


@pytest.mark.parametrize("n", [255, 256, 383, 384, 1279, 2048, 40000, 123457])
def test_c5_complex_frames_single_call(native, n):
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    x = _iq(n, 5)
    got = ch.process(x)
    exp = o.process(x)
    assert isinstance(got, list) and len(got) == exp.shape[0]
    if exp.shape[0]:
        got = np.stack(got)
        assert got.dtype == np.complex64
        assert rel_rms(got, exp) < TOL
        assert rel_rms(ch.arm_history, o.arm_history) == 0.0


def test_c5_state_carries_across_calls_and_reset(native):
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    x = _iq(60000, 6)
    cuts = [0, 5120, 10240, 10240 + 300, 10240 + 300 + 777, 40001, 60000]  # ragged, incl. < 9 frames
    for a, b in zip(cuts[:-1], cuts[1:]):
        got = ch.process_array(x[a:b])
        exp = o.process_vectorized(x[a:b])
        assert got.shape == exp.shape
        if exp.size:
            assert rel_rms(got, exp) < TOL
        assert np.array_equal(ch.arm_history, o.arm_history)
    ch.reset(); o.reset()
    assert not ch.arm_history.any()
    assert rel_rms(ch.process_array(x[:4096]), o.process(x[:4096])) < TOL


def test_c5_two_short_calls_drop_the_straddling_frame(native):
    # two 5120-sample calls give 78 frames, one 10240-sample call gives 79
    ch = _chan(native, 125_000_000, 488281)
    x = _iq(10240, 7)
    assert len(ch.process(x)) == 79
    ch.reset()
    assert len(ch.process(x[:5120])) + len(ch.process(x[5120:])) == 78


def test_c5_fused_fm_discriminator(native):
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    x = _iq(256 + 128 * 700, 8)
    rate = int(ch.channel_sample_rate)
    for part in (x[:50000], x[50000:]):
        got = ch.process_fm(part, rate)
        exp = channelize_fm(o.process_vectorized(part), rate)
        assert got.dtype == np.float32 and got.shape == exp.shape
        assert not got[0].any()
        period = 2 * np.pi * float(np.float32(rate / (2.0 * np.pi * 75000.0)))
        assert wrap_rel_rms(got, exp, period) < TOL


def test_c5_batched_chunks_equal_sequential_calls(native):
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    n, b = 30000, 4
    x = _iq(n * b, 9)
    exp = np.concatenate([o.process_vectorized(x[i * n:(i + 1) * n]) for i in range(b)])
    got = ch.process_batch(x, b)
    assert rel_rms(got, exp) < TOL
    assert np.array_equal(ch.arm_history, o.arm_history)
    ch.reset(); o.reset()
    rate = int(ch.channel_sample_rate)
    expf = np.concatenate([channelize_fm(o.process_vectorized(x[i * n:(i + 1) * n]), rate) for i in range(b)])
    gotf = ch.process_batch(x, b, fm=True, demod_sample_rate=rate)
    period = 2 * np.pi * float(np.float32(rate / (2.0 * np.pi * 75000.0)))
    assert wrap_rel_rms(gotf, expf, period) < TOL


def test_generic_channel_counts(native):
    # benchmark_dsp.py:117-123 workload: 8 MS/s, 25 kHz -> 320 channels (not a power of two)
    for fs, bw, t in ((8_000_000, 25000, 9), (2_400_000, 200000, 5), (1_000_000, 12500, 9)):
        ch = _chan(native, fs, bw, t)
        o = ChannelizerOracle(fs, bw, t)
        assert ch.channel_count == o.channel_count
        x = _iq(o.channel_count * 40 + 17, 10)
        for part in (x[: x.size // 3], x[x.size // 3:]):
            got, exp = ch.process_array(part), o.process_vectorized(part)
            assert rel_rms(got, exp) < TOL
        assert np.array_equal(ch.arm_history, o.arm_history)


def test_golden_channelizer_fixture(native):
    g = np.load(golden_path("channelizer_c5.npz"))
    ch = _chan(native, float(g["fs"]), int(g["bw"]))
    x = g["x"]
    got1 = ch.process_array(x[: int(g["cut"])])
    got2 = ch.process_array(x[int(g["cut"]):])
    assert rel_rms(got1, g["frames1"]) < TOL and rel_rms(got2, g["frames2"]) < TOL


def test_device_tensor_path(native):
    import torch

    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    x = _iq(70000, 11)
    xd = torch.from_numpy(x).cuda()
    got = ch.process_array(xd)
    assert got.is_cuda and got.dtype == torch.complex64
    assert rel_rms(got.cpu().numpy(), o.process_vectorized(x)) < TOL


def test_full_size_properties(native):
    """BASELINE config 5 size (6.25 M samples): linearity + shift-consistency, no oracle needed."""
    import torch

    ch = _chan(native, 125_000_000, 488281)
    n = 6_250_000
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda") * 0.5)
    b = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda") * 0.5)
    ya = ch.process_array(a); ch.reset()
    yb = ch.process_array(b); ch.reset()
    yab = ch.process_array(2.0 * a - 0.5 * b); ch.reset()
    assert ya.shape == (48827, 256)
    err = (yab - (2.0 * ya - 0.5 * yb)).abs().pow(2).mean().sqrt() / yab.abs().pow(2).mean().sqrt()
    assert float(err) < 1e-5
    # frame b of x[128*s:] equals frame b+s of x once the 9-block history is filled
    s = 16
    ys = ch.process_array(a[128 * s:].contiguous()); ch.reset()
    err2 = (ys[9:1000] - ya[9 + s:1000 + s]).abs().max()
    assert float(err2) < 1e-4 * float(ya.abs().max())


@pytest.mark.parametrize("world", [2, 4, 8])
def test_time_slabs_equal_the_unsharded_call(native, world):
    """multi-GPU time sharding emulated in one process: every rank's slab (9-frame halo) reproduces the rows of the
    unsharded process() call bit for bit, over two consecutive calls (history carried from the true block tail)."""
    import torch

    n = 256 + 128 * 1234 + 17
    blocks = [torch.from_numpy(_iq(n, 60 + i)).cuda() for i in range(2)]
    for fm in (False, True):
        whole = _chan(native, 125_000_000, 488281)
        ranks = [_chan(native, 125_000_000, 488281) for _ in range(world)]
        for blk in blocks:
            exp = whole.process_fm(blk) if fm else whole.process_array(blk)
            got = torch.zeros_like(exp)
            for r in range(world):
                rows, f0 = ranks[r].process_slab(blk, world, r, fm=fm)
                got[f0:f0 + rows.shape[0]] = rows
            torch.cuda.synchronize()
            if fm:
                # a slab's first emitted row is computed like any other row (not the call's zero row), except at f0 = 0
                assert torch.equal(got, exp)
            else:
                assert torch.equal(got, exp)
            assert np.array_equal(ranks[-1].arm_history, whole.arm_history)


@pytest.mark.parametrize("frames", [1, 2, 7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 32, 33, 255, 256, 257, 263, 700])
def test_c5_fused_fm_ragged_frame_counts(native, frames):
    """run / sub-tile boundaries of the pipelined kernel: prologue only, one fast sub-tile, ragged last sub-tile, several
    runs (first run emits R frames, later ones R-1 after a warm-up frame)."""
    ch = _chan(native, 125_000_000, 488281)
    o = ChannelizerOracle(125_000_000, 488281)
    n = 256 + 128 * (frames - 1) + 57          # 57 trailing samples that do not make a frame
    x = _iq(n, 100 + frames)
    rate = int(ch.channel_sample_rate)
    period = 2 * np.pi * float(np.float32(rate / (2.0 * np.pi * 75000.0)))
    for rep in range(2):                        # second call: carried history instead of zeros
        got = ch.process_fm(x, rate)
        exp = channelize_fm(o.process_vectorized(x), rate)
        assert got.shape == exp.shape == (frames, 256)
        assert wrap_rel_rms(got, exp, period) < TOL if frames > 1 else not got.any()
    y = ch.process_array(x)
    assert rel_rms(y, o.process_vectorized(x)) < TOL


def test_full_size_fused_fm_equals_discriminator_of_complex_frames(native):
    """BASELINE config 5 size, 4 chunks in one launch (R = 256 runs): the fused discriminator output equals
    angle(y_b * conj(y_{b-1})) * scale of the complex-mode output of the same kernel family, per chunk, d_0 = 0."""
    import torch

    ch = _chan(native, 125_000_000, 488281)
    n, b = 6_250_000, 4
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.view_as_complex(torch.randn((n * b, 2), generator=g, device="cuda") * 0.5)
    rate = int(ch.channel_sample_rate)
    y = ch.process_batch(x, b)
    ch.reset()
    d = ch.process_batch(x, b, fm=True, demod_sample_rate=rate)
    F = 48827
    assert tuple(y.shape) == (b * F, 256) and tuple(d.shape) == (b * F, 256)
    scale = float(np.float32(rate / (2.0 * np.pi * 75000.0)))
    y = y.reshape(b, F, 256)
    ref = torch.zeros((b, F, 256), dtype=torch.float32, device="cuda")
    ref[:, 1:] = torch.angle(y[:, 1:] * torch.conj(y[:, :-1])) * scale
    diff = d.reshape(b, F, 256) - ref
    period = 2 * np.pi * scale
    diff = diff - period * torch.round(diff / period)
    err = float(diff.double().pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt())
    assert err < 1e-4
    assert not bool(d.reshape(b, F, 256)[:, 0].any())


def test_channelize_samples_and_calculator(native):
    """channelize_samples (dsp/channelizer.py:234-268) for a 40-channel grid (generic path) and the C5 grid (fast path)
    against outputs of the reference itself and the oracle; the tone lands in the bin ChannelCalculator names."""
    from oracle.channelizer import ChannelCalculatorOracle
    from oracle.channelizer import channelize_samples as oracle_channelize
    from oracle.make_golden import channelize_samples_input
    from wavecap_sdr_b200.dsp.channelizer import ChannelCalculator, PolyphaseChannelizer, channelize_samples

    g = np.load(golden_path("channel_calc.npz"))
    x, fs, bw, center = channelize_samples_input()
    for j, target in enumerate((center - 75000.0, center + 200000.0, center)):
        y, rate = channelize_samples(x, fs, target, center, bw)
        assert rate == float(g[f"rate{j}"]) and y.dtype == np.complex64 and y.shape == g[f"chan{j}"].shape
        assert rel_rms(y, g[f"chan{j}"]) < TOL
    # the -75 kHz tone: strongest in the bin the calculator returns
    frames = PolyphaseChannelizer(fs, bw).process_array(x)
    k = ChannelCalculator(center, fs, bw).get_channel_index(center - 75000.0)
    assert k == 37 and int(np.argmax(np.mean(np.abs(frames[8:]) ** 2, axis=0))) == k
    # C5 grid, negative offset wrapping to the top bins
    rng = np.random.default_rng(62)
    n = 256 + 128 * 700 + 5
    xc = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1).astype(np.complex64)
    tgt = 100.0e6 - 9 * 488281.0
    y, rate = channelize_samples(xc, 125.0e6, tgt, 100.0e6, 488281)
    ye, re_ = oracle_channelize(xc, 125.0e6, tgt, 100.0e6, 488281)
    assert ChannelCalculatorOracle(100.0e6, 125.0e6, 488281).get_channel_index(tgt) == 247
    assert rate == re_ and y.shape == ye.shape and rel_rms(y, ye) < TOL


# ---- int16 capture format and the audio (/20) output mode: SURVEY §8d row "C5 + audio", cli.py:449-453 ------------------

def _cs16(n, seed, amp=6000):
    rng = np.random.default_rng(seed)
    return rng.integers(-amp, amp, size=(n, 2), dtype=np.int16)


@pytest.mark.parametrize("n", [256 + 128 * 77, 256 + 128 * 7 + 5, 256 + 128 * 600 + 127, 4000])
def test_cs16_input_equals_the_converted_cf32_call(native, n):
    """int16 I,Q scaled by 1/32768 is an exact float32 value, so the int16 path must give the very same bits as the
    complex64 path fed with the reference's own conversion (cli.py:449-453), and match the oracle to 1e-4."""
    from oracle import analog as oa

    q = _cs16(n, seed=n)
    xc = oa.cs16_to_cf32(q)
    a, b, o = _chan(native, 125_000_000, 488281), _chan(native, 125_000_000, 488281), ChannelizerOracle(125_000_000, 488281)
    cut = (n // 3) & ~3                                  # two calls: carried history comes from int16 rows too
    for part_q, part_c in ((q[:cut], xc[:cut]), (q[cut:], xc[cut:])):
        fa, fb = a.process_array(part_q), b.process_array(part_c)
        assert fa.dtype == np.complex64 and np.array_equal(fa, fb)
        exp = o.process_vectorized(part_c)
        if len(exp):
            assert rel_rms(fa, exp) < TOL
    assert np.array_equal(a.arm_history, b.arm_history)
    a.reset(), b.reset()
    assert np.array_equal(a.process_fm(q), b.process_fm(xc))


def test_cs16_batched_chunks_and_device_tensors(native):
    import torch
    from oracle import analog as oa

    n, B = 256 + 128 * 300, 3
    q = _cs16(n * B, seed=5)
    a, b = _chan(native, 125_000_000, 488281), _chan(native, 125_000_000, 488281)
    batched = a.process_batch(torch.from_numpy(q).cuda(), B, fm=True).cpu().numpy()
    seq = np.concatenate([b.process_fm(oa.cs16_to_cf32(q[i * n:(i + 1) * n])) for i in range(B)])
    assert np.array_equal(batched, seq)


def _audio_oracle(frames, demod_rate, audio_rate):
    from oracle import analog as oa

    return np.stack([oa.nbfm_demod(np.ascontiguousarray(frames[:, k]), demod_rate, audio_rate) for k in range(frames.shape[1])], axis=1)


def _knife_edge_channels(frames, tol=1e-5):
    """channels holding a discriminator sample within `tol` rad of +-pi: there the sign of the wrapped angle depends on the
    last bit of x[n] conj(x[n-1]) and a flip moves that sample by 2 pi (same reason the FM tests compare modulo 2 pi)."""
    ang = np.angle(frames[1:] * np.conj(frames[:-1]))
    return np.nonzero((np.pi - np.abs(ang) < tol).any(axis=0))[0]


def test_audio_mode_matches_nbfm_demod_of_every_channel(native):
    """wc_chan_process_ex(mode AUDIO) == nbfm_demod(extract_channel(process(x), k), 976560, 48828) for all 256 channels."""
    from conftest import parity_note

    fs, bw = 125_000_000, 488281
    n = 256 + 128 * 2999 + 17                                  # 3000 frames -> 150 audio samples per channel
    rng = np.random.default_rng(71)
    t = np.arange(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.02
    for k in range(0, 256, 5):                                 # FM carriers on every fifth bin centre
        dev = 5e3 + 270.0 * k
        x += 0.2 * np.exp(1j * (2 * np.pi * k * 488281.25 / fs * t + (dev / 2e3) * np.sin(2 * np.pi * (2e3 + 10 * k) / fs * t)))
    x = x.astype(np.complex64)
    ch, o = _chan(native, fs, bw), ChannelizerOracle(fs, bw)
    got = ch.process_audio(x)
    frames = o.process_vectorized(x)
    exp = _audio_oracle(frames, 976560, 48828)
    assert got.shape == exp.shape == (150, 256) and got.dtype == np.float32
    skip = set(_knife_edge_channels(frames).tolist())
    worst = 0.0
    for k in range(256):
        if k in skip:
            continue
        worst = max(worst, rel_rms(got[:, k], exp[:, k]))
    assert len(skip) <= 6 and worst < TOL, (len(skip), worst)
    parity_note(f"channelizer audio mode (/20): {256 - len(skip)} channels vs nbfm_demod, worst rel-RMS {worst:.1e}; "
                f"{len(skip)} channels skipped (a discriminator sample within 1e-5 rad of +-pi)")


def test_audio_mode_cs16_batched_host_and_device(native):
    """int16 in, audio out, several chunks per call: numpy (host C-ABI call with copies inside) == CUDA tensors == one
    chunk at a time; each chunk is an independent nbfm_demod call (RMS and zero-extended resampler per chunk)."""
    import torch
    from oracle import analog as oa

    n, B = 256 + 128 * 1203 + 60, 3                            # 1204 frames -> 61 audio samples (ragged last block)
    q = _cs16(n * B, seed=9, amp=3000)
    t = np.arange(n * B)
    q[:, 0] += (8000 * np.cos(2 * np.pi * (20 * 488281.25 / 125e6) * t + 3 * np.sin(2 * np.pi * 2e-5 * t))).astype(np.int16)
    q[:, 1] += (8000 * np.sin(2 * np.pi * (20 * 488281.25 / 125e6) * t + 3 * np.sin(2 * np.pi * 2e-5 * t))).astype(np.int16)
    a, b, c = (_chan(native, 125_000_000, 488281) for _ in range(3))
    host = a.process_audio(q, n_chunks=B)
    dev = b.process_audio(torch.from_numpy(q).cuda(), n_chunks=B).cpu().numpy()
    seq = np.concatenate([c.process_audio(q[i * n:(i + 1) * n]) for i in range(B)])
    assert host.shape == (61 * B, 256) and np.array_equal(host, dev) and np.array_equal(host, seq)
    o = ChannelizerOracle(125_000_000, 488281)
    frames = o.process_vectorized(oa.cs16_to_cf32(q[:n]))
    exp = _audio_oracle(frames[:, 18:23], 976560, 48828)
    assert rel_rms(host[:61, 20], exp[:, 2]) < TOL


@pytest.mark.parametrize("frames", [1, 6, 19, 20, 21, 41, 97])
def test_audio_mode_short_calls(native, frames):
    """audio mode on calls shorter than the 401-tap resampler / a decimation period: zero extension at both ends, one
    output for fewer than 20 frames, ceil(F / 20) in general — against nbfm_demod of the oracle's frames."""
    fs, bw = 125_000_000, 488281
    n = 256 + 128 * (frames - 1) + 13
    rng = np.random.default_rng(900 + frames)
    t = np.arange(n)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.01
         + 0.3 * np.exp(1j * (2 * np.pi * 7 * 488281.25 / fs * t + 4.0 * np.sin(2 * np.pi * 3e-5 * t)))).astype(np.complex64)
    got = _chan(native, fs, bw).process_audio(x)
    fr = ChannelizerOracle(fs, bw).process_vectorized(x)
    assert fr.shape[0] == frames and got.shape == (-(-frames // 20), 256)
    exp = _audio_oracle(fr[:, 5:10], 976560, 48828)
    skip = set((_knife_edge_channels(fr[:, 5:10]) + 5).tolist())
    for k in range(5, 10):
        if k not in skip:
            # a one-frame call has a single discriminator sample (0 by definition): rms below min_rms, audio 0 on both sides
            if np.max(np.abs(exp[:, k - 5])) == 0:
                assert np.max(np.abs(got[:, k])) == 0
            else:
                assert rel_rms(got[:, k], exp[:, k - 5]) < TOL, (frames, k)
