"""Oracle: polyphase channelizer (wavecapsdr/dsp/channelizer.py). Test infrastructure only."""
from __future__ import annotations

import numpy as np
from scipy import signal


def design_arms(sample_rate: float, channel_bandwidth: int, taps_per_channel: int = 9):
    """Channel count and polyphase arms, channelizer.py:53-55 and :69-89.

    M = int(fs/bw) made even; prototype = firwin(M*T-1, 0.9*bw/(fs/2), kaiser 8.0) in float64;
    arms[k, j] = proto[k + j*M] (missing last tap = 0)."""
    m = int(sample_rate / channel_bandwidth)
    m -= m % 2
    proto = signal.firwin(m * taps_per_channel - 1, (channel_bandwidth * 0.9) / (sample_rate / 2),
                          window=("kaiser", 8.0)).astype(np.float64)
    padded = np.zeros(m * taps_per_channel, dtype=np.float64)
    padded[: proto.size] = proto
    arms = padded.reshape(taps_per_channel, m).T.copy()
    return m, arms


class ChannelizerOracle:
    """Frame-by-frame restatement of PolyphaseChannelizer.process (channelizer.py:91-137):
    history columns shift right by one, column 0 takes the new M-sample block, the arm outputs are
    the row-wise dot product with the (float64) arms cast to complex64, then np.fft.fft -> complex64.
    Hop is M/2; a block never straddles two calls (trailing samples are dropped)."""

    def __init__(self, sample_rate: float, channel_bandwidth: int = 25000, taps_per_channel: int = 9):
        self.channel_count, self.arms = design_arms(sample_rate, channel_bandwidth, taps_per_channel)
        self.taps_per_channel = taps_per_channel
        self.channel_sample_rate = (sample_rate / self.channel_count) * 2  # channelizer.py:58
        self.reset()

    def reset(self) -> None:  # channelizer.py:139-142
        self.arm_history = np.zeros((self.channel_count, self.taps_per_channel), dtype=np.complex64)

    def process(self, samples: np.ndarray) -> np.ndarray:
        m = self.channel_count
        frames = []
        start = 0
        while start + m <= len(samples):
            hist = np.empty_like(self.arm_history)
            hist[:, 1:] = self.arm_history[:, :-1]
            hist[:, 0] = samples[start:start + m]
            self.arm_history = hist
            u = (hist * self.arms).sum(axis=1).astype(np.complex64)
            frames.append(np.fft.fft(u).astype(np.complex64))
            start += m // 2
        if not frames:
            return np.zeros((0, m), dtype=np.complex64)
        return np.stack(frames)

    def process_vectorized(self, samples: np.ndarray) -> np.ndarray:
        """Closed form: u_b[k] = sum_j h[k+jM] * blk_{b-j}[k]; all frames of a
        call at once. Same state semantics; used as the fast CPU port."""
        m, t = self.channel_count, self.taps_per_channel
        x = np.asarray(samples, dtype=np.complex64)
        if x.size < m:
            return np.zeros((0, m), dtype=np.complex64)
        nfr = (x.size - m) // (m // 2) + 1
        blocks = np.lib.stride_tricks.sliding_window_view(x, m)[:: m // 2][:nfr]
        # history blocks: arm_history[:, j] is the block fed j frames ago -> prepend oldest first
        prior = self.arm_history.T[::-1][-(t - 1):] if t > 1 else np.zeros((0, m), np.complex64)
        allb = np.concatenate([prior, blocks], axis=0)
        u = np.zeros((nfr, m), dtype=np.complex128)
        for j in range(t):
            u += allb[t - 1 - j: t - 1 - j + nfr] * self.arms[:, j][None, :]
        y = np.fft.fft(u.astype(np.complex64), axis=1).astype(np.complex64)
        tail = allb[-t:][::-1]
        self.arm_history = np.ascontiguousarray(tail.T).astype(np.complex64)
        return y


def quadrature_demod(iq: np.ndarray, sample_rate: int) -> np.ndarray:
    """dsp/fm.py:65-97: out[0]=0, out[n]=angle(x[n]*conj(x[n-1])) * float32(fs/(2*pi*75000))."""
    x = np.asarray(iq).astype(np.complex64, copy=False)
    out = np.zeros(x.size, dtype=np.float32)
    if x.size > 1:
        out[1:] = np.angle(x[1:] * np.conj(x[:-1])) * np.float32(sample_rate / (2.0 * np.pi * 75000.0))
    return out


def channelize_fm(frames: np.ndarray, demod_sample_rate: int) -> np.ndarray:
    """quadrature_demod applied to every extracted channel of one process() call -> [F, M] f32."""
    out = np.zeros(frames.shape, dtype=np.float32)
    for k in range(frames.shape[1]):
        out[:, k] = quadrature_demod(np.ascontiguousarray(frames[:, k]), demod_sample_rate)
    return out


class ChannelCalculatorOracle:
    """Frequency <-> FFT-bin bookkeeping, channelizer.py:161-231 (bin order = FFT order: 0 = DC, negative offsets wrap
    to the end; a negative offset beyond -channel_count is NOT wrapped, like the reference)."""

    def __init__(self, center_frequency: float, sample_rate: float, channel_bandwidth: int = 25000):
        self.center_frequency = center_frequency
        self.channel_bandwidth = channel_bandwidth
        self.channel_count = int(sample_rate / channel_bandwidth)          # :181-184
        if self.channel_count % 2 != 0:
            self.channel_count -= 1

    def get_channel_index(self, target_frequency: float) -> int:          # :186-214
        steps = int(round((target_frequency - self.center_frequency) / self.channel_bandwidth))
        if steps < 0:
            return self.channel_count + steps
        return steps % self.channel_count

    def get_channel_center_frequency(self, channel_index: int) -> float:  # :216-231
        if channel_index < self.channel_count // 2:
            return self.center_frequency + channel_index * self.channel_bandwidth
        return self.center_frequency + (channel_index - self.channel_count) * self.channel_bandwidth


def channelize_samples(samples, sample_rate, target_frequency, center_frequency, channel_bandwidth=25000):
    """channelizer.py:234-268: fresh channelizer, one process() call, the target bin of every frame."""
    ch = ChannelizerOracle(sample_rate, channel_bandwidth)
    idx = ChannelCalculatorOracle(center_frequency, sample_rate, channel_bandwidth).get_channel_index(target_frequency)
    frames = ch.process_vectorized(np.asarray(samples))
    return np.ascontiguousarray(frames[:, idx]).astype(np.complex64), ch.channel_sample_rate
