"""GPU: the reference's P25 control-channel monitor, `trunking.control_channel.ControlChannelMonitor.process_iq`
(control_channel.py:197-263: demodulate -> sync search -> NID BCH -> TSBK de-interleave / trellis / CRC -> TSBK parser), run
unmodified on a synthetic control channel — first untouched, then after install(), which now also reaches the by-name aliases
that module holds (`DSPC4FMDemodulator`, `P25CQPSKDemodulator`; `bch_decode`, `trellis_decode` in decoders/p25_frames.py).
The decoded TSBK results, the monitor's counters and its per-block diagnostics (BCH error counts, trellis metrics, DUID
histogram) must be IDENTICAL: "bit-exact dibits and decoded P25 frames" through the reference's own consumer.

The signal carries valid TSBKs: 80 random-but-plausible bits + CRC-16, 1/2-rate trellis (the reference's own
`trellis_encode`), block interleave (the inverse of its `deinterleave_data`), three blocks per TSDU behind sync + BCH-coded NID
with status symbols, C4FM-modulated at 48 kS/s with noise, carrier offset and a fractional timing offset."""
import numpy as np
import pytest

from conftest import parity_note
from oracle import build_ref
from oracle import bch as obch
from oracle.c4fm import modulate_c4fm
from oracle.cqpsk import modulate_cqpsk
from oracle.p25_framer import SYNC_DIBITS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not build_ref.staged(), reason="oracle/_ref not staged")]


def _crc16(bits80):
    c = 0
    for b in list(bits80) + [0] * 16:
        msb = (c >> 15) & 1
        c = ((c << 1) | int(b)) & 0xFFFF
        if msb:
            c ^= 0x1021
    return c ^ 0xFFFF


def _tsbk_block(rng, last: bool, opcode: int, trellis_encode, deinterleave_table) -> np.ndarray:
    """196 interleaved bits of one TSBK: LB | P | opcode(6) | MFID(8) | 64 data bits | CRC-16"""
    head = [int(last), 0] + [(opcode >> (5 - i)) & 1 for i in range(6)] + [0] * 8
    data = head + [int(v) for v in rng.integers(0, 2, 64)]
    crc = _crc16(data)
    bits96 = np.array(data + [(crc >> (15 - i)) & 1 for i in range(16)], dtype=np.uint8)
    dibits = np.concatenate([(bits96[0::2] << 1) | bits96[1::2], [0]]).astype(np.uint8)   # 48 + flush
    enc = np.asarray(trellis_encode(dibits), dtype=np.uint8)                              # 98 dibits
    bits196 = np.stack([(enc >> 1) & 1, enc & 1], axis=1).reshape(-1)
    # the receiver places interleaved bit i at position table[i] (decoders/p25_frames.py:534-563): the transmitter reads it there
    return bits196[np.asarray(deinterleave_table)]


def _tsdu_dibits(rng, nac, blocks):
    nid = list(obch.bch_encode((nac << 4) | 0x7))
    nid.append(sum(nid) & 1)
    payload = np.concatenate(blocks)
    body = list(SYNC_DIBITS) + [(nid[2 * i] << 1) | nid[2 * i + 1] for i in range(32)] + \
        [int((payload[2 * i] << 1) | payload[2 * i + 1]) for i in range(len(payload) // 2)]
    out = []
    for i, d in enumerate(body):
        out.append(int(d))
        if (i + 1) % 35 == 0:
            out.append(0)   # status symbol
    return out


def _signal(trellis_encode, table, seed, n_frames=10, lsm=False):
    rng = np.random.default_rng(seed)
    opcodes = [0x3A, 0x3B, 0x3C, 0x00, 0x02, 0x28, 0x2C, 0x3D, 0x39, 0x34]
    dibits = [int(v) for v in rng.integers(0, 4, 120)]
    for f in range(n_frames):
        blocks = [_tsbk_block(rng, last=(b == 2), opcode=opcodes[(3 * f + b) % len(opcodes)], trellis_encode=trellis_encode,
                              deinterleave_table=table) for b in range(3)]
        dibits += _tsdu_dibits(rng, 0x293, blocks) + [int(v) for v in rng.integers(0, 4, 30)]
    if lsm:
        return modulate_cqpsk(np.array(dibits), 48000, 4800, seed=seed)
    return modulate_c4fm(dibits, 48000, snr_db=22.0, cfo_hz=45.0, timing=0.37, seed=seed)


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in sorted(v.items()) if "time" not in str(k).lower() and "age" not in str(k).lower()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.generic,)):
        return v.item()
    if isinstance(v, bytes):
        return v.hex()
    if hasattr(v, "value") and hasattr(v, "name"):
        return str(v.name)
    return v


@pytest.mark.parametrize("modulation,chunk", [("c4fm", 12000), ("c4fm", 24000), ("lsm", 12000)])
def test_control_channel_monitor_decodes_the_same_tsbks(native, modulation, chunk):
    build_ref.load()
    import wavecapsdr.trunking.control_channel as cc
    from wavecapsdr.decoders.p25_frames import DATA_DEINTERLEAVE
    from wavecapsdr.dsp.fec.trellis import trellis_encode
    from wavecapsdr.trunking.config import TrunkingProtocol
    import wavecap_sdr_b200.install as b200

    lsm = modulation == "lsm"   # simulcast sites: CQPSK demodulator (control_channel.py:135-144)
    iq = _signal(trellis_encode, list(DATA_DEINTERLEAVE), seed=3, lsm=lsm)

    def run():
        mon = cc.ControlChannelMonitor(protocol=TrunkingProtocol.P25_PHASE1, sample_rate=48000,
                                        modulation=cc.P25Modulation.LSM if lsm else cc.P25Modulation.C4FM)
        raw = []
        mon.on_tsbk = lambda b: raw.append(bytes(b))
        results = []
        for s0 in range(0, len(iq), chunk):
            results += mon.process_iq(iq[s0:s0 + chunk])
        diag = {"duid": dict(mon._diag_duid_histogram), "bch": list(mon._diag_bch_errors), "trellis": list(mon._diag_trellis_metrics),
                "frames": mon.frames_decoded, "tsbk": mon.tsbk_decoded, "attempts": mon.tsbk_attempts, "crc_pass": mon.tsbk_crc_pass,
                "error_sum": mon.tsbk_error_sum, "rejected": mon.tsbk_rejected, "sync_losses": mon.sync_losses,
                "state": mon.sync_state.value, "demod": type(mon._demod).__module__}
        return _plain(results), raw, _plain(diag)

    ref_results, ref_raw, ref_diag = run()
    # the CPU chain decodes the signal (its own demodulator leaves 3-10 bit errors per NID on this synthetic channel and its
    # monitor drops frames at chunk boundaries: what it manages is the yardstick, not what was sent)
    assert ref_diag["demod"].startswith("wavecapsdr") and ref_diag["crc_pass"] >= 4 and ref_diag["attempts"] >= 15, ref_diag
    names = b200.install(0)
    try:
        assert "wavecapsdr.trunking.control_channel.DSPC4FMDemodulator (alias)" in names
        got_results, got_raw, got_diag = run()
    finally:
        b200.uninstall()
    assert got_diag.pop("demod").startswith("wavecap_sdr_b200")
    ref_diag.pop("demod")
    if lsm:
        # the CQPSK slicer of the reference goes through numpy's SVML arctan2: a symbol within ~1e-6 rad of a decision
        # boundary may fall on the other side there; the trellis absorbs such a dibit, so the per-block error
        # metrics may differ by that one dibit while everything decoded stays identical
        for key in ("bch", "trellis"):
            a, b = ref_diag.pop(key), got_diag.pop(key)
            assert len(a) == len(b) and sum(abs(x - y) for x, y in zip(a, b)) <= 2, (key, a, b)
    assert got_diag == ref_diag, (ref_diag, got_diag)
    assert got_raw == ref_raw
    assert got_results == ref_results
    parity_note(f"reference ControlChannelMonitor.process_iq ({modulation}), chunks of {chunk}: {ref_diag['frames']} frames, {ref_diag['attempts']} TSBK "
                f"blocks, {ref_diag['crc_pass']} CRC passes, {len(ref_results)} parsed results — identical after install() "
                f"(counters, BCH / trellis diagnostics, raw TSBK bytes, parsed fields)")


def test_control_channel_scanner_class_of_the_reference_after_install(native):
    """`trunking.cc_scanner.ControlChannelScanner` (what `TrunkingSystem` builds, system.py:997, and drives with
    `scan_all` / `log_scan_results` / `get_best_channel` / `should_roam` / `get_stats`, :1621-1708, :2767): the reference's
    class on the box's CPU against the class install() puts in its place, same band — dB values within 2e-4 dB, sync
    decisions, sample counts, best channel, ranking and roam decision identical."""
    from oracle import cc_scanner as oc

    build_ref.load()
    import wavecapsdr.trunking.cc_scanner as rs
    import wavecapsdr.trunking.system as rsys
    import wavecap_sdr_b200.install as b200

    x, center, freqs = oc.synth_band()

    def run():
        sc = rs.ControlChannelScanner(center_hz=center, sample_rate=1_200_000, control_channels=list(freqs))
        m = sc.scan_all(x)
        sc.log_scan_results()
        sc._current_channel_hz = freqs[3]
        return type(sc).__module__, m, sc.get_best_channel()[0], [f for f, _ in sc.get_channel_ranking()], sc.should_roam(freqs[3]), sc.get_stats()

    mod0, m0, best0, rank0, roam0, st0 = run()
    assert mod0.startswith("wavecapsdr")
    names = b200.install(0)
    try:
        assert "wavecapsdr.trunking.system.ControlChannelScanner (alias)" in names
        assert rsys.ControlChannelScanner.__module__.startswith("wavecap_sdr_b200")
        mod1, m1, best1, rank1, roam1, st1 = run()
    finally:
        b200.uninstall()
    assert mod1.startswith("wavecap_sdr_b200") and sorted(m1) == sorted(m0)
    worst = 0.0
    for f, a in m0.items():
        b = m1[f]
        for k in ("power_db", "peak_power_db", "noise_floor_db", "snr_db"):
            worst = max(worst, abs(getattr(a, k) - getattr(b, k)))
        assert a.sync_detected == b.sync_detected and a.sample_count == b.sample_count, f
    assert worst < 2e-4 and (best1, rank1, roam1) == (best0, rank0, roam0)
    assert st1["channels_measured"] == st0["channels_measured"] and sorted(st1["measurements"]) == sorted(st0["measurements"])
    assert st1["current_channel_hz"] == st0["current_channel_hz"]
    parity_note(f"reference ControlChannelScanner ({len(m0)} candidates at 1.2 MS/s): CPU class vs the installed one, worst |dB| difference "
                f"{worst:.1e}, sync decisions / best channel / ranking / roam decision identical")
