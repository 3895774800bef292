"""GPU soft sync detector with the call surface of `wavecapsdr.decoders.p25_framer.P25P1SoftSyncDetector`
(decoders/p25_framer.py:125-231): sliding 24-symbol correlation against the P25 sync word 0x5575F5FF77FF.
`process_batch` is one launch (csrc/p25.cu `wc_c4fm_sync_scores`); the message assembler / NID BCH / TSBK trellis
above it stay the reference's Python."""
from __future__ import annotations

import numpy as np

from ..dsp.p25.c4fm import _SoftSyncDetector


class P25P1SoftSyncDetector:
    SYNC_PATTERN = 0x5575F5FF77FF

    def __init__(self) -> None:
        self._det = _SoftSyncDetector()
        self.SYNC_PATTERN_SYMBOLS = np.array(
            [3.0 if ((self.SYNC_PATTERN >> ((23 - i) * 2)) & 3) == 1 else -3.0 for i in range(24)], dtype=np.float32)

    def reset(self) -> None:
        self._det.reset()

    def process(self, soft_symbol: float) -> float:
        return self._det.process(soft_symbol)

    def process_batch(self, soft_symbols) -> np.ndarray:
        s = np.asarray(soft_symbols, dtype=np.float32)
        if s.size == 0:
            return np.array([], dtype=np.float32)
        return self._det.process_block(s).astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# Message framer (decoders/p25_framer.py:33-92, 352-849) on the GPU: csrc/p25frame.cu runs the whole
# P25P1MessageFramer state machine, one thread per channel, with the NID BCH decode inside the kernel.
# ---------------------------------------------------------------------------------------------------
import ctypes as _C
import logging as _logging
from dataclasses import dataclass as _dataclass
from enum import IntEnum as _IntEnum
from typing import Callable as _Callable

from .. import _native as _N

_log = _logging.getLogger(__name__)

_MESSAGE_BITS = {0x0: 648, 0x3: 28, 0x5: 1568, 0xA: 1568, 0x7: 196, 0x17: 392, 0x27: 588, 0xC: 196, 0x1C: 392,
                 0x2C: 588, 0x3C: 784, 0x4C: 980, 0x5C: 1176, 0xF: 168, 0xD: 2000}


class P25P1DataUnitID(_IntEnum):
    """DUID values of the reference enum (p25_framer.py:33-56), same names."""

    HEADER_DATA_UNIT = 0x0
    TERMINATOR_DATA_UNIT = 0x3
    LOGICAL_LINK_DATA_UNIT_1 = 0x5
    TRUNKING_SIGNALING_BLOCK_1 = 0x7
    LOGICAL_LINK_DATA_UNIT_2 = 0xA
    PACKET_DATA_UNIT = 0xC
    TERMINATOR_DATA_UNIT_LINK_CONTROL = 0xF
    UNKNOWN = 0xE
    PLACE_HOLDER = 0xD
    TRUNKING_SIGNALING_BLOCK_2 = 0x17
    TRUNKING_SIGNALING_BLOCK_3 = 0x27
    PACKET_DATA_UNIT_BLOCK_1 = 0x1C
    PACKET_DATA_UNIT_BLOCK_2 = 0x2C
    PACKET_DATA_UNIT_BLOCK_3 = 0x3C
    PACKET_DATA_UNIT_BLOCK_4 = 0x4C
    PACKET_DATA_UNIT_BLOCK_5 = 0x5C

    @classmethod
    def from_value(cls, value: int) -> "P25P1DataUnitID":
        return cls(value) if value in cls._value2member_map_ else cls.UNKNOWN

    def get_message_length(self) -> int:
        return _MESSAGE_BITS.get(int(self), 196)

    def get_elapsed_dibit_length(self) -> int:
        return 57 + self.get_message_length() // 2


@_dataclass
class P25P1Message:
    """Assembled message, fields as the reference dataclass (p25_framer.py:351-360)."""

    duid: P25P1DataUnitID
    nac: int
    timestamp: int
    bits: np.ndarray
    corrected_bit_count: int = 0
    valid: bool = True


_ERR_TEXT = {
    1: lambda a, b, d: "Cannot dispatch placeholder message",
    2: lambda a, b, d: f"P25 {d} length {a} below minimum {b}",
    3: lambda a, b, d: f"P25 {d} length {a} is not aligned to 196-bit blocks",
    4: lambda a, b, d: f"P25 {d} length {a} did not match expected {b}",
    5: lambda a, b, d: f"Invalid dibit {a} for DUID {d}",
    6: lambda a, b, d: "p25 framer output buffers full",
}


def _duid_name(v: int) -> str:
    try:
        return P25P1DataUnitID(v).name
    except ValueError:
        return hex(v)


class P25FramerBank:
    """C independent framers advanced by one call. Inputs are [C][n] float32 soft symbols and uint8 dibits (numpy, or
    torch CUDA tensors straight from `C4FMBank.demodulate(..., return_device=True)`), `n_sym` the valid count per row.
    `process_batch` returns, per channel, (messages, nid_count, error, error_pos): messages are (duid, nac,
    symbols_total_at_dispatch, bits uint8, corrected_bit_count) tuples, error is None or the AssertionError text the
    reference would have raised out of `process_batch`, error_pos the symbol index it was raised on (-1)."""

    def __init__(self, n_channels: int) -> None:
        _N.ensure_init()
        self.n_channels = int(n_channels)
        h = _C.c_void_p()
        _N.check(_N.lib().wc_p25framer_create(self.n_channels, _C.byref(h)))
        self._h = h
        self.last_scores: np.ndarray | None = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _N.lib().wc_p25framer_destroy(h)
            except Exception:
                pass
            self._h = None

    def reset(self, channel: int = -1, full: bool = False) -> None:
        _N.check(_N.lib().wc_p25framer_reset(self._h, int(channel), 1 if full else 0))

    def get_state(self, channel: int = 0) -> dict:
        st = np.zeros(12, dtype=np.int32)
        _N.check(_N.lib().wc_p25framer_get_state(self._h, int(channel), _N.np_ptr(st)))
        keys = ("sync_detected", "nid_pointer", "dibit_counter", "status_symbol_counter", "assembler_active",
                "assembler_duid", "assembler_bits", "detected_nac", "detected_duid", "tracked_nac", "previous_duid",
                "symbols_total")
        return dict(zip(keys, (int(v) for v in st)))

    def process_batch(self, soft_symbols, dibits, n_sym=None, mode: int = 0, dispatch_enabled: bool = True,
                      want_scores: bool = False):
        l = _N.lib()
        C = self.n_channels
        if _N.is_torch_cuda(soft_symbols):
            return self._process_device(soft_symbols, dibits, n_sym, mode, dispatch_enabled)
        soft = np.ascontiguousarray(np.asarray(soft_symbols, dtype=np.float32).reshape(C, -1))
        dib = np.ascontiguousarray(np.asarray(dibits).astype(np.uint8).reshape(C, -1))
        if soft.shape != dib.shape:
            raise ValueError("soft symbols and dibits differ in shape")
        n = soft.shape[1]
        ns = None if n_sym is None else np.ascontiguousarray(np.asarray(n_sym, dtype=np.int32).reshape(C))
        mm, pc = l.wc_p25framer_max_msgs(n), l.wc_p25framer_pool_bytes(n)
        hdr = np.zeros((C, mm, 6), dtype=np.int32)
        hsym = np.zeros((C, mm), dtype=np.int64)
        pool = np.zeros((C, pc), dtype=np.uint8)
        summ = np.zeros((C, 8), dtype=np.int32)
        scores = np.zeros((C, max(n, 1)), dtype=np.float32) if want_scores else None
        _N.check(l.wc_p25framer_process_host(self._h, _N.np_ptr(soft), _N.np_ptr(dib), n,
                                             _N.np_ptr(ns) if ns is not None else None, int(mode),
                                             1 if dispatch_enabled else 0,
                                             _N.np_ptr(scores) if scores is not None else None, _N.np_ptr(hdr),
                                             _N.np_ptr(hsym), _N.np_ptr(pool), _N.np_ptr(summ)))
        self.last_scores = scores[:, :n] if scores is not None else None
        return self._unpack(hdr, hsym, pool, summ)

    def _process_device(self, soft, dibits, n_sym, mode, dispatch_enabled):
        import torch

        l = _N.lib()
        C = self.n_channels
        soft = soft.reshape(C, -1)
        dibits = dibits.reshape(C, -1)
        if soft.dtype != torch.float32 or dibits.dtype != torch.uint8 or not soft.is_contiguous() or not dibits.is_contiguous():
            raise ValueError("device inputs must be contiguous float32 / uint8 [C][n]")
        n = soft.shape[1]
        ns = None
        if n_sym is not None:
            ns = n_sym if _N.is_torch_cuda(n_sym) else torch.as_tensor(np.asarray(n_sym, dtype=np.int32), device=soft.device)
            ns = ns.to(torch.int32).contiguous()
        mm, pc = l.wc_p25framer_max_msgs(n), l.wc_p25framer_pool_bytes(n)
        hdr = torch.zeros((C, mm, 6), dtype=torch.int32, device=soft.device)
        hsym = torch.zeros((C, mm), dtype=torch.int64, device=soft.device)
        pool = torch.zeros((C, pc), dtype=torch.uint8, device=soft.device)
        summ = torch.zeros((C, 8), dtype=torch.int32, device=soft.device)
        _N.check(l.wc_p25framer_process(self._h, _C.c_void_p(soft.data_ptr()), _C.c_void_p(dibits.data_ptr()), n,
                                        _C.c_void_p(ns.data_ptr()) if ns is not None else None, n, int(mode),
                                        1 if dispatch_enabled else 0, None, _C.c_void_p(hdr.data_ptr()),
                                        _C.c_void_p(hsym.data_ptr()), _C.c_void_p(pool.data_ptr()),
                                        _C.c_void_p(summ.data_ptr()), _N.torch_stream_ptr()))
        return self._unpack(hdr.cpu().numpy(), hsym.cpu().numpy(), pool.cpu().numpy(), summ.cpu().numpy())

    @staticmethod
    def _unpack(hdr, hsym, pool, summ):
        out = []
        for c in range(hdr.shape[0]):
            msgs = []
            for m in range(int(summ[c, 0])):
                duid, nac, nbits, corrected, off, _ = (int(v) for v in hdr[c, m])
                msgs.append((duid, nac, int(hsym[c, m]), pool[c, off:off + nbits].copy(), corrected))
            err = None
            code = int(summ[c, 2])
            if code:
                err = _ERR_TEXT[code](int(summ[c, 4]), int(summ[c, 5]), _duid_name(int(summ[c, 6])))
            out.append((msgs, int(summ[c, 1]), err, int(summ[c, 3])))
        return out


class P25P1MessageFramer:
    """Single-channel framer with the reference's call surface (p25_framer.py:363-849). Messages reach the listener
    in dispatch order; when the reference would raise AssertionError out of the batch, the messages dispatched before
    that point are delivered first and the same error is raised."""

    DIBIT_LENGTH_NID = 33
    SYNC_DETECTION_THRESHOLD = 60.0

    def __init__(self) -> None:
        self._bank = P25FramerBank(1)
        self._message_listener: _Callable[[P25P1Message], None] | None = None
        self._running = False
        self._reference_timestamp = 0
        self._ts_base_symbols = 0
        # host-side hook of the reference's white-box tests (tests/test_p25_message_assertions.py): an assembler placed
        # here is dispatched by _dispatch_message() / _dispatch_tsbk() / _dispatch_pdu() below with the reference's length
        # rules. The batch path keeps its assemblers on the device (csrc/p25frame.cu) and never touches this attribute.
        self._message_assembler = None
        self._previous_duid = P25P1DataUnitID.PLACE_HOLDER
        self._detected_sync_bit_errors = 0

    # ---- dispatch of a host-side assembler (p25_framer.py:655-826) -----------------------------------------------
    _TSBK = (0x7, 0x17, 0x27)
    _PDU = (0xC, 0x1C, 0x2C, 0x3C, 0x4C, 0x5C)

    def _get_timestamp(self) -> int:
        return self._timestamp(self._bank.get_state(0)["symbols_total"])

    @staticmethod
    def _assert_message_length(bits, duid, allow_truncated: bool = False) -> None:
        code, name = int(duid), _duid_name(int(duid))
        if code == int(P25P1DataUnitID.PLACE_HOLDER):
            raise AssertionError(_ERR_TEXT[1](0, 0, name))
        have, want = int(np.asarray(bits).size), _MESSAGE_BITS.get(code, 196)
        if allow_truncated and have < want:
            return
        if code in P25P1MessageFramer._TSBK or code in P25P1MessageFramer._PDU:
            if have < want:
                raise AssertionError(_ERR_TEXT[2](have, want, name))
            if have % 196:
                raise AssertionError(_ERR_TEXT[3](have, want, name))
        elif have != want:
            raise AssertionError(_ERR_TEXT[4](have, want, name))

    def _broadcast(self, message: "P25P1Message") -> None:
        if self._running and self._message_listener is not None:
            try:
                self._message_listener(message)
            except Exception as e:
                _log.error(f"Error in message listener: {e}")

    def _emit(self, duid, bits, corrected: int) -> None:
        a = self._message_assembler
        self._broadcast(P25P1Message(duid=P25P1DataUnitID.from_value(int(duid)), nac=a.nac, timestamp=self._get_timestamp(),
                                     bits=bits, corrected_bit_count=corrected))

    def _dispatch_message(self) -> None:
        a = self._message_assembler
        if a is None:
            return
        self._previous_duid = a.duid
        if not self._running or self._message_listener is None:
            self._message_assembler = None
            return
        code, truncated = int(a.duid), bool(a.was_force_completed())
        if code in self._TSBK:
            self._dispatch_tsbk(allow_truncated=truncated)
        elif code in self._PDU:
            self._dispatch_pdu(allow_truncated=truncated)
        elif code == int(P25P1DataUnitID.PLACE_HOLDER):
            self._message_assembler = None
        else:
            self._dispatch_other(allow_truncated=truncated)

    def _dispatch_other(self, allow_truncated: bool = False) -> None:
        a = self._message_assembler
        if a is None:
            return
        bits = a.get_message_bits()
        self._assert_message_length(bits, a.duid, allow_truncated=allow_truncated)
        self._emit(a.duid, bits, self._detected_sync_bit_errors)
        self._message_assembler = None

    _dispatch_pdu = _dispatch_other   # one message with every collected block (p25_framer.py:808-826)

    def _dispatch_tsbk(self, allow_truncated: bool = False) -> None:
        """Blocks 1..3 of a TSDU leave as separate 196-bit messages; the assembler is re-armed for a continuation block
        when the next one has not arrived yet (p25_framer.py:746-806)."""
        a = self._message_assembler
        if a is None:
            return
        bits = a.get_message_bits()
        self._assert_message_length(bits, a.duid, allow_truncated=allow_truncated)
        block = self._TSBK.index(int(a.duid))
        while True:
            lo, hi = 196 * block, 196 * (block + 1)
            if len(bits) < hi:
                if block == 2:
                    self._message_assembler = None
                return
            self._emit(self._TSBK[block], bits[lo:hi], self._detected_sync_bit_errors if block == 0 else 0)
            if block == 2:
                self._message_assembler = None
                return
            if len(bits) >= hi + 196:
                a.set_duid(type(a.duid)(self._TSBK[block + 1]))
                self._assert_message_length(bits, a.duid, allow_truncated=allow_truncated)
                block += 1
            else:
                a.reconfigure(type(a.duid)(self._TSBK[block + 1]))
                return

    def start(self) -> None:
        self._running = True

    def stop(self) -> None:
        self._running = False

    def set_listener(self, listener: _Callable[[P25P1Message], None]) -> None:
        self._message_listener = listener

    def set_timestamp(self, timestamp: int) -> None:
        self._reference_timestamp = timestamp
        self._ts_base_symbols = self._bank.get_state(0)["symbols_total"]

    def reset(self) -> None:
        self._bank.reset(0, full=False)

    def _timestamp(self, symbols_total: int) -> int:
        if self._reference_timestamp > 0:
            return self._reference_timestamp + int(1000.0 * (symbols_total - self._ts_base_symbols) / 4800)
        return 0

    def _run(self, soft, dibits, mode: int) -> int:
        enabled = self._running and self._message_listener is not None
        (msgs, nids, err, _pos), = self._bank.process_batch(soft, dibits, mode=mode, dispatch_enabled=enabled)
        for duid, nac, sym, bits, corrected in msgs:
            message = P25P1Message(duid=P25P1DataUnitID(duid), nac=nac, timestamp=self._timestamp(sym), bits=bits,
                                   corrected_bit_count=corrected)
            try:
                self._message_listener(message)
            except Exception as e:  # the reference logs listener errors and carries on (p25_framer.py:829-835)
                _log.error(f"Error in message listener: {e}")
        if err is not None:
            raise AssertionError(err)
        return nids

    def process_batch(self, soft_symbols, dibits) -> int:
        d = np.asarray(dibits)
        s = np.asarray(soft_symbols)
        if d.size == 0 or s.size != d.size:  # p25_framer.py:478-480
            return 0
        return self._run(s.reshape(1, -1), d.reshape(1, -1), 0)

    def process_with_soft_sync(self, soft_symbol: float, dibit: int) -> bool:
        return self._run(np.array([[soft_symbol]], dtype=np.float32), np.array([[dibit]], dtype=np.uint8), 1) > 0

    def process(self, dibit: int) -> bool:
        return self._run(np.zeros((1, 1), dtype=np.float32), np.array([[dibit]], dtype=np.uint8), 2) > 0
