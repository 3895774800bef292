"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's P25 Phase 1 message framer.

Source: /root/reference/backend/wavecapsdr/decoders/p25_framer.py
  soft sync scores of a block      :193-231   (np.correlate 'valid' over [last 24 symbols | block], float32)
  process_batch                    :471-509   (sync callback BEFORE the symbol it was detected on)
  _process                         :517-579   (status symbol every 36th dibit, NID after 33 dibits, assembler)
  _check_nid / _nid_detected       :581-649   (status symbol at NID index 11, BCH, NAC tracker, counters 57 / 21)
  assembler + force_completion     :234-318
  NAC tracker                      :320-349
  dispatch + length assertions     :651-827
The state is a flat object and every reference method is a plain function here; AssertionErrors leave the same
partial state behind as the reference (e.g. `sync` stays set when the NID handler raised). Pinned to the live
reference by tests/golden/p25_framer.npz (oracle/make_golden.py:gen_p25_framer).
"""
from __future__ import annotations

import numpy as np

from . import bch as obch

SYNC_WORD = 0x5575F5FF77FF
SYNC_SYMBOLS = np.array([3.0 if ((SYNC_WORD >> ((23 - i) * 2)) & 3) == 1 else -3.0 for i in range(24)], dtype=np.float32)
SYNC_DIBITS = [(SYNC_WORD >> ((23 - i) * 2)) & 3 for i in range(24)]

HDU, TDU, LDU1, TSBK1, LDU2, PDU, TDULC, UNKNOWN, PLACEHOLDER, TSBK2, TSBK3 = 0x0, 0x3, 0x5, 0x7, 0xA, 0xC, 0xF, 0xE, 0xD, 0x17, 0x27
LENGTHS = {HDU: 648, TDU: 28, LDU1: 1568, LDU2: 1568, TSBK1: 196, TSBK2: 392, TSBK3: 588, PDU: 196, 0x1C: 392, 0x2C: 588,
           0x3C: 784, 0x4C: 980, 0x5C: 1176, TDULC: 168, PLACEHOLDER: 2000}
NAMES = {HDU: "HEADER_DATA_UNIT", TDU: "TERMINATOR_DATA_UNIT", LDU1: "LOGICAL_LINK_DATA_UNIT_1",
         TSBK1: "TRUNKING_SIGNALING_BLOCK_1", LDU2: "LOGICAL_LINK_DATA_UNIT_2", PDU: "PACKET_DATA_UNIT",
         TDULC: "TERMINATOR_DATA_UNIT_LINK_CONTROL", UNKNOWN: "UNKNOWN", PLACEHOLDER: "PLACE_HOLDER",
         TSBK2: "TRUNKING_SIGNALING_BLOCK_2", TSBK3: "TRUNKING_SIGNALING_BLOCK_3", 0x1C: "PACKET_DATA_UNIT_BLOCK_1",
         0x2C: "PACKET_DATA_UNIT_BLOCK_2", 0x3C: "PACKET_DATA_UNIT_BLOCK_3", 0x4C: "PACKET_DATA_UNIT_BLOCK_4",
         0x5C: "PACKET_DATA_UNIT_BLOCK_5"}
TSBKS = (TSBK1, TSBK2, TSBK3)
PDUS = (PDU, 0x1C, 0x2C, 0x3C, 0x4C, 0x5C)
NID_VALUES = (HDU, TDU, LDU1, TSBK1, LDU2, PDU, TDULC, UNKNOWN, PLACEHOLDER)


def length_of(duid: int) -> int:
    return LENGTHS.get(duid, 196)


class FramerOracle:
    def __init__(self):
        self.hist = np.zeros(24, dtype=np.float32)
        self.seen: dict[int, int] = {}
        self.tracked = 0
        self.symbols_total = 0
        self.detected_errs = 0
        self.ref_ts = 0
        self.ts_base = 0
        self.enabled = True
        self.out: list[tuple] = []
        self.reset()

    def reset(self):  # :837-849
        self.hist[:] = 0
        self.sync = False
        self.nid: list[int] = []
        self.nid_ptr = 0
        self.dibit_counter = 58
        self.status_counter = 36
        self.asm = None  # dict(nac, duid, bits, target, forced)
        self.assembly_required = False
        self.previous_duid = PLACEHOLDER
        self.detected_duid = PLACEHOLDER
        self.detected_nac = 0

    # -- scores --
    def scores(self, soft: np.ndarray) -> np.ndarray:
        soft = np.asarray(soft, dtype=np.float32)
        n = len(soft)
        ext = np.concatenate([self.hist, soft])
        sc = np.correlate(ext, SYNC_SYMBOLS, mode="valid")[-n:] if n else np.zeros(0, np.float32)
        self.hist = ext[-24:].astype(np.float32)
        return sc.astype(np.float32)

    # -- timestamps --
    def timestamp(self) -> int:
        return self.ref_ts + int(1000.0 * (self.symbols_total - self.ts_base) / 4800) if self.ref_ts > 0 else 0

    # -- dispatch --
    def _emit(self, duid, bits, corrected):
        if self.enabled:
            self.out.append((duid, self.asm["nac"], self.timestamp(), np.array(bits, dtype=np.uint8), corrected))

    def _check_length(self, nbits, duid, allow):
        if duid == PLACEHOLDER:
            raise AssertionError("Cannot dispatch placeholder message")
        exp = length_of(duid)
        if allow and nbits < exp:
            return
        if duid in TSBKS or duid in PDUS:
            if nbits < exp:
                raise AssertionError(f"P25 {NAMES[duid]} length {nbits} below minimum {exp}")
            if nbits % 196 != 0:
                raise AssertionError(f"P25 {NAMES[duid]} length {nbits} is not aligned to 196-bit blocks")
            return
        if nbits != exp:
            raise AssertionError(f"P25 {NAMES[duid]} length {nbits} did not match expected {exp}")

    def _dispatch(self):
        a = self.asm
        if a is None:
            return
        self.previous_duid = a["duid"]
        if not self.enabled:
            self.asm = None
            return
        allow = a["forced"]
        if a["duid"] in TSBKS:
            while True:
                duid, bits = a["duid"], a["bits"]
                self._check_length(len(bits), duid, allow)
                block = TSBKS.index(duid)
                if block < 2:
                    if len(bits) >= 196 * (block + 1):
                        self._emit(duid, bits[196 * block:196 * (block + 1)], self.detected_errs if block == 0 else 0)
                        a["duid"] = TSBKS[block + 1]
                        a["target"] = length_of(a["duid"])
                        if len(bits) >= 196 * (block + 2):
                            continue
                    return
                if len(bits) >= 588:
                    self._emit(duid, bits[392:588], 0)
                self.asm = None
                return
        if a["duid"] == PLACEHOLDER:
            self.asm = None
            return
        self._check_length(len(a["bits"]), a["duid"], allow)
        self._emit(a["duid"], a["bits"], self.detected_errs)
        self.asm = None

    def _receive(self, dibit):
        a = self.asm
        if dibit < 0 or dibit > 3:
            raise AssertionError(f"Invalid dibit {dibit} for DUID {NAMES[a['duid']]}")
        if len(a["bits"]) < a["target"]:
            a["bits"].append((dibit >> 1) & 1)
            if len(a["bits"]) < a["target"]:
                a["bits"].append(dibit & 1)

    def _force(self, next_duid):
        a = self.asm
        size = len(a["bits"])
        a["forced"] = True
        if a["duid"] == PLACEHOLDER:
            if size <= 28:
                a["duid"] = TDU
            elif next_duid == LDU1:
                if size <= 770:
                    a["duid"] = HDU
                elif size >= 1500:
                    a["duid"] = LDU2
            elif next_duid == LDU2:
                if size >= 1500:
                    a["duid"] = LDU1
            elif next_duid == TSBK1:
                if size >= 195:
                    a["duid"] = TSBK1
        if a["duid"] == PLACEHOLDER:
            a["duid"] = TDU
        a["target"] = length_of(a["duid"])

    # -- NID --
    def _check_nid(self) -> bool:
        d32 = self.nid[:11] + self.nid[12:33]
        bits = []
        for d in d32[:32]:
            bits += [(d >> 1) & 1, d & 1]
        data, errs = obch.bch_decode(bits[:63], self.tracked if self.tracked else None)
        if errs < 0:
            return False
        nac, dv = (data >> 4) & 0xFFF, data & 0xF
        duid = dv if dv in NID_VALUES else UNKNOWN
        if 0x001 <= nac <= 0xFFE:
            self.seen[nac] = self.seen.get(nac, 0) + 1
            if self.seen[nac] >= 3:
                self.tracked = nac
        self.detected_duid = PLACEHOLDER if duid == UNKNOWN else duid
        self.detected_nac = nac
        self.detected_errs = errs
        if self.asm is not None:
            if len(self.asm["bits"]) >= self.asm["target"]:
                if self.asm["duid"] != PLACEHOLDER:
                    self._dispatch()
            else:
                self._force(self.detected_duid)
                self._dispatch()
        self.assembly_required = True
        self.dibit_counter = 57
        self.status_counter = 21
        return True

    # -- one symbol --
    def step(self, dibit: int) -> bool:
        valid = False
        self.symbols_total += 1
        self.status_counter += 1
        if self.sync:
            self.nid.append(dibit)
            self.nid_ptr += 1
            if self.nid_ptr >= 33:
                valid = self._check_nid()
                self.sync = False
        if self.status_counter == 36:
            self.status_counter = 0
            self.dibit_counter += 1
            return False
        if self.asm is not None:
            if len(self.asm["bits"]) >= self.asm["target"]:
                self._dispatch()
                if self.asm is not None:
                    self._receive(dibit)
            else:
                self._receive(dibit)
        elif self.dibit_counter == 57:
            if self.assembly_required:
                self.asm = dict(nac=self.detected_nac, duid=self.detected_duid, bits=[], target=length_of(self.detected_duid), forced=False)
                self.assembly_required = False
            elif self.detected_nac > 0:
                self.detected_duid = PLACEHOLDER
                self.asm = dict(nac=self.detected_nac, duid=PLACEHOLDER, bits=[], target=2000, forced=False)
        elif self.dibit_counter >= 4800:
            self.dibit_counter -= 4800
        self.dibit_counter += 1
        return valid

    def _sync_hit(self):
        self.sync = True
        self.nid_ptr = 0
        self.nid = []

    def process_batch(self, soft, dibits) -> int:
        """Returns the NID count; messages accumulate in self.out. AssertionError propagates like the reference."""
        n = len(dibits)
        if n == 0 or len(soft) != n:
            return 0
        hits = self.scores(soft) > 60.0
        count = 0
        for i in range(n):
            if hits[i]:
                self._sync_hit()
            if self.step(int(dibits[i])):
                count += 1
        return count


# ---- synthetic frame source (valid NIDs via oracle.bch.bch_encode; status symbol after every 35 dibits) ----

def frame_dibits(rng, nac: int, duid: int, payload_bits: int, nid_errors: int = 0) -> list[int]:
    nid = list(obch.bch_encode((nac << 4) | duid))
    nid.append(sum(nid) & 1)
    if nid_errors:
        for p in rng.choice(63, nid_errors, replace=False):
            nid[p] ^= 1
    body = SYNC_DIBITS + [(nid[2 * i] << 1) | nid[2 * i + 1] for i in range(32)] + [int(v) for v in rng.integers(0, 4, payload_bits // 2)]
    out = []
    for i, d in enumerate(body):
        out.append(int(d))
        if (i + 1) % 35 == 0:
            out.append(int(rng.integers(0, 4)))
    return out


def dibits_to_soft(dibits) -> np.ndarray:
    return np.array([1.0, 3.0, -1.0, -3.0], dtype=np.float32)[np.asarray(dibits)]


def _stream(self, soft, dibits, mode: int) -> int:
    """mode 1: process_with_soft_sync per symbol (:438-457, sync callback AFTER the symbol); mode 2: process (:459-469)."""
    count = 0
    for s, d in zip(np.asarray(soft, dtype=np.float32), dibits):
        if self.step(int(d)):
            count += 1
        if mode == 1 and float(self.scores(np.array([s], dtype=np.float32))[0]) > 60.0:
            self._sync_hit()
    return count


FramerOracle.process_stream = _stream
