// P25 Phase 1 framing on the GPU, batched over channels:
//   * BCH(63,16,23) decoding of the Network ID      — dsp/fec/bch.py:52-222 (syndromes, Berlekamp-Massey, Chien),
//                                                      :533-641 (two-pass decode with the tracked NAC)
//   * soft sync correlation over a block of symbols  — decoders/p25_framer.py:193-231
//   * the message framer state machine               — decoders/p25_framer.py:363-849 (P25P1MessageFramer),
//                                                      :234-318 (assembler), trunking NAC tracker :320-349
//
// Layout. One handle owns C channels; all state lives in device memory (FramerState[C]). A call consumes
// soft/dibit rows [C][stride] (n_sym[c] valid symbols per row — exactly what the C4FM bank leaves on the device),
// runs (1) a grid-wide score kernel (one thread per symbol, 24-tap correlation, previous 24 symbols carried),
// (2) the sequential per-channel machine, one thread per channel: status-symbol stripping, NID collection,
// in-thread BCH decode, message assembly. Assembled messages go to a per-channel header list + bit pool that
// the host turns into P25P1Message objects. The reference raises AssertionError out of the middle of a batch in
// some flows; the machine stops that channel at the same symbol with the same partial state and reports a code.
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/wcsdr_b200.h"
#include "common.cuh"

namespace wc {

// ---- GF(2^6), primitive polynomial x^6 + x + 1 (bch.py:246,255-271) ----
// pw is doubled (alpha^i for i < 126) so that a product is pw[lg[a] + lg[b]] without a modulo. The tables are copied
// into shared memory by every kernel that decodes: lanes look up different elements, which a constant bank serialises.
__constant__ unsigned char c_gf_pw[128];
__constant__ unsigned char c_gf_lg[64];
__constant__ float c_fsync[24];
// syndrome table: g_syn[b][v] = XOR over the set bits p = 8b+k of byte value v of the packed odd syndromes
// (alpha^p, alpha^3p, ..., alpha^21p): .x/.y = S1..S19 in 6-bit fields, .z = S21. 8 x 256 x 16 B = 32 KB, L1/L2 resident.
__device__ uint4 g_syn[8 * 256];

constexpr int BCH_N = 63, BCH_T = 11;

struct Gf {
    const unsigned char* pw;
    const unsigned char* lg;
    __device__ __forceinline__ int mul(int a, int b) const { return (a && b) ? pw[lg[a] + lg[b]] : 0; }
    __device__ __forceinline__ int mul_log(int a, int log_b) const { return a ? pw[lg[a] + log_b] : 0; }
    __device__ __forceinline__ int sqr(int a) const { return a ? pw[2 * lg[a]] : 0; }
};

__device__ __forceinline__ void gf_load_shared(unsigned char* s_pw, unsigned char* s_lg) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_pw[i] = c_gf_pw[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_lg[i] = c_gf_lg[i];
}

// odd syndromes S1, S3, ..., S21 of r(x) = sum_i cw[i] x^(62-i) (bch.py:195-222); w holds cw[i] in bit (62-i)
__device__ __forceinline__ uint4 bch_odd_syndromes(unsigned long long w) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const uint4 t = g_syn[b * 256 + (int)((w >> (8 * b)) & 255ull)];
        acc.x ^= t.x;
        acc.y ^= t.y;
        acc.z ^= t.z;
    }
    return acc;
}

// One decoding attempt (bch.py:575-641). Returns the number of corrected bits or -1; *data = first 16 bits.
// Every array below is indexed with compile-time constants (loops fully unrolled), so it lives in registers:
// the locator update keeps x^m * B(x) instead of B(x) and m, which turns C[i+m] ^= q * B[i] into C[i] ^= q * Bx[i].
__device__ int bch_decode_once(const Gf gf, unsigned long long w, int* data) {
    const uint4 odd = bch_odd_syndromes(w);
    if ((odd.x | odd.y | odd.z) == 0u) {
        *data = (int)(w >> 47) & 0xFFFF;
        return 0;
    }
    // S[k] = S_(k+1): odd ones from the table, even ones by squaring (S_2j = S_j^2 in characteristic 2)
    int S[2 * BCH_T];
    {
        const unsigned long long lo = ((unsigned long long)odd.y << 32) | odd.x;
#pragma unroll
        for (int j = 0; j < 10; ++j) S[2 * j] = (int)((lo >> (6 * j)) & 63ull);
        S[20] = (int)(odd.z & 63u);
#pragma unroll
        for (int k = 2; k <= 2 * BCH_T; k += 2) S[k - 1] = gf.sqr(S[k / 2 - 1]);
    }
    // Berlekamp-Massey (bch.py:52-127), locator truncated to degree T like the reference
    int Cp[BCH_T + 1], Bx[BCH_T + 1], Sw[BCH_T + 1];  // Sw[i] = S[n - i] (0 when n < i)
#pragma unroll
    for (int i = 0; i <= BCH_T; ++i) Cp[i] = Bx[i] = Sw[i] = 0;
    Cp[0] = 1;
    Bx[1] = 1;  // x^1 * B(x), B = 1
    int L = 0, log_b = 0;
#pragma unroll
    for (int n = 0; n < 2 * BCH_T; ++n) {
#pragma unroll
        for (int i = BCH_T; i > 0; --i) Sw[i] = Sw[i - 1];
        Sw[0] = S[n];
        int d = Sw[0];
#pragma unroll
        for (int i = 1; i <= BCH_T; ++i)
            if (i <= L) d ^= gf.mul(Cp[i], Sw[i]);
        if (d == 0) {
#pragma unroll
            for (int i = BCH_T; i > 0; --i) Bx[i] = Bx[i - 1];
            Bx[0] = 0;
        } else {
            const int log_d = gf.lg[d];
            const int log_q = log_d + BCH_N - log_b;  // < 126
            const bool grow = n >= 2 * L;
            int old[BCH_T + 1];
#pragma unroll
            for (int i = 0; i <= BCH_T; ++i) {
                old[i] = Cp[i];
                Cp[i] ^= gf.mul_log(Bx[i], log_q >= BCH_N ? log_q - BCH_N : log_q);
            }
            if (grow) {
                L = n + 1 - L;
                log_b = log_d;
#pragma unroll
                for (int i = BCH_T; i > 0; --i) Bx[i] = old[i - 1];
            } else {
#pragma unroll
                for (int i = BCH_T; i > 0; --i) Bx[i] = Bx[i - 1];
            }
            Bx[0] = 0;
        }
    }
    if (L == 0 || L > BCH_T) return -1;
    // Chien search (bch.py:131-191): root alpha^i <-> error term x^((63-i)%63). term[k] = C[k] * alpha^(i*k) is
    // advanced by a constant factor per step.
    int term[BCH_T + 1];
#pragma unroll
    for (int k = 0; k <= BCH_T; ++k) term[k] = (k <= L) ? Cp[k] : 0;
    int found = 0;
    unsigned long long flips = 0ull;
    for (int i = 0; i < BCH_N && found < L; ++i) {
        int val = 0;
#pragma unroll
        for (int k = 0; k <= BCH_T; ++k) {
            val ^= term[k];
            term[k] = gf.mul_log(term[k], k);
        }
        if (val == 0) {
            flips ^= 1ull << ((BCH_N - i) % BCH_N);
            ++found;
        }
    }
    if (found != L) return -1;
    const unsigned long long fixed = w ^ flips;
    const uint4 chk = bch_odd_syndromes(fixed);
    if ((chk.x | chk.y | chk.z) != 0u) return -1;
    *data = (int)(fixed >> 47) & 0xFFFF;
    return L;
}

// BCH_63_16_23.decode (bch.py:533-573): second attempt with the NAC field overwritten by the tracked NAC
__device__ __noinline__ int bch_decode_nid(const unsigned char* s_pw, const unsigned char* s_lg, unsigned long long w,
                                           int tracked_nac, int* data) {
    Gf gf;
    gf.pw = s_pw;
    gf.lg = s_lg;
    int e = bch_decode_once(gf, w, data);
    if (e >= 0) return e;
    if (tracked_nac > 0) {
        const int cur = (int)(w >> 51) & 0xFFF;
        if (cur != tracked_nac) {
            const unsigned long long w2 = (w & ((1ull << 51) - 1ull)) | ((unsigned long long)(tracked_nac & 0xFFF) << 51);
            return bch_decode_once(gf, w2, data);
        }
    }
    *data = 0;
    return -1;
}

__global__ void bch_batch_kernel(const unsigned char* __restrict__ bits, const int* __restrict__ tracked, int B,
                                 int* __restrict__ data, int* __restrict__ errors) {
    __shared__ unsigned char s_pw[128], s_lg[64];
    gf_load_shared(s_pw, s_lg);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    unsigned long long w = 0ull;
    for (int k = 0; k < BCH_N; ++k) w |= (unsigned long long)(bits[(size_t)i * BCH_N + k] & 1) << (62 - k);
    int d = 0;
    const int e = bch_decode_nid(s_pw, s_lg, w, tracked ? tracked[i] : 0, &d);
    data[i] = (e >= 0) ? d : 0;
    errors[i] = e;
}

// ---- framer state (one per channel) ----
constexpr int DUID_HDU = 0x0, DUID_TDU = 0x3, DUID_LDU1 = 0x5, DUID_TSBK1 = 0x7, DUID_LDU2 = 0xA, DUID_PDU = 0xC,
              DUID_TDULC = 0xF, DUID_UNKNOWN = 0xE, DUID_PLACEHOLDER = 0xD, DUID_TSBK2 = 0x17, DUID_TSBK3 = 0x27;
constexpr int MAX_MSG_BITS = 2000;

// scalars + NID buffer: copied into registers / local memory for the duration of a call
struct FramerCore {
    long long symbols_total;   // every symbol ever processed (timestamps are derived from it on the host)
    int sync_detected, nid_pointer;
    int dibit_counter, status_counter;
    int asm_active, asm_nac, asm_duid, asm_nbits, asm_target, asm_forced;
    int assembly_required, previous_duid, detected_duid, detected_nac, detected_errs;
    int tracked_nac;
};
struct __align__(16) FramerState {
    FramerCore core;
    float hist[24];                // last 24 soft symbols (oldest first)
    unsigned char nid_buf[40];     // dibits collected after a sync hit
    __align__(16) unsigned char asm_bits[MAX_MSG_BITS];
    unsigned char nac_seen[4096];  // observation counts, saturating (NACTracker, p25_framer.py:320-349)
};

__device__ __forceinline__ int duid_from_value(int v) {  // P25P1DataUnitID.from_value, p25_framer.py:57-63
    switch (v) {
        case DUID_HDU: case DUID_TDU: case DUID_LDU1: case DUID_TSBK1: case DUID_LDU2: case DUID_PDU: case DUID_TDULC:
        case DUID_UNKNOWN: case DUID_PLACEHOLDER:
            return v;
        default:
            return DUID_UNKNOWN;
    }
}
__device__ __forceinline__ int duid_length(int d) {  // get_message_length, p25_framer.py:65-84
    switch (d) {
        case DUID_HDU: return 648;
        case DUID_TDU: return 28;
        case DUID_LDU1: case DUID_LDU2: return 1568;
        case DUID_TSBK1: return 196;
        case DUID_TSBK2: return 392;
        case DUID_TSBK3: return 588;
        case DUID_PDU: return 196;
        case 0x1C: return 392;
        case 0x2C: return 588;
        case 0x3C: return 784;
        case 0x4C: return 980;
        case 0x5C: return 1176;
        case DUID_TDULC: return 168;
        case DUID_PLACEHOLDER: return 2000;
        default: return 196;
    }
}
__device__ __forceinline__ bool duid_is_tsbk(int d) { return d == DUID_TSBK1 || d == DUID_TSBK2 || d == DUID_TSBK3; }
__device__ __forceinline__ bool duid_is_pdu(int d) {
    return d == DUID_PDU || d == 0x1C || d == 0x2C || d == 0x3C || d == 0x4C || d == 0x5C;
}

// per-call output cursor of one channel
struct FramerOut {
    int* hdr;               // [max_msgs][6]: duid, nac, nbits, corrected, bit_offset, (symbols_total at dispatch) low 31 bits
    long long* hdr_sym;     // [max_msgs] symbols_total at dispatch
    unsigned char* pool;    // message bits, one byte per bit
    int max_msgs, pool_cap;
    int n_msgs, pool_used;
    int err_code, err_a, err_b, err_duid;
    int dispatch_enabled;
    unsigned char* asm_bits;   // the channel's assembler bits, NID dibits and NAC observation counts stay in global memory
    unsigned char* nac_seen;
    unsigned char* nid_buf;
    const unsigned char* gf_pw;  // GF(64) tables in shared memory
    const unsigned char* gf_lg;
};

enum { ERR_NONE = 0, ERR_PLACEHOLDER = 1, ERR_BELOW_MIN = 2, ERR_NOT_ALIGNED = 3, ERR_LENGTH = 4, ERR_BAD_DIBIT = 5,
       ERR_OUTPUT_FULL = 6 };

__device__ __forceinline__ void emit_msg(FramerCore& s, FramerOut& o, int duid, int first, int count, int corrected) {
    if (!o.dispatch_enabled) return;  // _broadcast: running and a listener are required (p25_framer.py:829-835)
    const int padded = (count + 3) & ~3;
    if (o.n_msgs >= o.max_msgs || o.pool_used + padded > o.pool_cap) {
        if (!o.err_code) o.err_code = ERR_OUTPUT_FULL;
        return;
    }
    int* h = o.hdr + (size_t)o.n_msgs * 6;
    h[0] = duid;
    h[1] = s.asm_nac;
    h[2] = count;
    h[3] = corrected;
    h[4] = o.pool_used;
    h[5] = 0;
    o.hdr_sym[o.n_msgs] = s.symbols_total;
    // word copies: `first` is 0 / 196 / 392 and the pool cursor stays 4-byte aligned, both buffers are padded to words
    const unsigned int* __restrict__ src = reinterpret_cast<const unsigned int*>(o.asm_bits + first);
    unsigned int* __restrict__ dst = reinterpret_cast<unsigned int*>(o.pool + o.pool_used);
#pragma unroll 8
    for (int i = 0; i < padded / 4; ++i) dst[i] = src[i];
    o.pool_used += padded;
    ++o.n_msgs;
}

// _assert_message_length (p25_framer.py:651-688); true = raise
__device__ __forceinline__ bool length_check_fails(FramerOut& o, int nbits, int duid, bool allow_truncated) {
    if (duid == DUID_PLACEHOLDER) {
        o.err_code = ERR_PLACEHOLDER;
        o.err_duid = duid;
        return true;
    }
    const int expected = duid_length(duid);
    if (allow_truncated && nbits < expected) return false;
    if (duid_is_tsbk(duid) || duid_is_pdu(duid)) {
        if (nbits < expected) {
            o.err_code = ERR_BELOW_MIN; o.err_a = nbits; o.err_b = expected; o.err_duid = duid;
            return true;
        }
        if (nbits % 196 != 0) {
            o.err_code = ERR_NOT_ALIGNED; o.err_a = nbits; o.err_b = expected; o.err_duid = duid;
            return true;
        }
        return false;
    }
    if (nbits != expected) {
        o.err_code = ERR_LENGTH; o.err_a = nbits; o.err_b = expected; o.err_duid = duid;
        return true;
    }
    return false;
}

// _dispatch_message and its three flavours (p25_framer.py:690-827); true = raised
__device__ __forceinline__ bool dispatch_message(FramerCore& s, FramerOut& o) {
    if (!s.asm_active) return false;
    s.previous_duid = s.asm_duid;
    if (!o.dispatch_enabled) {
        s.asm_active = 0;
        return false;
    }
    const bool allow = s.asm_forced != 0;
    int duid = s.asm_duid;
    if (duid_is_tsbk(duid)) {
        // _dispatch_tsbk recurses block by block; every level re-checks the length
        for (;;) {
            duid = s.asm_duid;
            if (length_check_fails(o, s.asm_nbits, duid, allow)) return true;
            if (duid == DUID_TSBK1) {
                if (s.asm_nbits >= 196) {
                    emit_msg(s, o, duid, 0, 196, s.detected_errs);
                    s.asm_duid = DUID_TSBK2;
                    s.asm_target = duid_length(DUID_TSBK2);
                    if (s.asm_nbits >= 392) continue;
                }
                return false;
            }
            if (duid == DUID_TSBK2) {
                if (s.asm_nbits >= 392) {
                    emit_msg(s, o, duid, 196, 196, 0);
                    s.asm_duid = DUID_TSBK3;
                    s.asm_target = duid_length(DUID_TSBK3);
                    if (s.asm_nbits >= 588) continue;
                }
                return false;
            }
            if (s.asm_nbits >= 588) emit_msg(s, o, duid, 392, 196, 0);
            s.asm_active = 0;
            return false;
        }
    }
    if (duid == DUID_PLACEHOLDER && !duid_is_pdu(duid)) {
        s.asm_active = 0;
        return false;
    }
    // PDU flavours and everything else: whole message, then the assembler is dropped
    if (length_check_fails(o, s.asm_nbits, duid, allow)) return true;
    emit_msg(s, o, duid, 0, s.asm_nbits, s.detected_errs);
    s.asm_active = 0;
    return false;
}

__device__ __forceinline__ void asm_start(FramerCore& s, int nac, int duid) {
    s.asm_active = 1;
    s.asm_nac = nac;
    s.asm_duid = duid;
    s.asm_nbits = 0;
    s.asm_target = duid_length(duid);
    s.asm_forced = 0;
}
__device__ __forceinline__ void asm_receive(FramerCore& s, FramerOut& o, int dibit) {  // p25_framer.py:251-259
    if (s.asm_nbits < s.asm_target) {
        o.asm_bits[s.asm_nbits++] = (unsigned char)((dibit >> 1) & 1);
        if (s.asm_nbits < s.asm_target) o.asm_bits[s.asm_nbits++] = (unsigned char)(dibit & 1);
    }
}
// force_completion (p25_framer.py:287-317)
__device__ __forceinline__ void asm_force_completion(FramerCore& s, int next_duid) {
    const int size = s.asm_nbits;
    s.asm_forced = 1;
    if (s.asm_duid == DUID_PLACEHOLDER) {
        if (size <= 28) s.asm_duid = DUID_TDU;
        else if (next_duid == DUID_LDU1) {
            if (size <= 770) s.asm_duid = DUID_HDU;
            else if (size >= 1500) s.asm_duid = DUID_LDU2;
        } else if (next_duid == DUID_LDU2) {
            if (size >= 1500) s.asm_duid = DUID_LDU1;
        } else if (next_duid == DUID_TSBK1) {
            if (size >= 195) s.asm_duid = DUID_TSBK1;
        }
    }
    if (s.asm_duid == DUID_PLACEHOLDER) s.asm_duid = DUID_TDU;
    s.asm_target = duid_length(s.asm_duid);
}

// _check_nid + _nid_detected (p25_framer.py:581-649); returns 1 valid NID, 0 none, -1 raised
__device__ __forceinline__ int check_nid(FramerCore& s, FramerOut& o) {
    unsigned long long w = 0ull;
    int k = 0;
    for (int i = 0; i < 33 && k < 32; ++i) {
        if (i == 11) continue;  // status symbol inside the NID
        const int d = o.nid_buf[i] & 3;
        const int b0 = (d >> 1) & 1, b1 = d & 1;
        if (2 * k < BCH_N) w |= (unsigned long long)b0 << (62 - 2 * k);
        if (2 * k + 1 < BCH_N) w |= (unsigned long long)b1 << (62 - (2 * k + 1));
        ++k;
    }
    int decoded = 0;
    const int errs = bch_decode_nid(o.gf_pw, o.gf_lg, w, s.tracked_nac, &decoded);
    if (errs < 0) return 0;
    const int nac = (decoded >> 4) & 0xFFF;
    const int duid = duid_from_value(decoded & 0xF);
    if (nac >= 0x001 && nac <= 0xFFE) {
        unsigned char& cnt = o.nac_seen[nac];
        if (cnt < 255) ++cnt;
        if (cnt >= 3) s.tracked_nac = nac;
    }
    s.detected_duid = (duid == DUID_UNKNOWN) ? DUID_PLACEHOLDER : duid;
    s.detected_nac = nac;
    s.detected_errs = errs;
    if (s.asm_active) {
        if (s.asm_nbits >= s.asm_target) {
            if (s.asm_duid != DUID_PLACEHOLDER && dispatch_message(s, o)) return -1;
        } else {
            asm_force_completion(s, s.detected_duid);
            if (dispatch_message(s, o)) return -1;
        }
    }
    s.assembly_required = 1;
    s.dibit_counter = 57;
    s.status_counter = 21;
    return 1;
}

// _process (p25_framer.py:517-579); returns 1 valid NID, 0, -1 raised
__device__ __forceinline__ int process_symbol(FramerCore& s, FramerOut& o, int dibit) {
    int valid = 0;
    ++s.symbols_total;
    ++s.status_counter;
    if (s.sync_detected) {
        if (s.nid_pointer < 36) o.nid_buf[s.nid_pointer] = (unsigned char)dibit;
        ++s.nid_pointer;
        if (s.nid_pointer >= 33) {
            const int r = check_nid(s, o);
            if (r < 0) return -1;  // raised inside: _sync_detected stays set, like the reference
            valid = r;
            s.sync_detected = 0;
        }
    }
    if (s.status_counter == 36) {
        s.status_counter = 0;
        ++s.dibit_counter;
        return 0;
    }
    if (s.asm_active) {
        if (s.asm_nbits >= s.asm_target) {
            if (dispatch_message(s, o)) return -1;
            if (s.asm_active) {
                if (dibit > 3) { o.err_code = ERR_BAD_DIBIT; o.err_a = dibit; o.err_duid = s.asm_duid; return -1; }
                asm_receive(s, o, dibit);
            }
        } else {
            if (dibit > 3) { o.err_code = ERR_BAD_DIBIT; o.err_a = dibit; o.err_duid = s.asm_duid; return -1; }
            asm_receive(s, o, dibit);
        }
    } else if (s.dibit_counter == 57) {
        if (s.assembly_required) {
            asm_start(s, s.detected_nac, s.detected_duid);
            s.assembly_required = 0;
        } else if (s.detected_nac > 0) {
            s.detected_duid = DUID_PLACEHOLDER;
            asm_start(s, s.detected_nac, DUID_PLACEHOLDER);
        }
    } else if (s.dibit_counter >= 4800) {
        s.dibit_counter -= 4800;
    }
    ++s.dibit_counter;
    return valid;
}

// score[c][k] = sum_i sync[i] * s[k-23+i] (float32 result), previous 24 symbols from the carried history;
// hits[c][k] = score > 60 (SYNC_DETECTION_THRESHOLD). One thread per symbol, whole bank in one grid.
__global__ void framer_score_kernel(const float* __restrict__ soft, long long stride, const int* __restrict__ n_sym, int n_fixed,
                                    const FramerState* __restrict__ st, unsigned char* __restrict__ hits,
                                    float* __restrict__ scores, long long score_stride) {
    const int c = blockIdx.y;
    const int n = n_sym ? n_sym[c] : n_fixed;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* s = soft + (size_t)c * stride;
    const float* hist = st[c].hist;
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const int t = k - 23 + i;
        const float v = (t >= 0) ? s[t] : hist[24 + t];
        acc += (double)__fmul_rn(c_fsync[i], v);
    }
    const float sc = (float)acc;
    hits[(size_t)c * score_stride + k] = sc > 60.0f ? 1 : 0;
    if (scores) scores[(size_t)c * score_stride + k] = sc;
}

struct FramerArgs {
    FramerState* st;
    const float* soft;
    const unsigned char* dibits;
    long long stride;
    const int* n_sym;
    int n_fixed;
    const unsigned char* hits;
    long long hit_stride;
    int mode;              // 0 = process_batch order (sync callback before the symbol), 1 = process_with_soft_sync order,
                           // 2 = process(): no sync detection
    int dispatch_enabled;
    int C;
    int* hdr;              // [C][max_msgs][6]
    long long* hdr_sym;    // [C][max_msgs]
    unsigned char* pool;   // [C][pool_cap]
    int max_msgs, pool_cap;
    int* summary;          // [C][8]: n_msgs, nid_count, err_code, err_pos, err_a, err_b, err_duid, pool_used
};

constexpr int FR_TILE = 256;          // symbols staged per step
constexpr int FR_ROW = FR_TILE + 4;   // row pitch in bytes: 65 words, so the 32 channel rows sit in 32 different banks

// One warp = 32 channels. Dibits and sync hits are staged tile by tile into shared memory with coalesced loads
// (lanes along the symbol axis), then every lane walks its own channel's row; the scalar machine state lives in
// registers for the whole call, only the assembler's bit buffer and the NAC counts are touched in global memory.
__global__ void __launch_bounds__(32) framer_machine_kernel(const FramerArgs a) {
    __shared__ unsigned char sd[32][FR_ROW];
    __shared__ unsigned char sh[32][FR_ROW];
    __shared__ unsigned char s_pw[128], s_lg[64];
    gf_load_shared(s_pw, s_lg);
    __syncwarp();
    const int lane = threadIdx.x;
    const int c0 = blockIdx.x * 32;
    const int c = c0 + lane;
    const bool live = c < a.C;
    const int n = live ? (a.n_sym ? a.n_sym[c] : a.n_fixed) : 0;
    int n_max = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));

    FramerCore s;
    FramerOut o;
    if (live) {
        s = a.st[c].core;
        o.hdr = a.hdr + (size_t)c * a.max_msgs * 6;
        o.hdr_sym = a.hdr_sym + (size_t)c * a.max_msgs;
        o.pool = a.pool + (size_t)c * a.pool_cap;
        o.asm_bits = a.st[c].asm_bits;
        o.nac_seen = a.st[c].nac_seen;
        o.nid_buf = a.st[c].nid_buf;
    }
    o.gf_pw = s_pw;
    o.gf_lg = s_lg;
    o.max_msgs = a.max_msgs;
    o.pool_cap = a.pool_cap;
    o.n_msgs = 0;
    o.pool_used = 0;
    o.err_code = 0;
    o.err_a = o.err_b = o.err_duid = 0;
    o.dispatch_enabled = a.dispatch_enabled;
    int nid_count = 0, err_pos = -1;
    bool stopped = !live;

    for (int base = 0; base < n_max; base += FR_TILE) {
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            const int cr = c0 + r;
            if (cr >= a.C) break;
            const unsigned char* dr = a.dibits + (size_t)cr * a.stride + base;
            const unsigned char* hr = a.hits + (size_t)cr * a.hit_stride + base;
            const int nr = min(FR_TILE, (a.n_sym ? a.n_sym[cr] : a.n_fixed) - base);
#pragma unroll
            for (int k = 0; k < FR_TILE / 32; ++k) {
                const int i = lane + 32 * k;
                if (i < nr) {
                    sd[r][i] = dr[i];
                    sh[r][i] = (a.mode != 2) ? hr[i] : 0;
                }
            }
        }
        __syncwarp();
        if (!stopped) {
            const int lim = min(FR_TILE, n - base);
            for (int i = 0; i < lim; ++i) {
                const bool hit = sh[lane][i] != 0;
                if (a.mode == 0 && hit) {  // _sync_detected_callback (p25_framer.py:511-515)
                    s.sync_detected = 1;
                    s.nid_pointer = 0;
                }
                const int r = process_symbol(s, o, sd[lane][i]);
                if (r < 0) {
                    err_pos = base + i;
                    stopped = true;
                    break;
                }
                nid_count += r;
                if (a.mode == 1 && hit) {
                    s.sync_detected = 1;
                    s.nid_pointer = 0;
                }
            }
        }
    }
    if (!live) return;
    a.st[c].core = s;
    // batch order: the detector consumed the whole block before the machine ran (p25_framer.py:485-486).
    // per-symbol order: a symbol whose _process raised never reaches the detector (:448-455). process() never feeds it.
    if (a.mode != 2) {
        const int nh_sym = (a.mode == 1 && err_pos >= 0) ? err_pos : n;
        const float* sf = a.soft + (size_t)c * a.stride;
        float* hist = a.st[c].hist;
        float nh[24];
        for (int k = 0; k < 24; ++k) {
            const int t = nh_sym - 24 + k;
            nh[k] = (t >= 0) ? sf[t] : hist[24 + t];
        }
        for (int k = 0; k < 24; ++k) hist[k] = nh[k];
    }
    int* sm = a.summary + (size_t)c * 8;
    sm[0] = o.n_msgs;
    sm[1] = nid_count;
    sm[2] = o.err_code;
    sm[3] = err_pos;
    sm[4] = o.err_a;
    sm[5] = o.err_b;
    sm[6] = o.err_duid;
    sm[7] = o.pool_used;
}

__global__ void framer_reset_kernel(FramerState* st, int C, int channel, int keep_tracker) {
    const int c = blockIdx.x;
    if (c >= C || (channel >= 0 && c != channel)) return;
    FramerState& f = st[c];
    if (threadIdx.x == 0) {
        FramerCore& s = f.core;
        for (int i = 0; i < 24; ++i) f.hist[i] = 0.f;
        s.sync_detected = 0;
        s.nid_pointer = 0;
        s.dibit_counter = 58;
        s.status_counter = 36;
        s.asm_active = 0;
        s.asm_nac = s.asm_duid = s.asm_nbits = s.asm_target = s.asm_forced = 0;
        s.assembly_required = 0;
        s.previous_duid = DUID_PLACEHOLDER;
        s.detected_duid = DUID_PLACEHOLDER;
        s.detected_nac = 0;
        if (!keep_tracker) {
            s.symbols_total = 0;
            s.detected_errs = 0;
            s.tracked_nac = 0;
        }
    }
    if (!keep_tracker)
        for (int i = threadIdx.x; i < 4096; i += blockDim.x) f.nac_seen[i] = 0;
}


// ---------------------------------------------------------------------------------------------
// 1/2-rate trellis (Viterbi) decoder: dsp/fec/trellis.py:89-289 TrellisDecoder.decode, one thread per block.
// Survivors are register-exchange paths (2 bits per step). Output semantics of the reference: from the 12th step on,
// every step emits the decision 12 steps back of the path that is best AT THAT STEP (first minimum over states);
// the last 11 decisions come from the final best path. Metrics are float64 like the reference's Python floats
// (hard decisions: bit-error counts, exact).
// ---------------------------------------------------------------------------------------------
__constant__ unsigned char c_tr_nibble[16] = {2, 12, 1, 15, 14, 0, 13, 3, 9, 7, 10, 4, 5, 11, 6, 8};  // [state][input]
constexpr int TR_DEPTH = 12;
constexpr int TR_MAXW = 32;   // 64-bit words per survivor: up to 1024 steps (2048 dibits) per block

struct TrellisArgs {
    const unsigned char* dibits;   // [B][stride]
    long long stride;
    const int* n_each;             // [B] or null
    int n_fixed;
    const double* soft;            // [B][stride] or null
    int tsbk;                      // 1: rows are 196 message BITS -> 98 dibits, deinterleaved (decoders/p25.py:2037-2087)
    unsigned char* out;            // tsbk 0: decoded dibits [B][out_stride]; tsbk 1: 96 decoded bits [B][96]
    long long out_stride;
    int* n_out;                    // [B] decoded dibits (tsbk 0)
    int* metric;                   // [B] int(best path metric)
    int* fields;                   // tsbk 1: [B][4] last_block, protected, opcode, mfid
    unsigned char* data8;          // tsbk 1: [B][8] payload bytes
    int B;
};

__device__ __forceinline__ int tsbk_deint(int i) {   // decoders/p25.py:2552-2660 as a rule
    if (i >= 96) return 24 + (i - 96);
    const int grp = i >> 3, r = i & 7;
    const int base = (r < 2) ? 0 : (r < 4) ? 26 : (r < 6) ? 50 : 74;
    return base + 2 * grp + (r & 1);
}

__global__ void __launch_bounds__(64) trellis12_kernel(const TrellisArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const unsigned char* row = a.dibits + (long long)b * a.stride;
    const double* srow = a.soft ? a.soft + (long long)b * a.stride : nullptr;
    int n = a.tsbk ? 98 : (a.n_each ? a.n_each[b] : a.n_fixed);
    n &= ~1;   // odd length: the last dibit is dropped (:228-232)
    const int steps_total = n >> 1;
    const int W = (steps_total + 31) >> 5;
    unsigned long long path[2][4][TR_MAXW];
    double metric[4] = {0.0, INFINITY, INFINITY, INFINITY};
    for (int s = 0; s < 4; ++s)
        for (int w = 0; w < W; ++w) path[0][s][w] = 0ull;
    int cur = 0, n_out = 0;
    unsigned char* out = a.out + (long long)b * a.out_stride;
    unsigned char first48[48];
    auto emit = [&](int d) {
        if (a.tsbk) {
            if (n_out < 48) first48[n_out] = (unsigned char)d;
        } else {
            out[n_out] = (unsigned char)d;
        }
        ++n_out;
    };
    auto get = [&](int i) -> int {
        if (!a.tsbk) return row[i] & 3;   // the reference indexes its tables with the dibit; > 3 would raise there
        const int j = tsbk_deint(i);
        return ((row[2 * j] & 1) << 1) | (row[2 * j + 1] & 1);
    };
    for (int step = 0; step < steps_total; ++step) {
        const int r0 = get(2 * step), r1 = get(2 * step + 1);
        double s0 = 0.0, s1 = 0.0;
        if (srow) {
            s0 = srow[2 * step];
            s1 = srow[2 * step + 1];
        }
        double nm[4];
        int bp[4];
        for (int ns = 0; ns < 4; ++ns) {
            double best = INFINITY;
            int bprev = 0;
            for (int p = 0; p < 4; ++p) {
                const int nib = c_tr_nibble[p * 4 + ns];
                const int e0 = nib >> 2, e1 = nib & 3;
                double br;
                if (srow) {
                    const double l0 = (e0 == 0) ? 1.0 : (e0 == 1) ? 3.0 : (e0 == 2) ? -1.0 : -3.0;
                    const double l1 = (e1 == 0) ? 1.0 : (e1 == 1) ? 3.0 : (e1 == 2) ? -1.0 : -3.0;
                    br = (s0 - l0) * (s0 - l0) + (s1 - l1) * (s1 - l1);
                } else {
                    br = (double)(__popc(r0 ^ e0) + __popc(r1 ^ e1));
                }
                const double m = metric[p] + br;
                if (m < best) {
                    best = m;
                    bprev = p;
                }
            }
            nm[ns] = best;
            bp[ns] = bprev;
        }
        const int nxt = cur ^ 1;
        for (int ns = 0; ns < 4; ++ns) {
            for (int w = 0; w < W; ++w) path[nxt][ns][w] = path[cur][bp[ns]][w];
            const unsigned long long dec = (nm[ns] < INFINITY) ? (unsigned long long)ns : 0ull;
            path[nxt][ns][step >> 5] |= dec << (2 * (step & 31));
            metric[ns] = nm[ns];
        }
        cur = nxt;
        const int steps = step + 1;
        if (steps >= TR_DEPTH) {
            int best = 0;
            for (int st = 1; st < 4; ++st)
                if (metric[st] < metric[best]) best = st;
            const int k = steps - TR_DEPTH;
            emit((int)((path[cur][best][k >> 5] >> (2 * (k & 31))) & 3ull));
        }
    }
    int best = 0;
    for (int st = 1; st < 4; ++st)
        if (metric[st] < metric[best]) best = st;
    for (int k = max(0, steps_total - TR_DEPTH + 1); k < steps_total; ++k)
        emit((int)((path[cur][best][k >> 5] >> (2 * (k & 31))) & 3ull));
    const double mb = metric[best];
    a.metric[b] = (steps_total == 0) ? 0 : (int)mb;
    if (!a.tsbk) {
        a.n_out[b] = n_out;
        return;
    }
    unsigned char* bits = a.out + (long long)b * 96;
    for (int i = 0; i < 48; ++i) {
        bits[2 * i] = (first48[i] >> 1) & 1;
        bits[2 * i + 1] = first48[i] & 1;
    }
    int v = 0;
    for (int i = 2; i < 8; ++i) v = (v << 1) | bits[i];
    int* f = a.fields + (long long)b * 4;
    f[0] = bits[0];
    f[1] = bits[1];
    f[2] = v;
    v = 0;
    for (int i = 8; i < 16; ++i) v = (v << 1) | bits[i];
    f[3] = v;
    for (int k = 0; k < 8; ++k) {
        v = 0;
        for (int i = 0; i < 8; ++i) v = (v << 1) | bits[16 + 8 * k + i];
        a.data8[(long long)b * 8 + k] = (unsigned char)v;
    }
}

// The tables live in __constant__ / __device__ symbols, i.e. once per device: one bit per device ordinal, filled under a
// mutex (framers are created from several threads when a process serves several captures).
static std::atomic<unsigned long long> g_tables_ready{0};
static std::mutex g_tables_mu;
static int ensure_tables() {
    int dev = 0;
    WC_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (g_tables_ready.load(std::memory_order_acquire) & bit) return 0;
    std::lock_guard<std::mutex> lk(g_tables_mu);
    if (g_tables_ready.load(std::memory_order_acquire) & bit) return 0;
    unsigned char pw[128], lg[64];
    memset(pw, 0, sizeof(pw));
    memset(lg, 0, sizeof(lg));
    int x = 1;
    for (int i = 0; i < 63; ++i) {
        pw[i] = (unsigned char)x;
        lg[x] = (unsigned char)i;
        x <<= 1;
        if (x & 64) x ^= 0x43;
    }
    for (int i = 63; i < 128; ++i) pw[i] = pw[i - 63];
    float sync[24];
    const unsigned long long pat = 0x5575F5FF77FFull;
    for (int i = 0; i < 24; ++i) sync[i] = (((pat >> ((23 - i) * 2)) & 3ull) == 1ull) ? 3.0f : -3.0f;
    // packed odd syndromes of every single bit, then of every byte value at every byte position
    std::vector<uint4> tab(8 * 256);
    for (int b = 0; b < 8; ++b)
        for (int v = 0; v < 256; ++v) {
            unsigned long long lo = 0ull;
            unsigned int hi = 0u;
            for (int k = 0; k < 8; ++k) {
                if (!((v >> k) & 1)) continue;
                const int p = 8 * b + k;  // bit 63 never occurs in a 63-bit word; its row stays consistent anyway
                for (int j = 0; j < 11; ++j) {
                    const unsigned long long e = pw[((2 * j + 1) * p) % 63];
                    if (j < 10) lo ^= e << (6 * j);
                    else hi ^= (unsigned int)e;
                }
            }
            tab[b * 256 + v] = make_uint4((unsigned int)(lo & 0xffffffffull), (unsigned int)(lo >> 32), hi, 0u);
        }
    WC_CUDA(cudaMemcpyToSymbol(c_gf_pw, pw, sizeof(pw)));
    WC_CUDA(cudaMemcpyToSymbol(c_gf_lg, lg, sizeof(lg)));
    WC_CUDA(cudaMemcpyToSymbol(c_fsync, sync, sizeof(sync)));
    WC_CUDA(cudaMemcpyToSymbol(g_syn, tab.data(), sizeof(uint4) * tab.size()));
    g_tables_ready.fetch_or(bit, std::memory_order_release);
    return 0;
}

}  // namespace wc

using namespace wc;

struct wc_p25framer {
    int C = 0;
    FramerState* d_state = nullptr;
    float* d_scores = nullptr;   size_t scores_cap = 0;   // [C][n] (host-call staging of the optional score output)
    unsigned char* d_hits = nullptr; size_t hits_cap = 0;  // [C][n] score > 60
    // host-call staging
    void* d_soft = nullptr;      size_t soft_cap = 0;
    void* d_dibits = nullptr;    size_t dib_cap = 0;
    void* d_hdr = nullptr;       size_t hdr_cap = 0;
    void* d_hsym = nullptr;      size_t hsym_cap = 0;
    void* d_pool = nullptr;      size_t pool_cap_b = 0;
    int* d_summary = nullptr;
    int* d_nsym = nullptr;
    cudaStream_t stream = nullptr;
};

static int grow(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    WC_CUDA(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

extern "C" {

int wc_bch_decode(const unsigned char* bits63_dev, const int* tracked_nac_dev, int count, int* data_dev, int* errors_dev,
                  void* stream_v) {
    WC_REQUIRE(bits63_dev && data_dev && errors_dev, "wc_bch_decode: null argument");
    if (count <= 0) return 0;
    if (ensure_tables()) return -2;
    bch_batch_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(bits63_dev, tracked_nac_dev, count, data_dev,
                                                                             errors_dev);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_bch_decode_host(const unsigned char* bits63_host, const int* tracked_nac_host, int count, int* data_host,
                       int* errors_host) {
    WC_REQUIRE(bits63_host && data_host && errors_host, "wc_bch_decode_host: null argument");
    if (count <= 0) return 0;
    unsigned char* d_bits = nullptr;
    int *d_tr = nullptr, *d_data = nullptr, *d_err = nullptr;
    WC_CUDA(cudaMalloc(&d_bits, (size_t)count * 63));
    WC_CUDA(cudaMalloc(&d_data, sizeof(int) * (size_t)count));
    WC_CUDA(cudaMalloc(&d_err, sizeof(int) * (size_t)count));
    if (tracked_nac_host) WC_CUDA(cudaMalloc(&d_tr, sizeof(int) * (size_t)count));
    int rc = 0;
    cudaMemcpy(d_bits, bits63_host, (size_t)count * 63, cudaMemcpyHostToDevice);
    if (d_tr) cudaMemcpy(d_tr, tracked_nac_host, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice);
    rc = wc_bch_decode(d_bits, d_tr, count, d_data, d_err, nullptr);
    if (!rc) {
        cudaMemcpy(data_host, d_data, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost);
        if (cudaMemcpy(errors_host, d_err, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("wc_bch_decode_host: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -2;
        }
    }
    cudaFree(d_bits);
    cudaFree(d_data);
    cudaFree(d_err);
    if (d_tr) cudaFree(d_tr);
    return rc;
}

int wc_p25framer_create(int n_channels, wc_p25framer** out) {
    WC_REQUIRE(out != nullptr, "wc_p25framer_create: out is null");
    WC_REQUIRE(n_channels >= 1 && n_channels <= 65536, "wc_p25framer_create: n_channels %d out of range", n_channels);
    if (ensure_tables()) return -2;
    wc_p25framer* h = new wc_p25framer();
    h->C = n_channels;
    if (cudaMalloc(&h->d_state, sizeof(FramerState) * (size_t)n_channels) != cudaSuccess ||
        cudaMalloc(&h->d_summary, sizeof(int) * 8 * (size_t)n_channels) != cudaSuccess ||
        cudaMalloc(&h->d_nsym, sizeof(int) * (size_t)n_channels) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("wc_p25framer_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return -2;
    }
    cudaMemset(h->d_state, 0, sizeof(FramerState) * (size_t)n_channels);
    framer_reset_kernel<<<n_channels, 128, 0, h->stream>>>(h->d_state, n_channels, -1, 0);
    cudaStreamSynchronize(h->stream);
    *out = h;
    return 0;
}

void wc_p25framer_destroy(wc_p25framer* h) {
    if (!h) return;
    cudaFree(h->d_state);
    cudaFree(h->d_summary);
    cudaFree(h->d_nsym);
    if (h->d_scores) cudaFree(h->d_scores);
    if (h->d_hits) cudaFree(h->d_hits);
    if (h->d_soft) cudaFree(h->d_soft);
    if (h->d_dibits) cudaFree(h->d_dibits);
    if (h->d_hdr) cudaFree(h->d_hdr);
    if (h->d_hsym) cudaFree(h->d_hsym);
    if (h->d_pool) cudaFree(h->d_pool);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

/* P25P1MessageFramer.reset (p25_framer.py:837-849) keeps the NAC tracker and the symbol clock; full = 1 also clears those */
int wc_p25framer_reset(wc_p25framer* h, int channel, int full) {
    WC_REQUIRE(h != nullptr, "wc_p25framer_reset: null handle");
    WC_REQUIRE(channel >= -1 && channel < h->C, "wc_p25framer_reset: channel %d out of range", channel);
    framer_reset_kernel<<<h->C, 128, 0, h->stream>>>(h->d_state, h->C, channel, full ? 0 : 1);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int wc_p25framer_max_msgs(int n_symbols) { return n_symbols / 57 + 8; }
int wc_p25framer_pool_bytes(int n_symbols) { return 2 * n_symbols + 2 * MAX_MSG_BITS + 4 * wc_p25framer_max_msgs(n_symbols); }

int wc_p25framer_process(wc_p25framer* h, const float* soft_dev, const unsigned char* dibits_dev, long long chan_stride,
                         const int* n_sym_dev, int n_symbols, int mode, int dispatch_enabled, float* scores_dev,
                         int* msg_hdr_dev, long long* msg_sym_dev, unsigned char* msg_bits_dev, int* summary_dev,
                         void* stream_v) {
    WC_REQUIRE(h && soft_dev && dibits_dev && msg_hdr_dev && msg_sym_dev && msg_bits_dev && summary_dev,
               "wc_p25framer_process: null argument");
    WC_REQUIRE(n_symbols >= 0 && chan_stride >= n_symbols, "wc_p25framer_process: bad n_symbols / chan_stride");
    WC_REQUIRE(mode >= 0 && mode <= 2, "wc_p25framer_process: bad mode %d", mode);
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int nmax = n_symbols > 0 ? n_symbols : 1;
    if (grow((void**)&h->d_hits, &h->hits_cap, (size_t)h->C * nmax)) return -2;
    if (n_symbols > 0 && mode != 2) {
        dim3 grid((n_symbols + 127) / 128, h->C);
        framer_score_kernel<<<grid, 128, 0, stream>>>(soft_dev, chan_stride, n_sym_dev, n_symbols, h->d_state, h->d_hits,
                                                      scores_dev, nmax);
    }
    FramerArgs a;
    a.st = h->d_state;
    a.soft = soft_dev;
    a.dibits = dibits_dev;
    a.stride = chan_stride;
    a.n_sym = n_sym_dev;
    a.n_fixed = n_symbols;
    a.hits = h->d_hits;
    a.hit_stride = nmax;
    a.mode = mode;
    a.dispatch_enabled = dispatch_enabled;
    a.C = h->C;
    a.hdr = msg_hdr_dev;
    a.hdr_sym = msg_sym_dev;
    a.pool = msg_bits_dev;
    a.max_msgs = wc_p25framer_max_msgs(n_symbols);
    a.pool_cap = wc_p25framer_pool_bytes(n_symbols);
    a.summary = summary_dev;
    framer_machine_kernel<<<(h->C + 31) / 32, 32, 0, stream>>>(a);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_p25framer_process_host(wc_p25framer* h, const float* soft_host, const unsigned char* dibits_host, int n_symbols,
                              const int* n_sym_host, int mode, int dispatch_enabled, float* scores_host, int* msg_hdr_host,
                              long long* msg_sym_host, unsigned char* msg_bits_host, int* summary_host) {
    WC_REQUIRE(h && soft_host && dibits_host && msg_hdr_host && msg_sym_host && msg_bits_host && summary_host,
               "wc_p25framer_process_host: null argument");
    WC_REQUIRE(n_symbols >= 0, "wc_p25framer_process_host: negative n_symbols");
    const int C = h->C;
    const int nmax = n_symbols > 0 ? n_symbols : 1;
    const int mm = wc_p25framer_max_msgs(n_symbols), pc = wc_p25framer_pool_bytes(n_symbols);
    if (grow(&h->d_soft, &h->soft_cap, sizeof(float) * (size_t)C * nmax)) return -2;
    if (grow(&h->d_dibits, &h->dib_cap, (size_t)C * nmax)) return -2;
    if (grow(&h->d_hdr, &h->hdr_cap, sizeof(int) * 6 * (size_t)C * mm)) return -2;
    if (grow(&h->d_hsym, &h->hsym_cap, sizeof(long long) * (size_t)C * mm)) return -2;
    if (grow(&h->d_pool, &h->pool_cap_b, (size_t)C * pc)) return -2;
    if (grow((void**)&h->d_scores, &h->scores_cap, sizeof(float) * (size_t)C * nmax)) return -2;
    cudaStream_t s = h->stream;
    if (n_symbols > 0) {
        WC_CUDA(cudaMemcpyAsync(h->d_soft, soft_host, sizeof(float) * (size_t)C * n_symbols, cudaMemcpyHostToDevice, s));
        WC_CUDA(cudaMemcpyAsync(h->d_dibits, dibits_host, (size_t)C * n_symbols, cudaMemcpyHostToDevice, s));
    }
    if (n_sym_host) WC_CUDA(cudaMemcpyAsync(h->d_nsym, n_sym_host, sizeof(int) * (size_t)C, cudaMemcpyHostToDevice, s));
    int rc = wc_p25framer_process(h, (const float*)h->d_soft, (const unsigned char*)h->d_dibits, nmax,
                                  n_sym_host ? h->d_nsym : nullptr, n_symbols, mode, dispatch_enabled,
                                  scores_host ? h->d_scores : nullptr,
                                  (int*)h->d_hdr, (long long*)h->d_hsym, (unsigned char*)h->d_pool, h->d_summary, s);
    if (rc) return rc;
    WC_CUDA(cudaMemcpyAsync(summary_host, h->d_summary, sizeof(int) * 8 * (size_t)C, cudaMemcpyDeviceToHost, s));
    WC_CUDA(cudaMemcpyAsync(msg_hdr_host, h->d_hdr, sizeof(int) * 6 * (size_t)C * mm, cudaMemcpyDeviceToHost, s));
    WC_CUDA(cudaMemcpyAsync(msg_sym_host, h->d_hsym, sizeof(long long) * (size_t)C * mm, cudaMemcpyDeviceToHost, s));
    WC_CUDA(cudaMemcpyAsync(msg_bits_host, h->d_pool, (size_t)C * pc, cudaMemcpyDeviceToHost, s));
    if (scores_host && n_symbols > 0)
        WC_CUDA(cudaMemcpyAsync(scores_host, h->d_scores, sizeof(float) * (size_t)C * n_symbols, cudaMemcpyDeviceToHost, s));
    WC_CUDA(cudaStreamSynchronize(s));
    return 0;
}

/* state12 = {sync_detected, nid_pointer, dibit_counter, status_counter, asm_active, asm_duid, asm_nbits, detected_nac,
 * detected_duid, tracked_nac, previous_duid, symbols_total (low 31 bits)} */
int wc_p25framer_get_state(wc_p25framer* h, int channel, int* state12) {
    WC_REQUIRE(h && state12, "wc_p25framer_get_state: null argument");
    WC_REQUIRE(channel >= 0 && channel < h->C, "wc_p25framer_get_state: channel %d out of range", channel);
    static FramerState tmp;  // ~6 KB: keep it off the stack
    WC_CUDA(cudaStreamSynchronize(h->stream));
    WC_CUDA(cudaMemcpy(&tmp, h->d_state + channel, sizeof(FramerState), cudaMemcpyDeviceToHost));
    const FramerCore& k = tmp.core;
    state12[0] = k.sync_detected;
    state12[1] = k.nid_pointer;
    state12[2] = k.dibit_counter;
    state12[3] = k.status_counter;
    state12[4] = k.asm_active;
    state12[5] = k.asm_duid;
    state12[6] = k.asm_nbits;
    state12[7] = k.detected_nac;
    state12[8] = k.detected_duid;
    state12[9] = k.tracked_nac;
    state12[10] = k.previous_duid;
    state12[11] = (int)(k.symbols_total & 0x7fffffff);
    return 0;
}

/* dsp/fec/trellis.py:214-272 TrellisDecoder.decode for `count` blocks: dibits uint8 [count][stride] (n_dibits[b] valid, or
 * n_fixed when n_dibits is NULL; an odd length drops its last dibit), soft float64 [count][stride] or NULL (hard decisions)
 * -> decoded dibits [count][out_stride] (n_out[b] = n/2 of them), metric[b] = int(best path metric). */
int wc_trellis12_decode(const unsigned char* dibits_dev, long long stride, const int* n_dibits_dev, int n_fixed,
                        const double* soft_dev, int count, unsigned char* out_dev, long long out_stride, int* n_out_dev,
                        int* metric_dev, void* stream_v) {
    WC_REQUIRE(dibits_dev && out_dev && n_out_dev && metric_dev, "wc_trellis12_decode: null argument");
    WC_REQUIRE(n_fixed >= 0 && n_fixed <= 2 * 32 * TR_MAXW && stride >= n_fixed && out_stride >= n_fixed / 2,
               "wc_trellis12_decode: block length %d outside [0, %d]", n_fixed, 2 * 32 * TR_MAXW);
    if (count <= 0) return 0;
    TrellisArgs a;
    memset(&a, 0, sizeof(a));
    a.dibits = dibits_dev;
    a.stride = stride;
    a.n_each = n_dibits_dev;
    a.n_fixed = n_fixed;
    a.soft = soft_dev;
    a.tsbk = 0;
    a.out = out_dev;
    a.out_stride = out_stride;
    a.n_out = n_out_dev;
    a.metric = metric_dev;
    a.B = count;
    trellis12_kernel<<<(count + 63) / 64, 64, 0, (cudaStream_t)stream_v>>>(a);
    WC_CUDA(cudaGetLastError());
    return 0;
}

/* decoders/p25.py:2037-2109 for `count` TSBK blocks: 196 message bits (as P25P1Message.bits) -> 98 dibits -> deinterleave
 * (:2552-2660) -> trellis decode -> 96 bits; fields [count][4] = last_block, protected, opcode, mfid; data8 = payload. */
int wc_tsbk_decode(const unsigned char* bits196_dev, int count, unsigned char* bits96_dev, int* metric_dev, int* fields_dev,
                   unsigned char* data8_dev, void* stream_v) {
    WC_REQUIRE(bits196_dev && bits96_dev && metric_dev && fields_dev && data8_dev, "wc_tsbk_decode: null argument");
    if (count <= 0) return 0;
    TrellisArgs a;
    memset(&a, 0, sizeof(a));
    a.dibits = bits196_dev;
    a.stride = 196;
    a.n_fixed = 98;
    a.tsbk = 1;
    a.out = bits96_dev;
    a.out_stride = 96;
    a.metric = metric_dev;
    a.fields = fields_dev;
    a.data8 = data8_dev;
    a.B = count;
    trellis12_kernel<<<(count + 63) / 64, 64, 0, (cudaStream_t)stream_v>>>(a);
    WC_CUDA(cudaGetLastError());
    return 0;
}

int wc_tsbk_decode_host(const unsigned char* bits196_host, int count, unsigned char* bits96_host, int* metric_host,
                        int* fields_host, unsigned char* data8_host) {
    WC_REQUIRE(bits196_host && bits96_host && metric_host && fields_host && data8_host, "wc_tsbk_decode_host: null argument");
    if (count <= 0) return 0;
    unsigned char *d_in = nullptr, *d_out = nullptr, *d_d8 = nullptr;
    int *d_m = nullptr, *d_f = nullptr;
    WC_CUDA(cudaMalloc(&d_in, (size_t)count * 196));
    WC_CUDA(cudaMalloc(&d_out, (size_t)count * 96));
    WC_CUDA(cudaMalloc(&d_d8, (size_t)count * 8));
    WC_CUDA(cudaMalloc(&d_m, sizeof(int) * (size_t)count));
    WC_CUDA(cudaMalloc(&d_f, sizeof(int) * 4 * (size_t)count));
    cudaMemcpy(d_in, bits196_host, (size_t)count * 196, cudaMemcpyHostToDevice);
    int rc = wc_tsbk_decode(d_in, count, d_out, d_m, d_f, d_d8, nullptr);
    if (!rc) {
        cudaMemcpy(bits96_host, d_out, (size_t)count * 96, cudaMemcpyDeviceToHost);
        cudaMemcpy(metric_host, d_m, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost);
        cudaMemcpy(fields_host, d_f, sizeof(int) * 4 * (size_t)count, cudaMemcpyDeviceToHost);
        if (cudaMemcpy(data8_host, d_d8, (size_t)count * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("wc_tsbk_decode_host: %s", cudaGetErrorString(cudaGetLastError()));
            rc = -2;
        }
    }
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_d8);
    cudaFree(d_m);
    cudaFree(d_f);
    return rc;
}

}  // extern "C"
