"""Transmit side of the DVB-S outer code (EN 300 421 4.4.1-4.4.3), test infrastructure: energy dispersal over groups of
eight TS packets, RS(204,188), convolutional interleaver I = 12, M = 17 -- the stream DVBSDemod::process hands to its
frame loop (dvbs/module_dvbs_demod.cpp:91-106) in frames of 8 x 204 bytes."""
import numpy as np

import orclib


def prbs_bytes(n):
    """n bytes of the dispersal sequence 1 + x^14 + x^15 from the load value (dvbs_scrambling.h:16-28, reg = 0xa9)"""
    reg, out = 0xA9, np.zeros(n, np.uint8)
    for i in range(n):
        res = 0
        for _ in range(8):
            fb = ((reg >> 13) ^ (reg >> 14)) & 1
            reg = ((reg << 1) | fb) & 0x7FFF
            res = (res << 1) | fb
        out[i] = res
    return out


_PRBS = None


def scramble(packets):
    """packets [n][188] with 0x47 sync, n a multiple of 8 -> dispersed packets, first sync of every group inverted"""
    global _PRBS
    if _PRBS is None:
        _PRBS = prbs_bytes(8 * 188)
    p = np.array(packets, np.uint8).reshape(-1, 8 * 188).copy()
    seq = _PRBS.copy()
    # after the load the generator first serves byte 1 of packet 0; the sync bytes of packets 1..7 consume 8 clocks each
    mask = np.zeros(8 * 188, np.uint8)
    mask[1:] = seq[:8 * 188 - 1]
    mask[188::188] = 0
    p ^= mask
    p[:, 0] = 0xB8
    return p.reshape(-1, 188)


def rs_encode(packets):
    o = orclib.oracle()
    out = np.zeros((len(packets), 204), np.uint8)
    out[:, :188] = packets
    par = np.zeros(16, np.uint8)
    for i, pk in enumerate(np.ascontiguousarray(packets, np.uint8)):
        o.orc_rs204_parity(pk, par)
        out[i, 188:] = par
    return out


def interleave(stream):
    """byte n goes through a FIFO of 17 (n % 12) cells (dvbs_interleaving.h:23-27,43-55), zeros to begin with"""
    s = np.asarray(stream, np.uint8)
    out = np.zeros_like(s)
    n = np.arange(len(s))
    src = n - 17 * 12 * (n % 12)
    ok = src >= 0
    out[ok] = s[src[ok]]
    return out


def outer_stream(ngroups, rng, payload=None, codeword_errors=None):
    """-> (ts packets [8 ngroups][188], interleaved channel bytes [8 ngroups * 204]); codeword_errors = (lo, hi): that many
    byte errors in every RS codeword BEFORE interleaving (<= 8 is always correctable, wherever the channel puts them)"""
    n = 8 * ngroups
    ts = rng.integers(0, 256, (n, 188), dtype=np.uint8) if payload is None else np.array(payload, np.uint8)
    ts[:, 0] = 0x47
    coded = rs_encode(scramble(ts))
    if codeword_errors is not None:
        coded = add_errors(coded.reshape(-1), rng, per_packet=codeword_errors).reshape(-1, 204)
    return ts, interleave(coded.reshape(-1))


def add_errors(stream, rng, per_packet=(0, 9), burst_every=0):
    """byte errors in the channel stream: a random number per 204 bytes"""
    s = np.array(stream, np.uint8)
    for p in range(len(s) // 204):
        k = int(rng.integers(per_packet[0], per_packet[1] + 1))
        pos = rng.choice(204, k, replace=False) + 204 * p
        s[pos] ^= rng.integers(1, 256, k, dtype=np.uint8)
    return s


# ---- inner code (EN 300 421 4.4.3): K = 7 convolutional code, punctured, as soft bits -----------------------------------
RATES = ["1/2", "2/3", "3/4", "5/6", "7/8"]
# which of X (first polynomial) and Y (second) survive, per information bit of the puncturing period (EN 300 421 table 2)
_PUNCT = {0: ([1], [1]), 1: ([1, 0], [1, 1]), 2: ([1, 0, 1], [1, 1, 0]), 3: ([1, 0, 1, 0, 1], [1, 1, 0, 1, 0]),
          4: ([1, 0, 0, 0, 1, 0, 1], [1, 1, 1, 1, 0, 1, 0])}


def _parity(v):
    v = v.astype(np.uint8)
    v ^= v >> 4
    v ^= v >> 2
    v ^= v >> 1
    return v & 1


def conv_encode(bits, state=0):
    """the reference's CCEncoder (cc_encoder.cpp:92-104): register << 1 | bit, outputs parity(reg & 79), parity(reg & 109)"""
    b = np.asarray(bits, np.uint8) & 1
    hist = np.concatenate([[(state >> k) & 1 for k in range(5, -1, -1)], b]).astype(np.int64)   # oldest first
    reg = np.zeros(len(b), np.int64)
    for k in range(7):
        reg |= hist[6 - k:len(hist) - k] << k
    return _parity(reg & 79), _parity(reg & 109)


def puncture(x, y, rate):
    """-> the transmitted bit sequence of one rate (index into RATES)"""
    px, py = _PUNCT[rate]
    n = len(x)
    keep = np.zeros((n, 2), bool)
    keep[:, 0] = np.resize(np.array(px, bool), n)
    keep[:, 1] = np.resize(np.array(py, bool), n)
    return np.stack([x, y], 1)[keep]


def inner_softs(bits, rate, rng, amp=60.0, sigma=0.0, phase=0, lead=0):
    """decoded-domain bits -> the signed soft bits DVBSymToSoftBlock hands to the Viterbi decoder: bit 1 = +amp; `lead`
    soft bits of noise first (moves the puncturing / pairing phase); phase 1: the constellation turned so that the decoder
    has to lock on its 90 degree branch (it maps (a, b) -> (b, -a), rotation.cpp:33-42)"""
    x, y = conv_encode(bits)
    tx = puncture(x, y, rate).astype(np.float64) * 2 - 1
    s = np.concatenate([rng.normal(0, 20, lead), tx * amp])
    if len(s) % 2:
        s = s[:-1]
    if sigma:
        s = s + rng.normal(0, sigma, len(s))
    s = np.clip(np.rint(s), -127, 127)
    if phase == 1:
        a, b = s[0::2].copy(), s[1::2].copy()
        s[0::2], s[1::2] = -b, a
    return s.astype(np.int8)
