"""CPU-side checks: the C ABI library loads and exports what include/dvbs2fec.h declares, the in-tree
transmitter agrees with the oracle, the expanded code tables satisfy the standard's structure, and the
product refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orclib
from fec import pkg, MODCODS, QPSK_MODCOD_OF_RATE
from orclib import ALL_CODES, code_params


def test_library_exports_every_declared_symbol():
    hdr = open(pkg.INCLUDE).read()
    names = set(re.findall(r"\b(dvbs2fec_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 25
    L = pkg.lib()
    for n in sorted(names):
        assert hasattr(L, n), n


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.DVBS2FecError) as e:
        pkg.DVBS2Decoder()
    assert e.value.code == pkg.ENODEV


def test_oracle_is_not_reachable_from_the_product():
    root = os.path.dirname(pkg.LIB_PATH)
    for dirpath, _, files in os.walk(root):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in text and "orclib" not in text and "libdvbs2_oracle" not in text, f


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_modcod_info_matches_code_tables(short, rate):
    p = code_params(short, rate)
    info = pkg.modcod_info(QPSK_MODCOD_OF_RATE[rate], bool(short))
    assert (info["nldpc"], info["kldpc"], info["kbch"], info["bch_t"], info["links_total"]) == (
        p["N"], p["K"], p["kbch"], p["t"], p["links"])
    assert info["plframe_symbols"] == 90 + p["N"] // 2


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_transmitter_matches_oracle_and_parity_checks(short, rate):
    """product encoder == oracle encoder on the same payload; H c^T = 0 through the oracle's schedule"""
    o = orclib.oracle()
    p = code_params(short, rate)
    rng = np.random.default_rng(rate + 31 * short)
    payload = rng.integers(0, 256, p["kbch"] // 8, dtype=np.uint8)
    code = pkg.encode_fecframe(QPSK_MODCOD_OF_RATE[rate], bool(short), payload)
    # oracle path: scramble (self-inverse) -> BCH -> LDPC
    frame = np.zeros(p["K"] // 8, np.uint8)
    frame[: p["kbch"] // 8] = payload
    o.orc_descramble(short, rate, frame)
    assert o.orc_bch_encode(short, rate, frame) == 0
    want = np.zeros(p["N"], np.uint8)
    o.orc_ldpc_encode_bits(short, rate, np.unpackbits(frame), want)
    assert np.array_equal(code, want)
    # a noiseless codeword needs 0 iterations and 0 corrections, and gives the payload back
    llr = np.where(code > 0, -20, 20).astype(np.int8)
    bb = np.zeros(p["kbch"] // 8, np.uint8)
    it, co = C.c_int(), C.c_int()
    o.orc_decode_frame(short, rate, llr, 25, bb, C.byref(it), C.byref(co))
    assert (it.value, co.value) == (0, 0)
    assert np.array_equal(bb, payload)


@pytest.mark.parametrize("modcod,short", [(4, 0), (12, 0), (13, 1), (18, 0), (22, 1), (24, 0), (27, 1)])
def test_modulator_is_inverse_of_reference_demapper(modcod, short):
    """noiseless symbols from the in-tree mapper, demapped by the oracle: hard decisions == code bits"""
    const, ctype, rate, g1, g2 = MODCODS[modcod]
    o = orclib.oracle()
    c = o.orc_const_create(ctype, g1, g2)
    n = 16200 if short else 64800
    rng = np.random.default_rng(modcod)
    bits = rng.integers(0, 2, n, dtype=np.uint8)
    pl = pkg.modulate(modcod, bool(short), False, bits).view(np.float32)
    llr = np.zeros(n, np.int8)
    o.orc_bb_to_soft(c, const, short, rate, np.ascontiguousarray(pl), llr)
    assert np.array_equal((llr < 0).astype(np.uint8), bits)
    assert (llr != 0).all()
    # pilots: same payload symbols, 36 zeros after every 1440
    plp = pkg.modulate(modcod, bool(short), True, bits)
    nsym = n // (const + 2)
    assert plp.size == 90 + nsym + 36 * ((nsym - 1) // 1440)
    data = np.concatenate([plp[90 + 1476 * k: 90 + 1476 * k + 1440] for k in range((nsym + 1439) // 1440)])[:nsym]
    assert np.array_equal(data, pl.view(np.complex64)[90:90 + nsym])
    o.orc_const_destroy(c)


def test_bb_scrambler_sequence_start():
    """EN 302 307 5.2.2: register 100101010000000, output = XOR of the last two stages; first byte 0x03..."""
    o = orclib.oracle()
    z = np.zeros(32400 // 8, np.uint8)
    o.orc_descramble(0, 3, z)
    sr = [1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0]
    bits = []
    for _ in range(64):
        b = sr[13] ^ sr[14]
        bits.append(b)
        sr = [b] + sr[:-1]
    assert np.array_equal(np.unpackbits(z[:8]), np.array(bits, np.uint8))


def test_bbheader_crc8_restatement_agrees_with_standard_encoder():
    """EN 302 307 5.1.6: CRC-8 with g(x) = x^8+x^7+x^6+x^4+x^2+1 over the first 72 BBHEADER bits, appended MSB
    first; the reference parser's check (bbframe_ts_parser.cpp:66-80), as restated in the oracle, must then be 0."""
    o = orclib.oracle()
    rng = np.random.default_rng(1)
    for _ in range(50):
        hdr = rng.integers(0, 256, 10, dtype=np.uint8)
        reg = 0
        for n in range(72):
            bit = (int(hdr[n >> 3]) >> (7 - (n & 7))) & 1
            fb = ((reg >> 7) & 1) ^ bit
            reg = ((reg << 1) & 0xFF) ^ (0xD5 if fb else 0)
        hdr[9] = reg
        assert o.orc_bbheader_crc8(hdr) == 0
        hdr[3] ^= 0x10
        assert o.orc_bbheader_crc8(hdr) != 0
