"""Pins the C restatement (oracle/) against the reference's own sources compiled unmodified
(oracle/_ref).  Runs wherever oracle/_ref/libdvbs2_ref.so exists (the build container; the GPU box
gets the prebuilt file).  CPU only."""
import ctypes as C

import numpy as np
import pytest

import orclib
from orclib import ALL_CODES, code_params

pytestmark = pytest.mark.skipif(not orclib.have_ref(), reason="compiled reference not built")


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_schedule_matches_reference_iterator(short, rate):
    """pos[] built by the oracle == the (bit, check) edges the reference's LDPC<TABLE> iterator yields,
    sorted per check by ascending bit, permuted to layered order (layered_decoder.hh:95-119)."""
    o, r = orclib.oracle(), orclib.ref()
    p = code_params(short, rate)
    R = p["N"] - p["K"]
    q = p["q"]
    nlinks = p["links"] - 2 * R + 1
    bits = np.zeros(nlinks, np.int32)
    chks = np.zeros(nlinks, np.int32)
    assert r.ref_ldpc_links(short, rate, bits, chks, nlinks) == nlinks
    cnl = o.orc_ldpc_schedule(short, rate, None, None)
    pos = np.zeros(R * cnl, np.uint16)
    cnc = np.zeros(q, np.uint8)
    o.orc_ldpc_schedule(short, rate, pos.ctypes.data, cnc.ctypes.data)
    pos = pos.reshape(R, cnl)
    # reference edges grouped by check, in iteration (= ascending bit) order
    order = np.argsort(chks, kind="stable")
    counts = np.bincount(chks, minlength=R)
    assert counts.max() == cnl
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    sb = bits[order]
    for i in range(q):
        for j in (0, 1, 179, 359):
            chk = q * j + i
            n = counts[chk]
            assert n == cnc[i]
            assert np.array_equal(pos[360 * i + j, :n], sb[starts[chk]: starts[chk] + n])


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_ldpc_encoder_matches_reference(short, rate):
    o, r = orclib.oracle(), orclib.ref()
    p = code_params(short, rate)
    rng = np.random.default_rng(100 + rate + 50 * short)
    data = rng.integers(0, 2, p["K"], dtype=np.uint8)
    a = np.zeros(p["N"], np.uint8)
    b = np.zeros(p["N"], np.uint8)
    assert o.orc_ldpc_encode_bits(short, rate, data, a) == 0
    assert r.ref_ldpc_encode_bits(short, rate, data, b) == 0
    assert np.array_equal(a, b)


def _ldpc_case(short, rate, esn0, seed, max_trials, kind="awgn"):
    p = code_params(short, rate)
    rng = np.random.default_rng(seed)
    if kind == "awgn":
        _, code = orclib.encode_frame(short, rate, rng)
        llr = orclib.awgn_llr(code, esn0, rng)
    elif kind == "random":
        llr = rng.integers(-128, 128, p["N"], dtype=np.int8)
    elif kind == "saturated":
        _, code = orclib.encode_frame(short, rate, rng)
        llr = np.where(code > 0, -128, 127).astype(np.int8)
        flips = rng.integers(0, p["N"], 40)
        llr[flips] = -llr[flips].astype(np.int16).clip(-127, 127).astype(np.int8)
    elif kind == "zeros":
        _, code = orclib.encode_frame(short, rate, rng)
        llr = orclib.awgn_llr(code, esn0, rng)
        llr[rng.integers(0, p["N"], 200)] = 0
    a = llr.copy()
    b = llr.copy()
    ra = orclib.oracle().orc_ldpc_decode(short, rate, a, max_trials)
    rb = orclib.ref().ref_ldpc_decode(short, rate, b, max_trials)
    return ra, rb, a, b


# Es/N0 chosen per rate so that each code sees converging and non-converging frames
_SNR = {0: -2.0, 1: -1.0, 2: 0.0, 3: 1.3, 4: 2.6, 5: 3.4, 6: 4.3, 7: 5.0, 8: 5.5, 10: 6.5, 11: 6.7}


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_ldpc_decode_matches_reference(short, rate):
    seen = set()
    for k, (dsnr, kind) in enumerate([(0.3, "awgn"), (-0.6, "awgn"), (3.0, "awgn"), (0.3, "zeros"),
                                      (0, "random"), (0, "saturated")]):
        ra, rb, a, b = _ldpc_case(short, rate, _SNR[rate] + dsnr + (0.4 if short else 0), 1000 * rate + k + 7 * short,
                                  25 if k != 2 else 16, kind)
        assert ra == rb, (kind, ra, rb)
        assert np.array_equal(a, b), kind
        seen.add(ra)
    assert len(seen) >= 2  # both converged and non-converged / different iteration counts were exercised


@pytest.mark.parametrize("short,rate", [(0, 3), (0, 5), (0, 10), (1, 3), (1, 10)])
def test_bch_matches_reference(short, rate):
    o, r = orclib.oracle(), orclib.ref()
    p = code_params(short, rate)
    rng = np.random.default_rng(5 + rate)
    K, kbch, t = p["K"], p["kbch"], p["t"]
    for nerr in [0, 1, 2, 2, 3, t - 1, t, t + 1, t + 3, 40]:
        frame = np.zeros(K // 8, np.uint8)
        frame[: kbch // 8] = rng.integers(0, 256, kbch // 8, dtype=np.uint8)
        enc_a = frame.copy()
        enc_b = frame.copy()
        assert o.orc_bch_encode(short, rate, enc_a) == 0
        r.ref_bch_encode(short, rate, enc_b)
        assert np.array_equal(enc_a, enc_b)
        errs = rng.choice(K, nerr, replace=False)
        for e in errs:
            enc_a[e >> 3] ^= 0x80 >> (e & 7)
        enc_b = enc_a.copy()
        ca = o.orc_bch_decode(short, rate, enc_a)
        cb = r.ref_bch_decode(short, rate, enc_b)
        assert ca == cb, (nerr, ca, cb)
        assert np.array_equal(enc_a, enc_b)
        if nerr <= t:
            assert ca == nerr


def test_bch_adjacent_double_errors_hit_artin_schreier_path():
    """degree-2 locators go through the half-trace table incl. its skipped entry (reed_solomon_error_correction.hh:80-83)"""
    o, r = orclib.oracle(), orclib.ref()
    short, rate = 1, 3
    p = code_params(short, rate)
    rng = np.random.default_rng(9)
    K, kbch = p["K"], p["kbch"]
    base = np.zeros(K // 8, np.uint8)
    base[: kbch // 8] = rng.integers(0, 256, kbch // 8, dtype=np.uint8)
    o.orc_bch_encode(short, rate, base)
    mism = 0
    for trial in range(300):
        e = rng.choice(K, 2, replace=False)
        a = base.copy()
        for x in e:
            a[x >> 3] ^= 0x80 >> (x & 7)
        b = a.copy()
        ca, cb = o.orc_bch_decode(short, rate, a), r.ref_bch_decode(short, rate, b)
        assert ca == cb and np.array_equal(a, b)
        mism += ca != 2
    assert mism <= 2  # the all-ones quirk is rare


@pytest.mark.parametrize("short,rate", [(0, 3), (1, 0), (0, 11)])
def test_descrambler_matches_reference(short, rate):
    o, r = orclib.oracle(), orclib.ref()
    p = code_params(short, rate)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, p["K"] // 8, dtype=np.uint8)
    b = a.copy()
    o.orc_descramble(short, rate, a)
    r.ref_descramble(short, rate, b)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("ctype,g1,g2", [(1, 0, 0), (3, 0, 0), (4, 3.15, 0), (4, 2.57, 0), (5, 2.53, 4.30), (5, 2.84, 5.27)])
def test_demapper_matches_reference(ctype, g1, g2):
    o, r = orclib.oracle(), orclib.ref()
    c = o.orc_const_create(ctype, g1, g2)
    bits = o.orc_const_bits(c)
    rng = np.random.default_rng(ctype)
    n = 20000 if ctype != 5 else 3000
    sym = (rng.normal(0, 0.6, (n, 2))).astype(np.float32)
    sym[:50] *= 40  # far outside the LUT grid
    sym[50] = (np.nan, 0.1)
    sym[51] = (1e30, -1e30)
    got = np.zeros((n, bits), np.int8)
    for i in range(n):
        o.orc_demod_soft_lut(c, float(sym[i, 0]), float(sym[i, 1]), got[i])
    want = np.zeros((n, bits), np.int8)
    assert r.ref_demap(ctype, g1, g2, np.ascontiguousarray(sym.reshape(-1)), n, want.reshape(-1)) == bits
    assert np.array_equal(got, want)
    if ctype != 5:
        # the whole LUT: feed the grid centres back
        lut = np.ctypeslib.as_array(o.orc_const_lut(c), shape=(256, 256, bits))
        xs = ((np.arange(256) - 128 + 0.5) / 256 * 1.5).astype(np.float32)
        grid = np.stack(np.meshgrid(xs, xs, indexing="ij"), -1).reshape(-1, 2)
        want = np.zeros((65536, bits), np.int8)
        r.ref_demap(ctype, g1, g2, np.ascontiguousarray(grid.reshape(-1)), 65536, want.reshape(-1))
        assert np.array_equal(lut.reshape(65536, bits), want)
    # modulator
    pts = np.zeros(2, np.float32)
    allsym = np.arange(1 << bits, dtype=np.uint8)
    want_m = np.zeros(2 << bits, np.float32)
    r.ref_mod(ctype, g1, g2, allsym, 1 << bits, want_m)
    for s in range(1 << bits):
        o.orc_mod(c, s, pts)
        assert np.array_equal(pts, want_m[2 * s: 2 * s + 2])
    o.orc_const_destroy(c)


@pytest.mark.parametrize("const,short,rate", [(0, 0, 3), (1, 0, 4), (1, 0, 5), (1, 1, 4), (2, 0, 5), (3, 0, 11), (3, 1, 6)])
def test_deinterleaver_matches_reference(const, short, rate):
    o, r = orclib.oracle(), orclib.ref()
    n = 16200 if short else 64800
    rng = np.random.default_rng(1)
    x = rng.integers(-128, 128, n, dtype=np.int8)
    a = np.zeros(n, np.int8)
    b = np.zeros(n, np.int8)
    o.orc_deinterleave(const, short, rate, x, a)
    r.ref_deinterleave(const, short, rate, x.copy(), b)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("codenum", [0, 1, 2, 17, 4096, 131071, 262141])
def test_pl_scrambling_matches_reference(codenum):
    """S2Scrambling: Rn sequence (through descramble of probe symbols), descramble and scramble, 33282 symbols =
    the longest PLFRAME payload with pilots"""
    o, r = orclib.oracle(), orclib.ref()
    if not hasattr(r, "ref_pl_descramble"):
        pytest.skip("oracle/_ref was built without s2_scrambling.cpp")
    rn = np.zeros(131072, np.uint8)
    o.orc_pl_rn(codenum, rn)
    n = 33282
    x = np.random.default_rng(codenum).normal(0, 1, 2 * n).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    o.orc_pl_descramble(rn, x, n, a)
    r.ref_pl_descramble(codenum, x, n, b, 0)
    assert np.array_equal(a, b)
    o.orc_pl_scramble(rn, x, n, a)
    r.ref_pl_descramble(codenum, x, n, b, 1)
    assert np.array_equal(a, b)
    back = np.zeros_like(x)
    o.orc_pl_descramble(rn, a, n, back)
    assert np.array_equal(back, x)
    assert set(np.unique(rn[:n]).tolist()) == {0, 1, 2, 3}
