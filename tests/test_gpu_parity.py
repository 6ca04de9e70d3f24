"""Parity of the CUDA decode stage (through the C ABI) against the CPU oracle.  Needs a B200.

Bit-exact bar: posterior LLR bytes + iteration counts (LDPC), corrected codeword bytes + correction
counts (BCH), BBFRAME bytes (whole chain), int8 LLRs (LUT demapper).  32APSK demapping uses device
expf/logf instead of glibc's: tolerance |diff| <= 1 LSB on at most 0.5 % of the LLRs (stated below)."""
import ctypes as C

import numpy as np
import pytest

import orclib
from fec import pkg, MODCODS, QPSK_MODCOD_OF_RATE
from orclib import ALL_CODES, code_params

pytestmark = pytest.mark.gpu

_SNR = {0: -2.0, 1: -1.0, 2: 0.0, 3: 1.3, 4: 2.6, 5: 3.4, 6: 4.3, 7: 5.0, 8: 5.5, 10: 6.5, 11: 6.7}


@pytest.fixture(scope="module")
def dec():
    d = pkg.DVBS2Decoder(max_batch=256, max_trials=25)
    yield d
    d.close()


def make_llrs(short, rate, n, seed, spread=1.2):
    """n frames around the code's threshold: converging, slow and failing ones, plus corner cases"""
    p = code_params(short, rate)
    rng = np.random.default_rng(seed)
    out = np.zeros((n, p["N"]), np.int8)
    for i in range(n):
        _, code = orclib.encode_frame(short, rate, rng)
        snr = _SNR[rate] + (0.4 if short else 0) + spread * (i / max(n - 1, 1) - 0.35)
        out[i] = orclib.awgn_llr(code, snr, rng)
    return out


def oracle_ldpc(short, rate, llr, max_trials):
    o = orclib.oracle()
    post = llr.copy()
    iters = np.zeros(len(llr), np.int16)
    for i in range(len(llr)):
        iters[i] = o.orc_ldpc_decode(short, rate, post[i], max_trials)
    return post, iters


@pytest.mark.parametrize("short,rate", ALL_CODES)
def test_ldpc_bit_exact_all_codes(dec, short, rate):
    dec.setDemodParams(QPSK_MODCOD_OF_RATE[rate], bool(short), False)
    n = 7 if not short else 13  # odd: exercises the half-empty last pair
    llr = make_llrs(short, rate, n, 17 + rate + 40 * short)
    llr[1, ::97] = 0                      # zero LLRs keep rows "bad"
    llr[2] = np.where(llr[2] < 0, -128, 127)  # saturated input
    want_post, want_it = oracle_ldpc(short, rate, llr, 25)
    got = llr.copy()
    it = dec.ldpc_decode(got, 25)
    assert np.array_equal(it, want_it), (it, want_it)
    assert np.array_equal(got, want_post)
    assert len(set(want_it.tolist())) >= 2


@pytest.mark.parametrize("max_trials", [1, 3, 16])
def test_ldpc_iteration_cap(dec, max_trials):
    short, rate = 1, 3
    dec.setDemodParams(4, True, False)
    llr = make_llrs(short, rate, 8, 5)
    want_post, want_it = oracle_ldpc(short, rate, llr, max_trials)
    got = llr.copy()
    it = dec.ldpc_decode(got, max_trials)
    assert np.array_equal(it, want_it)
    assert np.array_equal(got, want_post)


def test_ldpc_random_and_codeword_inputs(dec):
    short, rate = 1, 6
    dec.setDemodParams(7, True, False)
    p = code_params(short, rate)
    rng = np.random.default_rng(3)
    llr = rng.integers(-128, 128, (6, p["N"]), dtype=np.int8)
    _, code = orclib.encode_frame(short, rate, rng)
    llr[4] = np.where(code > 0, -9, 9)   # already a codeword: 0 iterations
    llr[5] = 0                            # all-zero LLRs
    want_post, want_it = oracle_ldpc(short, rate, llr, 25)
    got = llr.copy()
    it = dec.ldpc_decode(got, 25)
    assert want_it[4] == 0
    assert np.array_equal(it, want_it)
    assert np.array_equal(got, want_post)


@pytest.mark.parametrize("short,rate", [(0, 3), (0, 5), (0, 10), (1, 3), (1, 0)])
def test_bch_bit_exact(dec, short, rate):
    dec.setDemodParams(QPSK_MODCOD_OF_RATE[rate], bool(short), False)
    o = orclib.oracle()
    p = code_params(short, rate)
    K, kbch, t = p["K"], p["kbch"], p["t"]
    rng = np.random.default_rng(11 + rate)
    nerrs = [0, 1, 2, 2, 2, 3, 5, t - 1, t, t, t + 1, t + 2, 30, 200]
    frames = np.zeros((len(nerrs), K // 8), np.uint8)
    for i, ne in enumerate(nerrs):
        frames[i, : kbch // 8] = rng.integers(0, 256, kbch // 8, dtype=np.uint8)
        assert o.orc_bch_encode(short, rate, frames[i]) == 0
        where = rng.choice(K, ne, replace=False)
        if i == 4:
            where = np.array([K - 1, K - 2])  # last two parity bits
        for e in where:
            frames[i, e >> 3] ^= 0x80 >> (e & 7)
    want = frames.copy()
    want_c = np.array([o.orc_bch_decode(short, rate, want[i]) for i in range(len(nerrs))], np.int16)
    got = frames.copy()
    got_c = dec.bch_decode(got)
    assert np.array_equal(got_c, want_c), (got_c, want_c)
    assert np.array_equal(got, want)
    assert (want_c[:10] == np.array(nerrs[:10])).all()


def test_descrambler_bit_exact(dec):
    dec.setDemodParams(4, False, False)
    o = orclib.oracle()
    rng = np.random.default_rng(2)
    fr = rng.integers(0, 256, (3, dec.K // 8), dtype=np.uint8)
    want = fr.copy()
    for i in range(3):
        o.orc_descramble(0, 3, want[i])
    assert np.array_equal(dec.descramble(fr.copy()), want)


@pytest.mark.parametrize("modcod,short", [(4, 0), (4, 1), (12, 0), (13, 0), (13, 1), (18, 0), (21, 1), (23, 0)])
def test_demapper_lut_bit_exact(dec, modcod, short):
    const, ctype, rate, g1, g2 = MODCODS[modcod]
    dec.setDemodParams(modcod, bool(short), False)
    o = orclib.oracle()
    c = o.orc_const_create(ctype, g1, g2)
    rng = np.random.default_rng(modcod)
    n = 2
    pl = np.zeros((n, dec.plframe_symbols, 2), np.float32)
    for i in range(n):
        bits = rng.integers(0, 2, dec.N, dtype=np.uint8)
        pl[i] = pkg.modulate(modcod, bool(short), False, bits).view(np.float32).reshape(-1, 2)
    pl += rng.normal(0, 0.15, pl.shape).astype(np.float32)
    pl[0, 100:140] *= 30            # clipped by the LUT index clamp
    pl[0, 150] = (np.nan, 0.3)
    pl[0, 151] = (3e30, -3e30)
    want = np.zeros((n, dec.N), np.int8)
    for i in range(n):
        o.orc_bb_to_soft(c, const, short, rate, np.ascontiguousarray(pl[i].reshape(-1)), want[i])
    got = dec.bb_to_soft(pl)
    assert np.array_equal(got, want)
    o.orc_const_destroy(c)


@pytest.mark.parametrize("modcod,short,pilots,codenum", [(4, 0, False, -1), (12, 0, False, -1), (13, 1, True, 5), (18, 0, False, -1), (21, 1, True, 0)])
def test_quantised_symbols_path_gives_identical_frames(dec, modcod, short, pilots, codenum):
    """dvbs2fec_quantize_plframes (host: the reference's LUT index arithmetic incl. NaN / out-of-range, pilot removal,
    PL descrambling) + dvbs2fec_decode_plframes_idx == dvbs2fec_decode_plframes on the same symbols, frame by
    frame; also through the queue"""
    try:
        dec.set_pl_scrambling(codenum)
        dec.setDemodParams(modcod, bool(short), pilots)
        rng = np.random.default_rng(500 + modcod)
        n = 5
        payload = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
        pl = np.stack([pkg.modulate(modcod, bool(short), pilots, pkg.encode_fecframe(modcod, bool(short), payload[i])) for i in range(n)])
        x = pl.view(np.float32).reshape(n, -1) + rng.normal(0, 0.06, (n, dec.plframe_symbols * 2)).astype(np.float32)
        if codenum >= 0:
            o = orclib.oracle()
            rn = np.zeros(131072, np.uint8)
            o.orc_pl_rn(codenum, rn)
            nsym = dec.plframe_symbols - 90
            for i in range(n):
                out = np.zeros(2 * nsym, np.float32)
                o.orc_pl_scramble(rn, np.ascontiguousarray(x[i, 180:]), nsym, out)
                x[i, 180:] = out
        x[0, 400:480] *= 30
        x[0, 500] = np.nan
        x[0, 503] = 3e30
        x[1, 600] = -np.inf
        want_bb, want_res = dec.decode_plframes(x)
        idx = dec.quantize_plframes(x)
        assert idx.shape == (n, dec.N // pkg.modcod_info(modcod, bool(short))["bits"], 2)
        bb, res = dec.decode_plframes_idx(idx)
        assert np.array_equal(bb, want_bb)
        for k in ("ldpc_iters", "bch_corr", "flags"):
            assert np.array_equal(res[k], want_res[k])
        assert np.array_equal(bb[2:], payload[2:])
        for i in range(n):
            dec.submit_plframe_idx(idx[i], 40 + i)
        dec.flush()
        qbb, qres = dec.collect(n, timeout_us=5_000_000)
        assert np.array_equal(qbb, want_bb) and list(qres["tag"]) == list(range(40, 40 + n))
    finally:
        dec.set_pl_scrambling(-1)
    dec.setDemodParams(28, False, False)
    with pytest.raises(pkg.DVBS2FecError):
        dec.quantize_plframes(np.zeros((1, dec.plframe_symbols * 2), np.float32))


@pytest.mark.parametrize("modcod,short", [(24, 1), (28, 0)])
def test_demapper_32apsk_within_one_lsb(dec, modcod, short):
    const, ctype, rate, g1, g2 = MODCODS[modcod]
    dec.setDemodParams(modcod, bool(short), False)
    o = orclib.oracle()
    c = o.orc_const_create(ctype, g1, g2)
    rng = np.random.default_rng(modcod)
    bits = rng.integers(0, 2, dec.N, dtype=np.uint8)
    pl = pkg.modulate(modcod, bool(short), False, bits).view(np.float32).reshape(1, -1, 2).copy()
    pl += rng.normal(0, 0.05, pl.shape).astype(np.float32)
    want = np.zeros((1, dec.N), np.int8)
    o.orc_bb_to_soft(c, const, short, rate, np.ascontiguousarray(pl.reshape(-1)), want[0])
    got = dec.bb_to_soft(pl)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    # tolerance: device expf/logf vs glibc -> at most 1 LSB, on at most 0.5 % of the LLRs
    assert diff.max() <= 1
    assert (diff != 0).mean() <= 0.005
    o.orc_const_destroy(c)


def oracle_chain(short, rate, llr, max_trials):
    o = orclib.oracle()
    p = code_params(short, rate)
    bb = np.zeros((len(llr), p["kbch"] // 8), np.uint8)
    its = np.zeros(len(llr), np.int16)
    cor = np.zeros(len(llr), np.int16)
    for i in range(len(llr)):
        a, b = C.c_int(), C.c_int()
        o.orc_decode_frame(short, rate, llr[i].copy(), max_trials, bb[i], C.byref(a), C.byref(b))
        its[i], cor[i] = a.value, b.value
    return bb, its, cor


@pytest.mark.parametrize("short,rate,n", [(0, 3, 24), (1, 3, 40), (0, 11, 6), (1, 10, 30)])
def test_whole_stage_bit_exact(dec, short, rate, n):
    dec.setDemodParams(QPSK_MODCOD_OF_RATE[rate], bool(short), False)
    llr = make_llrs(short, rate, n, 99 + rate, spread=0.9)
    want_bb, want_it, want_c = oracle_chain(short, rate, llr, 25)
    bb, res = dec.decode_batch(llr)
    assert np.array_equal(res["ldpc_iters"], want_it)
    assert np.array_equal(res["bch_corr"], want_c)
    assert np.array_equal(bb, want_bb)
    assert np.array_equal(res["tag"], np.arange(n))
    flags = (want_it < 0) * 1 + (want_c < 0) * 2
    assert np.array_equal(res["flags"] & 3, flags)
    crc_bad = np.array([orclib.oracle().orc_bbheader_crc8(np.ascontiguousarray(want_bb[i])) != 0 for i in range(n)])
    assert np.array_equal((res["flags"] & 4) != 0, crc_bad)
    assert dec.last_launch_count() >= 2


def test_payload_round_trip_from_symbols(dec):
    """transmit -> AWGN -> demap + LDPC + BCH + descramble returns the payload (config 3/4 style chain)"""
    for modcod, short, sigma in [(4, 1, 0.10), (13, 0, 0.10), (18, 1, 0.05), (27, 1, 0.02)]:
        dec.setDemodParams(modcod, bool(short), False)
        rng = np.random.default_rng(modcod)
        n = 4
        payload = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
        pl = np.zeros((n, dec.plframe_symbols), np.complex64)
        for i in range(n):
            pl[i] = pkg.modulate(modcod, bool(short), False, pkg.encode_fecframe(modcod, bool(short), payload[i]))
        noisy = pl.view(np.float32) + rng.normal(0, sigma, (n, dec.plframe_symbols * 2)).astype(np.float32)
        bb, res = dec.decode_plframes(noisy)
        assert (res["bch_corr"] >= 0).all(), (modcod, res)
        assert np.array_equal(bb, payload), modcod


def test_queue_returns_frames_in_submission_order(dec):
    short, rate = 1, 3
    dec.setDemodParams(4, True, False)
    llr = make_llrs(short, rate, 21, 1234)
    want_bb, want_it, want_c = oracle_chain(short, rate, llr, 25)
    for i in range(len(llr)):
        dec.submit_llr(llr[i], 1000 + i)
    dec.flush()
    bb, res = dec.collect(64, timeout_us=2_000_000)
    assert len(res) == len(llr)
    assert np.array_equal(res["tag"], 1000 + np.arange(len(llr)))
    assert np.array_equal(bb, want_bb)
    assert np.array_equal(res["ldpc_iters"], want_it)
    assert np.array_equal(res["bch_corr"], want_c)


def test_queue_many_small_batches_partial_collects_and_backpressure():
    """max_batch 4: 37 frames pass through ten staging batches; collect takes odd-sized bites; a producer that
    never collects eventually gets EAGAIN ("queue full") and loses nothing once it does collect"""
    short, rate = 1, 3
    d = pkg.DVBS2Decoder(max_batch=4, max_latency_us=500, max_trials=25)
    try:
        d.setDemodParams(4, True, False)
        llr = make_llrs(short, rate, 37, 4321)
        want_bb, want_it, want_c = oracle_chain(short, rate, llr, 25)
        got_bb, got_res = [], []
        refused = 0
        for i in range(len(llr)):
            while True:
                rc = pkg.lib().dvbs2fec_submit_llr(d._h, llr[i].ctypes.data_as(C.c_void_p), i)
                if rc != pkg.EAGAIN:
                    break
                refused += 1   # all four staging batches hold uncollected frames: take a bite and retry
                bb, res = d.collect(3, timeout_us=100_000)
                got_bb.append(bb.copy()); got_res.append(res.copy())
            assert rc == 0
            if i % 5 == 4:
                bb, res = d.collect(3, timeout_us=0)
                got_bb.append(bb.copy()); got_res.append(res.copy())
        assert refused > 0
        d.flush()
        while sum(len(r) for r in got_res) < len(llr):
            bb, res = d.collect(7, timeout_us=2_000_000)
            assert len(res) > 0
            got_bb.append(bb.copy()); got_res.append(res.copy())
        bb, res = np.concatenate(got_bb), np.concatenate(got_res)
        assert np.array_equal(res["tag"], np.arange(len(llr)))
        assert np.array_equal(bb, want_bb)
        assert np.array_equal(res["ldpc_iters"], want_it)
        assert np.array_equal(res["bch_corr"], want_c)
        # back-pressure: 4 staging batches of up to 4 frames (a batch may close early on its latency deadline)
        accepted = 0
        for i in range(40):
            rc = pkg.lib().dvbs2fec_submit_llr(d._h, llr[i % len(llr)].ctypes.data_as(C.c_void_p), 100 + i)
            if rc == pkg.EAGAIN:
                break
            assert rc == 0
            accepted += 1
        assert 4 <= accepted <= 16
        d.flush()
        tags = []
        while len(tags) < accepted:
            _, res = d.collect(64, timeout_us=2_000_000)
            assert len(res) > 0
            tags += list(res["tag"])
        assert tags == list(range(100, 100 + accepted))
    finally:
        d.close()


def test_queue_mixed_llr_and_plframe_inputs_keep_order(dec):
    """LLR frames and PLFRAMEs submitted alternately in runs: each run becomes its own batch, order is kept"""
    modcod, short = 4, True
    dec.setDemodParams(modcod, short, False)
    rng = np.random.default_rng(77)
    n = 9
    payload = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
    codes = [pkg.encode_fecframe(modcod, short, payload[i]) for i in range(n)]
    for i in range(n):
        if (i // 3) % 2 == 0:
            dec.submit_llr(np.where(codes[i] > 0, -40, 40).astype(np.int8), i)
        else:
            pl = pkg.modulate(modcod, short, False, codes[i]).view(np.float32)
            dec.submit_plframe(pl + rng.normal(0, 0.05, pl.shape).astype(np.float32), i)
    dec.flush()
    bb, res = dec.collect(64, timeout_us=2_000_000)
    assert list(res["tag"]) == list(range(n))
    assert (res["bch_corr"] >= 0).all()
    assert np.array_equal(bb, payload)


def test_modcod_switch_and_errors(dec):
    with pytest.raises(pkg.DVBS2FecError):
        dec.setDemodParams(0)
    with pytest.raises(pkg.DVBS2FecError):
        dec.setDemodParams(29)
    with pytest.raises(pkg.DVBS2FecError):
        dec.setDemodParams(11, True)      # short 9/10 does not exist (reference leaves ldpc uninitialised)
    dec.setDemodParams(6, False, False)
    assert (dec.N, dec.K, dec.kbch) == (64800, 43200, 43040)
    dec.setDemodParams(1, True, True)
    assert (dec.N, dec.K, dec.kbch) == (16200, 3240, 3072)
    assert dec.plframe_symbols == 90 + 8100 + 36 * 5


def test_stage_objects_mirror_reference_interface():
    rng = np.random.default_rng(8)
    ldpc = pkg.BBFrameLDPC(True, 3)
    bch = pkg.BBFrameBCH(True, 3, decoder=ldpc.dec)
    descr = pkg.BBFrameDescrambler(True, 3, decoder=ldpc.dec)
    assert ldpc.dataSize() == 7200 and bch.dataSize() == 7032
    payload, code = orclib.encode_frame(1, 3, rng)
    frame = orclib.awgn_llr(code, 3.0, rng)
    want = frame.copy()
    want_it = orclib.oracle().orc_ldpc_decode(1, 3, want, 16)
    assert ldpc.decode(frame, 16) == want_it
    assert np.array_equal(frame, want)
    packed = np.packbits((frame[:7200] < 0).astype(np.uint8))
    assert bch.decode(packed) == 0
    descr.work(packed)
    orclib.oracle().orc_descramble(1, 3, payload_copy := np.concatenate([payload, np.zeros(21, np.uint8)]))
    assert np.array_equal(packed[: 7032 // 8], payload_copy[: 7032 // 8])
    ldpc.dec.close()


def test_bbheader_crc_flag(dec):
    """flag bit 4 mirrors BBFrameTSParser's CRC-8 gate: clear for a well-formed BBHEADER, set otherwise"""
    o = orclib.oracle()
    dec.setDemodParams(4, True, False)
    rng = np.random.default_rng(77)
    n = 12
    payload = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
    for i in range(0, n, 2):   # make every other header valid: choose the CRC byte that zeroes the check
        for b in range(256):
            payload[i, 9] = b
            if o.orc_bbheader_crc8(np.ascontiguousarray(payload[i])) == 0:
                break
        else:
            raise AssertionError("no CRC byte found")
    llr = np.stack([np.where(pkg.encode_fecframe(4, True, payload[i]) > 0, -10, 10).astype(np.int8) for i in range(n)])
    bb, res = dec.decode_batch(llr)
    assert np.array_equal(bb, payload)
    want = np.array([o.orc_bbheader_crc8(np.ascontiguousarray(payload[i])) != 0 for i in range(n)])
    assert (~want[0::2]).all()
    assert np.array_equal((res["flags"] & pkg.FLAG_BBHEADER_CRC_FAIL) != 0, want)


@pytest.mark.parametrize("modcod,short,pilots,codenum", [(4, True, False, 0), (13, True, True, 5), (18, True, False, 262141),
                                                         (4, False, True, 1)])
def test_pl_descrambling_inside_the_demapper(dec, modcod, short, pilots, codenum):
    """PL-scrambled symbols (S2Scrambling::scramble restated by the oracle) with dvbs2fec_set_pl_scrambling(codenum)
    give exactly the LLRs -- and BBFRAMEs -- of the descrambled symbols with the descrambler off"""
    o = orclib.oracle()
    try:
        dec.set_pl_scrambling(-1)
        dec.setDemodParams(modcod, short, pilots)
        rng = np.random.default_rng(modcod + codenum % 97)
        n = 3
        payload = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
        pl = np.stack([pkg.modulate(modcod, short, pilots, pkg.encode_fecframe(modcod, short, payload[i])) for i in range(n)])
        clean = pl.view(np.float32).reshape(n, -1) + rng.normal(0, 0.05, (n, dec.plframe_symbols * 2)).astype(np.float32)
        rn = np.zeros(131072, np.uint8)
        o.orc_pl_rn(codenum, rn)
        scr = clean.copy()
        nsym = dec.plframe_symbols - 90
        for i in range(n):
            out = np.zeros(2 * nsym, np.float32)
            o.orc_pl_scramble(rn, np.ascontiguousarray(clean[i, 180:]), nsym, out)
            scr[i, 180:] = out
        want_llr = dec.bb_to_soft(clean)
        want_bb, want_res = dec.decode_plframes(clean)
        dec.set_pl_scrambling(codenum)
        assert np.array_equal(dec.bb_to_soft(scr), want_llr)
        bb, res = dec.decode_plframes(scr)
        assert np.array_equal(bb, want_bb) and np.array_equal(res["ldpc_iters"], want_res["ldpc_iters"])
        assert np.array_equal(bb, payload)
        assert not np.array_equal(dec.bb_to_soft(clean), want_llr)   # the descrambler really is on
    finally:
        dec.set_pl_scrambling(-1)
    with pytest.raises(pkg.DVBS2FecError):
        dec.set_pl_scrambling(262142)


def test_zero_copy_submit_acquire_commit(dec):
    """frames written straight into the page-locked batch (dvbs2fec_acquire_llr / dvbs2fec_commit), mixed with copying
    submits, come back in order and equal to decode_batch; a second acquire without commit is refused"""
    from test_gpu_properties import make_batch
    dec.setDemodParams(4, True, False)
    llr, payload = make_batch(4, True, 11, 3.0, 77)
    want_bb, want_res = dec.decode_batch(llr)
    d = pkg.DVBS2Decoder(max_batch=4, max_latency_us=300)
    d.setDemodParams(4, True, False)
    for i in range(11):
        if i % 3 == 2:
            d.submit_llr(llr[i], i)
        else:
            slot = d.acquire_llr()
            if i == 0:
                with pytest.raises(pkg.DVBS2FecError):
                    d.acquire_llr()
                with pytest.raises(pkg.DVBS2FecError):
                    d.submit_llr(llr[i], 99)
            slot[:] = llr[i]
            if i == 5:
                import time
                time.sleep(0.01)   # longer than max_latency_us: the partial batch must wait for this frame
            d.commit(i)
    d.flush()
    bb, res = d.collect(16, timeout_us=5_000_000)
    while len(res) < 11:
        b2, r2 = d.collect(16, timeout_us=5_000_000)
        bb, res = np.concatenate([bb, b2]), np.concatenate([res, r2])
    assert list(res["tag"]) == list(range(11))
    assert np.array_equal(bb, want_bb) and np.array_equal(res["ldpc_iters"], want_res["ldpc_iters"])
    d.close()
