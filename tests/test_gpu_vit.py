"""Row 8(f)-4, inner half, on the GPU: dvbs2fec_dvbs_viterbi_* / dvbs2fec_dvbs_sts_* against the CPU oracle, which
tests/test_vit_oracle.py pins to the reference's own sources.  Bit-exact: decoded bits, BER, lock state after every call."""
import numpy as np
import pytest

import dvbs_stream
from fec import pkg
from test_vit_oracle import CASES, OrcViterbi, same_stats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_viterbi_matches_oracle(case):
    """search, lock, decode at every rate, both phases, odd puncturing offsets; calls of one to five blocks"""
    rate, sigma, phase, lead, seed = case
    rng = np.random.default_rng(100 + seed)
    bits = rng.integers(0, 2, 130000, dtype=np.uint8)
    s = dvbs_stream.inner_softs(bits, rate, rng, sigma=sigma, phase=phase, lead=lead)
    nb = min(len(s) // 8192, 14)
    o, g = OrcViterbi(), pkg.DVBSViterbi()
    k = 0
    while k < nb:
        n = min(int(rng.integers(1, 6)), nb - k)
        seg = s[k * 8192:(k + n) * 8192]
        want = o.process(seg, fill=k & 1)
        got = g.process(seg, out=np.full(len(seg), k & 1, np.uint8))
        assert got.shape == want.shape and np.array_equal(got, want), (k, n)
        assert same_stats(g.stats(), o.stats()), (k, g.stats(), o.stats())
        k += n
    assert g.stats()[1] == 1 and g.stats()[2] == rate
    g.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_viterbi_lock_lost_and_found_matches_oracle(seed):
    """noise, signal, noise until the lock is given up, another rate, degenerate input -- in ONE call (the batch is cut
    where the state changes) and again block by block"""
    rng = np.random.default_rng(200 + seed)
    r1, r2 = [(1, 3), (4, 0), (2, 1)][seed]
    noise = lambda n: np.clip(np.rint(rng.normal(0, 40, n * 8192)), -128, 127).astype(np.int8)
    sig = lambda r, n, lead: dvbs_stream.inner_softs(rng.integers(0, 2, 8192 * n, dtype=np.uint8), r, rng, sigma=[12.0, 10.0, 9.0, 7.0, 5.0][r],
                                                     lead=lead)[:n * 8192]
    s = np.concatenate([noise(2), sig(r1, 4, 2), noise(5), sig(r2, 4, 0), np.zeros(8192, np.int8), np.full(8192, -128, np.int8), noise(7)])
    o, g = OrcViterbi(0.15, 3), pkg.DVBSViterbi(0.15, 3)
    want = o.process(s, fill=7)
    got = g.process(s, out=np.full(len(s), 7, np.uint8))
    assert np.array_equal(got, want) and same_stats(g.stats(), o.stats())
    g.reset()
    o = OrcViterbi(0.15, 3)
    for k in range(len(s) // 8192):
        seg = s[k * 8192:(k + 1) * 8192]
        want = o.process(seg, fill=7)
        got = g.process(seg, out=np.full(8192, 7, np.uint8))
        assert np.array_equal(got, want), k
        assert same_stats(g.stats(), o.stats()), (k, g.stats(), o.stats())
    g.close()


@pytest.mark.parametrize("rate", [0, 1, 2, 3, 4])
def test_viterbi_large_batch(rate):
    """200 blocks in one call: equal to the oracle; the transmitted bits come back (rate 5/6: where the reference decodes)"""
    rng = np.random.default_rng(300 + rate)
    nbits = [4096, 5462, 6144, 6827, 7168][rate] * 200
    bits = rng.integers(0, 2, nbits, dtype=np.uint8)
    s = dvbs_stream.inner_softs(bits, rate, rng, sigma=[14.0, 11.0, 10.0, 8.0, 6.0][rate])
    s = s[:len(s) // 8192 * 8192]
    g = pkg.DVBSViterbi()
    got = g.process(s)
    want = OrcViterbi().process(s)
    assert np.array_equal(got, want)
    assert g.stats()[1:3] == (1, rate)
    if rate != 3:
        assert np.mean(got[200:-200] == bits[200:len(got) - 200]) > 0.999
    g.close()


def test_syms_to_soft_matches_oracle():
    import orclib
    rng = np.random.default_rng(5)
    o = orclib.oracle()
    ho = o.orc_sts_create()
    g = pkg.DVBSViterbi()
    for n in [1, 4095, 1, 5000, 0, 12289, 3, 100000]:
        syms = (rng.normal(0, 0.9, (n, 2))).astype(np.float32)
        if n > 10:
            syms[3:7] = [[1.27, -1.27], [1.2701, -1.2799], [5.0, -7.0], [0.0049, -0.0099]]
        want = np.zeros(2 * n + 8192, np.int8)
        k = o.orc_sts_process(ho, n, np.ascontiguousarray(syms.reshape(-1)), want)
        got = g.syms_to_soft(syms)
        assert len(got) == k and np.array_equal(got, want[:k])
    g.close()


def test_viterbi_argument_errors():
    g = pkg.DVBSViterbi()
    with pytest.raises(pkg.DVBS2FecError):
        g.process(np.zeros(1000, np.int8))
    assert len(g.process(np.zeros(0, np.int8))) == 0
    with pytest.raises(pkg.DVBS2FecError):
        pkg.DVBSViterbi(device=99)
    g.close()


# ---- the whole decode stage of the module: symbols -> TS packets ------------------------------------------------------------
class OracleChain:
    """DVBSDemod::process (module_dvbs_demod.cpp:78-100) behind demod.process, out of the oracle's pieces (each pinned to the
    reference's class); the bit buffer persists from call to call like the handle's"""

    def __init__(self, stride):
        import orclib
        from test_dvbs_oracle import OrcDeframer, OrcOuter
        self.o = orclib.oracle()
        self.sts = self.o.orc_sts_create()
        self.vit, self.defr, self.outer, self.stride = OrcViterbi(), OrcDeframer(), OrcOuter(), stride
        self.bits = np.zeros(0, np.uint8)

    def process(self, syms):
        x = np.ascontiguousarray(syms, np.complex64).view(np.float32).reshape(-1)
        n = len(x) // 2
        soft = np.zeros(2 * n + 8192, np.int8)
        k = self.o.orc_sts_process(self.sts, n, x, soft)
        if not k:
            return np.zeros((0, 188), np.uint8)
        if len(self.bits) < 2 * n + 8192 + 8192:
            self.bits = np.concatenate([self.bits, np.zeros(2 * n + 8192 + 8192 - len(self.bits), np.uint8)])
        nb = self.vit.proc(self.vit.h, k, soft[:k].copy(), self.bits)
        frames, _ = self.defr.work(self.bits[:nb])
        nfr = min(len(frames), nb // 1632 + 8)
        if not nfr:
            return np.zeros((0, 188), np.uint8)
        ts, err = self.outer.process(frames.reshape(-1), nfr, self.stride)
        return ts


def dvbs_symbols(nframes, rate, rng, sigma=0.08, lead=0):
    """TS packets -> outer code -> inner code -> QPSK symbols (bit 1 = +0.6) with noise"""
    ts, ch = dvbs_stream.outer_stream(nframes, rng)
    bits = np.unpackbits(ch)
    x, y = dvbs_stream.conv_encode(bits)
    tx = dvbs_stream.puncture(x, y, rate).astype(np.float32) * 2 - 1
    tx = np.concatenate([rng.normal(0, 0.3, lead).astype(np.float32), tx])
    tx = tx[:len(tx) // 2 * 2] * 0.6 + rng.normal(0, sigma, len(tx) // 2 * 2).astype(np.float32)
    return ts, (tx[0::2] + 1j * tx[1::2]).astype(np.complex64)


@pytest.mark.parametrize("case", [(0, 1632, 0), (1, 1632, 1), (2, 204, 2), (4, 1632, 3), (3, 1632, 4)])
def test_demod_chain_matches_oracle_and_recovers_the_packets(case):
    """symbols in, TS packets out, in calls of uneven sizes: equal to the chain of oracle pieces; with back-to-back frames
    (stride 1632) and a rate the reference decodes properly the transmitted packets come back"""
    rate, stride, seed = case
    rng = np.random.default_rng(400 + seed)
    ts, syms = dvbs_symbols(40, rate, rng, lead=2 * int(rng.integers(0, 50)))
    o, g = OracleChain(stride), pkg.DVBSDemod(frame_stride=stride)
    cuts = sorted(set(int(c) for c in rng.integers(0, len(syms), 4)) | {0, len(syms)})
    got_all = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        want = o.process(syms[lo:hi])
        got = g.process(syms[lo:hi])
        assert got.shape == want.shape and np.array_equal(got, want), (lo, hi)
        got_all.append(got)
    st = g.stats()
    assert st["viterbi_lock"] == 1 and st["viterbi_rate"] == rate
    got_all = np.concatenate(got_all)
    if stride == 1632 and rate != 3:
        assert len(got_all) >= 8 * 20
        sent = {p.tobytes() for p in ts}
        good = sum(p.tobytes() in sent for p in got_all[24:])
        assert good >= len(got_all) - 24 - 8
    g.close()


def test_demod_chain_reset_stats_and_errors():
    """reset() gives the behaviour of a fresh handle; the module's statistics; ENOSPC when the TS output does not fit; bad
    arguments are refused without touching the state"""
    import ctypes as C
    rng = np.random.default_rng(9)
    ts, syms = dvbs_symbols(30, 0, rng, lead=14)
    g = pkg.DVBSDemod(frame_stride=1632)
    first = g.process(syms)
    st = g.stats()
    assert len(first) > 100 and st["viterbi_lock"] == 1 and st["viterbi_rate"] == 0 and st["frames_done"] * 8 == len(first)
    assert st["deframer_err"] == 0 and st["rs_avg"] == 0.0 and 0 <= st["viterbi_ber"] < 0.15
    g.reset()
    assert np.array_equal(g.process(syms), first)
    g.reset()
    L = pkg.lib()
    x = np.ascontiguousarray(syms).view(np.float32).reshape(-1)
    small = np.zeros(1504, np.uint8)
    rc = L.dvbs2fec_dvbs_demod_process(g._p, len(syms), x.ctypes.data_as(C.c_void_p), small.ctypes.data_as(C.c_void_p), len(small))
    assert rc == -28      # DVBS2FEC_ENOSPC
    assert L.dvbs2fec_dvbs_demod_process(g._p, -1, None, None, 0) == -22
    h = C.c_void_p()
    assert L.dvbs2fec_dvbs_demod_create(0, 0.15, 20, 100, C.byref(h)) == -22      # frame stride is 204 or 1632
    g.close()


def test_deframer_argument_errors():
    import ctypes as C
    L = pkg.lib()
    d = pkg.DVBSTSDeframer()
    one = np.zeros(8, np.uint8)
    assert L.dvbs2fec_dvbs_deframer_work(d._p, one.ctypes.data_as(C.c_void_p), -1, one.ctypes.data_as(C.c_void_p), 1) == -22
    assert L.dvbs2fec_dvbs_deframer_work(d._p, None, 8, one.ctypes.data_as(C.c_void_p), 1) == -22
    assert L.dvbs2fec_dvbs_deframer_work_device(d._p, one.ctypes.data_as(C.c_void_p), 1 << 24, one.ctypes.data_as(C.c_void_p), 1, None, None) == -22
    assert d.work(np.ones(5000, np.uint8)).shape == (0, 1632)
    d.close()


def test_viterbi_calls_longer_than_the_internal_batches(monkeypatch):
    """700 locked blocks in one call with the decode batch limited to 256 blocks (8192 by default), and 150 blocks of
    noise (search batches grow to 64 blocks): equal to the oracle"""
    rng = np.random.default_rng(808)
    bits = rng.integers(0, 2, 4096 * 701, dtype=np.uint8)
    s = dvbs_stream.inner_softs(bits, 0, rng, sigma=14.0)[:700 * 8192]
    monkeypatch.setenv("DVBS2FEC_VIT_MAX_BATCH", "256")
    g = pkg.DVBSViterbi()
    got = g.process(s)
    want = OrcViterbi().process(s)
    assert np.array_equal(got, want) and g.stats()[1:3] == (1, 0)
    g.reset()
    noise = np.clip(np.rint(rng.normal(0, 40, 150 * 8192)), -127, 127).astype(np.int8)
    o = OrcViterbi()
    assert len(g.process(noise)) == len(o.process(noise)) == 0
    assert same_stats(g.stats(), o.stats())
    tail = s[:8 * 8192]
    assert np.array_equal(g.process(tail), o.process(tail)) and same_stats(g.stats(), o.stats())
    g.close()
