"""Row 8(f)-4, inner half: the CPU restatement of the self-locking punctured Viterbi decoder and of the soft-symbol block
(oracle/oracle_vit.c) against the reference's own viterbi_all.cpp / cc_decoder.cpp / cc_encoder.cpp / depunc.h /
rotation.cpp / dvbs_syms_to_soft.cpp compiled unmodified (oracle/_ref).  Byte and index work: bit-exact, lock state and
BER (a float made of two integer counts) included."""
import ctypes as C

import numpy as np
import pytest

import dvbs_stream
import orclib


class OrcViterbi:
    def __init__(self, thr=0.15, max_outsync=20):
        self.o = orclib.oracle()
        self.h = self.o.orc_vit_create(thr, max_outsync)
        self.proc, self.stat = self.o.orc_vit_process, self.o.orc_vit_stats

    def process(self, softs, fill=0):
        """-> decoded bits (one per byte).  The output buffer starts from `fill`: the reference leaves holes at rate 5/6"""
        x = np.array(softs, np.int8)          # (the reference rotates its input in place)
        out = np.full(len(x) + 8192, fill, np.uint8)
        n = self.proc(self.h, len(x), x, out)
        return out[:n].copy()

    def stats(self):
        b = C.c_float()
        v = [C.c_int() for _ in range(5)]
        self.stat(self.h, C.byref(b), *[C.byref(i) for i in v])
        return (b.value,) + tuple(i.value for i in v)      # ber, state, rate, phase, shift, invalid


class RefViterbi(OrcViterbi):
    def __init__(self, thr=0.15, max_outsync=20):
        self.o = orclib.ref()
        self.h = self.o.ref_vit_create(thr, max_outsync)
        assert self.o.ref_vit_layout_ok(self.h)
        self.proc, self.stat = self.o.ref_vit_process, self.o.ref_vit_stats


def same_stats(a, b):
    """rate / phase / shift mean something only while locked"""
    if not (a[0] == b[0] or (np.isnan(a[0]) and np.isnan(b[0]))) or a[1] != b[1] or a[5] != b[5]:
        return False
    return a[1] == 0 or a[2:5] == b[2:5]


def test_decoder_locks_and_decodes_every_rate():
    """the transmitted bits come back (after the lock, away from the block edges) at every rate; 5/6 is the exception the
    reference builds in: its decoder runs 6799 of 6826 steps per call, so only the stretch it does decode is compared"""
    rng = np.random.default_rng(1)
    for rate in range(5):
        bits = rng.integers(0, 2, 60000, dtype=np.uint8)
        s = dvbs_stream.inner_softs(bits, rate, rng, sigma=12.0)
        s = s[:len(s) // 8192 * 8192]
        v = OrcViterbi()
        out = v.process(s)
        st = v.stats()
        assert st[1] == 1 and st[2] == rate and st[3] == 0, (rate, st)
        if rate != 3:
            ok = np.mean(out[100:len(out) - 100] == bits[100:len(out) - 100])
            assert ok > 0.995, (rate, ok)
        else:
            assert np.mean(out[200:6000] == bits[200:6000]) > 0.995


needs_ref = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_vit_create"),
                               reason="oracle/_ref/libdvbs2_ref.so (with the Viterbi decoder) not built")

CASES = [  # rate, sigma, phase, lead, seed
    (0, 10.0, 0, 0, 0), (0, 25.0, 1, 1, 1), (1, 10.0, 0, 0, 2), (1, 18.0, 1, 2, 3), (1, 10.0, 0, 4, 4), (2, 12.0, 0, 0, 5),
    (2, 10.0, 1, 2, 6), (3, 8.0, 0, 0, 7), (3, 8.0, 0, 7, 8), (3, 8.0, 1, 13, 9), (4, 6.0, 0, 0, 10), (4, 6.0, 1, 6, 11)]


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_viterbi_matches_reference(case):
    """lock search (52 candidates per call until one locks), decoding at the lock, BER and the lock parameters after every
    call; calls of one to five blocks; the puncturing phase moved by `lead` soft bits, the constellation by 90 degrees"""
    rate, sigma, phase, lead, seed = case
    rng = np.random.default_rng(100 + seed)
    bits = rng.integers(0, 2, 130000, dtype=np.uint8)
    s = dvbs_stream.inner_softs(bits, rate, rng, sigma=sigma, phase=phase, lead=lead)
    nb = min(len(s) // 8192, 14)
    a, b = OrcViterbi(), RefViterbi()
    k = 0
    while k < nb:
        n = min(int(rng.integers(1, 6)), nb - k)
        seg = s[k * 8192:(k + n) * 8192]
        oa, ob = a.process(seg, fill=k & 1), b.process(seg, fill=k & 1)
        assert oa.shape == ob.shape and np.array_equal(oa, ob), (k, n)
        assert same_stats(a.stats(), b.stats()), (k, a.stats(), b.stats())
        k += n
    assert a.stats()[1] == 1 and a.stats()[2] == rate


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_viterbi_lock_lost_and_found_matches_reference(seed):
    """noise only (the search runs every call), then a signal, then noise until the lock is given up after max_outsync bad
    calls, then another rate: the chained start states of the ten decoders, the re-encoders' registers and the
    depuncturers' state carry over all of it"""
    rng = np.random.default_rng(200 + seed)
    r1, r2 = [(1, 3), (4, 0), (2, 1)][seed]
    noise = lambda n: np.clip(np.rint(rng.normal(0, 40, n * 8192)), -128, 127).astype(np.int8)
    sig = lambda r, n, lead: dvbs_stream.inner_softs(rng.integers(0, 2, 8192 * n, dtype=np.uint8), r, rng, sigma=[12.0, 10.0, 9.0, 7.0, 5.0][r],
                                                 lead=lead)[:n * 8192]
    s = np.concatenate([noise(2), sig(r1, 4, 2), noise(5), sig(r2, 4, 0), np.zeros(8192, np.int8), np.full(8192, -128, np.int8)])
    a, b = OrcViterbi(0.15, 3), RefViterbi(0.15, 3)
    seen = set()
    for k in range(len(s) // 8192):
        seg = s[k * 8192:(k + 1) * 8192]
        oa, ob = a.process(seg, fill=7), b.process(seg, fill=7)
        assert np.array_equal(oa, ob), k
        assert same_stats(a.stats(), b.stats()), (k, a.stats(), b.stats())
        seen.add(a.stats()[1:3] if a.stats()[1] else (0, -1))
    assert (1, r1) in seen and (1, r2) in seen and (0, -1) in seen


@needs_ref
def test_syms_to_soft_matches_reference():
    rng = np.random.default_rng(5)
    o, r = orclib.oracle(), orclib.ref()
    ho, hr = o.orc_sts_create(), r.ref_sts_create()
    for n in [1, 4095, 1, 5000, 0, 12289, 3]:
        syms = (rng.normal(0, 0.9, (n, 2))).astype(np.float32)
        if n > 10:
            syms[3] = [1.27, -1.27]
            syms[4] = [1.2701, -1.2799]
            syms[5] = [5.0, -7.0]
            syms[6] = [0.0049, -0.0099]
        oa, ob = np.zeros(2 * n + 8192, np.int8), np.zeros(2 * n + 8192, np.int8)
        na = o.orc_sts_process(ho, n, np.ascontiguousarray(syms.reshape(-1)), oa)
        nb = r.ref_sts_process(hr, n, np.ascontiguousarray(syms.reshape(-1)), ob)
        assert na == nb and np.array_equal(oa, ob)
