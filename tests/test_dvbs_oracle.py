"""Row 8(f)-4, byte-domain half: the oracle's DVB-S outer decoder (oracle/oracle_dvbs.c: deinterleaver, RS(204,188),
descrambler) against the reference's own DVBSInterleaving / DVBSReedSolomon (vendored libcorrect) / DVBSScrambling driven
as DVBSDemod::process drives them (dvbs/module_dvbs_demod.cpp:91-106), compiled unmodified into oracle/_ref."""
import numpy as np
import pytest

import dvbs_stream
import orclib

needs_ref = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_dvbs_outer_create"),
                               reason="oracle/_ref/libdvbs2_ref.so (with the DVB-S outer decoder) not built")


class OrcOuter:
    def __init__(self):
        self.o = orclib.oracle()
        self.h = self.o.orc_dvbs_outer_create()

    def process(self, buf, nframes, stride=1632):
        buf = np.ascontiguousarray(buf, np.uint8)
        assert len(buf) >= (nframes - 1) * stride + 1632
        out = np.zeros(nframes * 1504, np.uint8)
        err = np.zeros(nframes * 8, np.int32)
        self.o.orc_dvbs_outer_process(self.h, buf, nframes, stride, out, err)
        return out.reshape(-1, 188), err


class RefOuter(OrcOuter):
    def __init__(self):
        self.o = orclib.ref()
        self.h = self.o.ref_dvbs_outer_create()

    def process(self, buf, nframes, stride=1632):
        buf = np.ascontiguousarray(buf, np.uint8)
        out = np.zeros(nframes * 1504, np.uint8)
        err = np.zeros(nframes * 8, np.int32)
        self.o.ref_dvbs_outer_process(self.h, buf, nframes, stride, out, err)
        return out.reshape(-1, 188), err


def test_clean_stream_round_trip():
    """transmit side written from EN 300 421 -> the oracle's decoder: the TS packets come back, eleven packets late
    (the deinterleaver's delay), error counts zero"""
    rng = np.random.default_rng(1)
    ts, ch = dvbs_stream.outer_stream(6, rng)
    out, err = OrcOuter().process(ch, 6)
    # the first 11 packets are the deinterleaver filling up; the descrambler locks at the first inverted sync it sees
    first = 16
    assert np.array_equal(out[first:], ts[first - 11:len(ts) - 11])
    assert (err[first:] == 0).all()


def test_correctable_errors_are_corrected_and_counted():
    rng = np.random.default_rng(2)
    ts, bad = dvbs_stream.outer_stream(8, rng, codeword_errors=(0, 8))
    out, err = OrcOuter().process(bad, 8)
    assert np.array_equal(out[24:], ts[13:len(ts) - 11])
    assert err[24:].max() <= 8 and err[24:].sum() > 0


@needs_ref
def test_rs_parity_matches_libcorrect():
    rng = np.random.default_rng(3)
    a, b = np.zeros(16, np.uint8), np.zeros(16, np.uint8)
    for _ in range(50):
        m = rng.integers(0, 256, 188, dtype=np.uint8)
        orclib.oracle().orc_rs204_parity(m, a)
        orclib.ref().ref_rs204_parity(m, b)
        assert np.array_equal(a, b)


@needs_ref
@pytest.mark.parametrize("case", [((0, 0), 1632, 0), ((0, 8), 1632, 1), ((0, 12), 1632, 2), ((6, 20), 1632, 3), ((0, 9), 204, 4),
                                  ((100, 204), 1632, 5)])
def test_outer_decoder_matches_reference(case):
    """clean, correctable, mixed and hopeless packets (libcorrect gives up: the previous packet's bytes come out and the
    difference is 'counted'; or it miscorrects), the module's own frame stride of 204 bytes, noise only; state carried
    over calls (FIFOs, the decoder's output buffer, the descrambler register)"""
    span, stride, seed = case
    rng = np.random.default_rng(10 + seed)
    ts, ch = dvbs_stream.outer_stream(12, rng)
    bad = dvbs_stream.add_errors(ch, rng, per_packet=span)
    if seed == 1:      # errors in the last parity byte of some packets: libcorrect's position 255 (log[1])
        bad[203::204 * 3] ^= 0x5A
    a, b = OrcOuter(), RefOuter()
    nfr = 12 if stride == 1632 else (len(bad) - 1632) // stride + 1
    cuts = [0, 1, 4, nfr] if stride == 1632 else [0, 3, 40, nfr]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        seg = bad[lo * stride:]
        oa, ea = a.process(seg, hi - lo, stride)
        ob, eb = b.process(seg, hi - lo, stride)
        assert np.array_equal(ea, eb)
        assert np.array_equal(oa, ob)


# ---- TS deframer ----------------------------------------------------------------------------------------------------
class OrcDeframer:
    def __init__(self):
        self.o = orclib.oracle()
        self.h = self.o.orc_dvbs_deframer_create()
        self.work_fn, self.stats_fn = self.o.orc_dvbs_deframer_work, self.o.orc_dvbs_deframer_stats

    def work(self, bits):
        import ctypes as C
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros((len(bits) // 100 + 4) * 1632, np.uint8)      # room for the frames a test stream can hold
        n = self.work_fn(self.h, bits, len(bits), out)
        a, b = C.c_int(), C.c_int()
        self.stats_fn(self.h, C.byref(a), C.byref(b))
        return out[:n * 1632].reshape(n, 1632).copy(), (a.value, b.value)


class RefDeframer(OrcDeframer):
    def __init__(self):
        self.o = orclib.ref()
        self.h = self.o.ref_dvbs_deframer_create()
        self.work_fn, self.stats_fn = self.o.ref_dvbs_deframer_work, self.o.ref_dvbs_deframer_stats


def deframer_bits(nframes, rng, lead=777, flip=0.0, invert=False):
    """the unpacked bit stream behind the Viterbi decoder: junk, then frames of 8 x 204 bytes with their sync bytes"""
    ts, ch = dvbs_stream.outer_stream(nframes, rng)
    bits = np.unpackbits(ch)
    bits = np.concatenate([rng.integers(0, 2, lead, dtype=np.uint8), bits, rng.integers(0, 2, 300, dtype=np.uint8)])
    if flip:
        bits ^= (rng.random(len(bits)) < flip).astype(np.uint8)
    if invert:
        bits ^= 1
    return ch.reshape(nframes, 1632), bits


def test_deframer_finds_the_frames():
    rng = np.random.default_rng(4)
    frames, bits = deframer_bits(3, rng)
    got, st = OrcDeframer().work(bits)
    assert len(got) == 3 and np.array_equal(got, frames) and st == (0, 0)
    frames, bits = deframer_bits(2, rng, invert=True)
    got, st = OrcDeframer().work(bits)
    assert len(got) == 2 and np.array_equal(got, frames)


needs_ref_def = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_dvbs_deframer_create"),
                                   reason="oracle/_ref/libdvbs2_ref.so (with the TS deframer) not built")


@needs_ref_def
@pytest.mark.parametrize("case", [(0.0, False, 0), (0.02, False, 1), (0.02, True, 2), (0.5, False, 3)])
def test_deframer_matches_reference(case):
    """clean, 2 % bit errors (sync bytes with a few wrong bits still lock, some frames are missed), inverted stream, noise
    only; the stream cut into calls at odd places (the window crosses calls)"""
    flip, invert, seed = case
    rng = np.random.default_rng(20 + seed)
    frames, bits = deframer_bits(3, rng, lead=int(rng.integers(1, 2000)), flip=flip, invert=invert)
    a, b = OrcDeframer(), RefDeframer()
    cuts = sorted(set(int(c) for c in rng.integers(0, len(bits), 5)) | {0, len(bits)})
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ga, sa = a.work(bits[lo:hi])
        gb, sb = b.work(bits[lo:hi])
        assert ga.shape == gb.shape and np.array_equal(ga, gb) and sa == sb
