"""Row 8(f)-4, byte-domain half, on the GPU: dvbs2fec_dvbs_outer_* (deinterleaver, RS(204,188), descrambler of the DVB-S
chain) against the CPU oracle, which tests/test_dvbs_oracle.py pins to the reference's own classes over the vendored
libcorrect.  Byte work: bit-exact."""
import numpy as np
import pytest

import dvbs_stream
from fec import pkg
from test_dvbs_oracle import OrcOuter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [((0, 0), 1632, 0), ((0, 8), 1632, 1), ((0, 12), 1632, 2), ((6, 20), 1632, 3), ((0, 9), 204, 4),
                                  ((100, 204), 1632, 5), ((7, 9), 1632, 6)])
def test_outer_decoder_matches_oracle(case):
    """clean, correctable, mixed and hopeless packets (the previous packet's bytes come out, or libcorrect miscorrects),
    the module's own frame stride of 204 bytes; FIFO contents, the decoder's output buffer and the descrambler register
    carried over calls of uneven sizes"""
    span, stride, seed = case
    rng = np.random.default_rng(10 + seed)
    ts, ch = dvbs_stream.outer_stream(24, rng)
    bad = dvbs_stream.add_errors(ch, rng, per_packet=span)
    if seed == 1:
        bad[203::204 * 3] ^= 0x5A            # the last parity byte: libcorrect's position 255
    o, g = OrcOuter(), pkg.DVBSOuterDecoder()
    nfr = 24 if stride == 1632 else (len(bad) - 1632) // stride + 1
    cuts = [0, 1, 2, 9, nfr] if stride == 1632 else [0, 3, 40, nfr]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        seg = bad[lo * stride:]
        want, we = o.process(seg, hi - lo, stride)
        got, ge = g.process(seg, hi - lo, stride)
        assert np.array_equal(ge, we)
        assert np.array_equal(got, want)
    g.reset()
    o2 = OrcOuter()
    want, we = o2.process(bad, 3)
    got, ge = g.process(bad, 3)
    assert np.array_equal(got, want) and np.array_equal(ge, we)
    g.close()


def test_outer_decoder_large_batch_round_trip_and_oracle():
    """2048 frames (16384 packets, 3.3 MB) in one call, up to 8 byte errors per codeword: every TS packet comes back, eleven
    packets late; equal to the oracle"""
    rng = np.random.default_rng(77)
    ts, bad = dvbs_stream.outer_stream(2048, rng, codeword_errors=(0, 8))
    g = pkg.DVBSOuterDecoder()
    got, ge = g.process(bad, 2048)
    assert np.array_equal(got[24:], ts[13:len(ts) - 11])
    assert ge[24:].max() <= 8
    want, we = OrcOuter().process(bad, 2048)
    assert np.array_equal(got, want) and np.array_equal(ge, we)
    g.close()


def test_outer_decoder_argument_errors():
    g = pkg.DVBSOuterDecoder()
    with pytest.raises(ValueError):
        g.process(np.zeros(1000, np.uint8), 1)
    out, err = g.process(np.zeros(0, np.uint8), 0)
    assert out.shape == (0, 188)
    with pytest.raises(pkg.DVBS2FecError):
        pkg.DVBSOuterDecoder(device=99)
    g.close()


# ---- K10: TS deframer ------------------------------------------------------------------------------------------------
from test_dvbs_oracle import OrcDeframer, deframer_bits


@pytest.mark.parametrize("case", [(0.0, False, 0), (0.02, False, 1), (0.02, True, 2), (0.5, False, 3), (0.004, False, 4)])
def test_deframer_matches_oracle(case):
    """clean, bit errors (sync bytes with a few wrong bits still lock), inverted stream, noise only; the stream cut into
    calls at odd places (the window crosses calls, down to calls of one bit)"""
    flip, invert, seed = case
    rng = np.random.default_rng(30 + seed)
    frames, bits = deframer_bits(5, rng, lead=int(rng.integers(1, 3000)), flip=flip, invert=invert)
    o, g = OrcDeframer(), pkg.DVBSTSDeframer()
    cuts = sorted(set(int(c) for c in rng.integers(0, len(bits), 6)) | {0, 1, 2, len(bits)})
    total = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        want, ws = o.work(bits[lo:hi])
        got = g.work(bits[lo:hi])
        assert got.shape == want.shape and np.array_equal(got, want)
        st = g.stats()
        assert st[:2] == ws and st[2] == len(want)
        total += len(got)
    if flip == 0.0:
        assert total == 5
    g.reset()
    want, _ = OrcDeframer().work(bits[:20000])
    assert np.array_equal(g.work(bits[:20000]), want)
    g.close()


def test_deframer_large_call_and_frame_limit():
    """300 frames (3.9 Mbit) in one call: all found, in order; max_frames bounds what is written, stats tell what was found"""
    rng = np.random.default_rng(5)
    frames, bits = deframer_bits(300, rng, lead=12345)
    g = pkg.DVBSTSDeframer()
    got = g.work(bits, max_frames=400)
    assert np.array_equal(got, frames)
    g.reset()
    got = g.work(bits, max_frames=7)
    assert np.array_equal(got, frames[:7]) and g.stats()[2] == 300
    assert g.work(np.zeros(0, np.uint8)).shape == (0, 1632)
    with pytest.raises(pkg.DVBS2FecError):
        pkg.DVBSTSDeframer(device=99)
    g.close()
