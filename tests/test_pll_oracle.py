"""Row 8(f)-2 oracle (oracle/oracle_pll.c) against the reference's own S2PLLBlock (dvbs2/dvbs2_pll.cpp), compiled
unmodified into oracle/_ref (SDR++ core's PhaseControlLoop / phasor / complex_t replaced by oracle/shim/): output
symbols and loop state bit for bit, frame after frame (the loop state carries over)."""
import numpy as np
import pytest

import orclib
import plstream

needs_ref = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_pll_create"),
                               reason="oracle/_ref/libdvbs2_ref.so (with the PLL) not built")

# (constellation type as dsp::constellation_type_t, bits, g1, g2)
CONST = {"qpsk": (1, 2, 0.0, 0.0), "8psk": (3, 3, 0.0, 0.0), "16apsk": (4, 4, 3.15, 0.0), "32apsk": (5, 5, 2.53, 4.30)}


def frames_for(name, slots, pilots, nframes, rng, esn0_db, cfo, codenum, modcod=4, short=False):
    """nframes aligned PLFRAMEs (as S2PLSyncBlock delivers them) with a residual carrier offset and noise"""
    ctype, bits, g1, g2 = CONST[name]
    pls = (modcod << 2) | int(short) << 1 | int(pilots)
    rfs = orclib.oracle().orc_raw_frame_size(slots, int(pilots))
    x = plstream.stream(pls, slots, pilots, nframes, rng, esn0_db=esn0_db, lead=0, cfo=cfo, phase=0.2, codenum=codenum,
                        bits=min(bits, 3))
    return pls, x.reshape(nframes, rfs)


class OrcPll:
    def __init__(self, bw, name, slots, pilots, pls, codenum):
        ctype, _, g1, g2 = CONST[name]
        self.o = orclib.oracle()
        self.h = self.o.orc_pll_create(bw, ctype, g1, g2, slots, int(pilots), pls, codenum)
        self.total = (slots + 1) * 90 + self.o.orc_pll_pilot_cnt(self.h) * 36

    def process(self, frame):
        out = np.zeros(2 * len(frame), np.float32)
        st = np.zeros(3, np.float32)
        n = self.o.orc_pll_process(self.h, np.ascontiguousarray(frame).view(np.float32), out, st)
        return out[:2 * n].view(np.complex64).copy(), st


class RefPll(OrcPll):
    def __init__(self, bw, name, slots, pilots, pls, codenum):
        ctype, _, g1, g2 = CONST[name]
        self.o = orclib.ref()
        self.h = self.o.ref_pll_create(bw, ctype, g1, g2, slots, int(pilots), pls, codenum)
        self.total = (slots + 1) * 90 + self.o.ref_pll_pilot_cnt(self.h) * 36

    def process(self, frame):
        out = np.zeros(2 * len(frame), np.float32)
        st = np.zeros(3, np.float32)
        n = self.o.ref_pll_process(self.h, len(frame), np.ascontiguousarray(frame).view(np.float32), out, st)
        return out[:2 * n].view(np.complex64).copy(), st


def test_pilot_count_follows_the_reference_rule():
    o = orclib.oracle()
    for slots, want in ((360, 1), (90, 1), (36, 1), (1620, 1), (3060, 2)):
        h = o.orc_pll_create(0.005, 1, 0, 0, slots, 1, 17, 0)
        assert o.orc_pll_pilot_cnt(h) == want     # counted from the slot number (dvbs2_pll.h:50-58, SURVEY note N2)
        o.orc_pll_destroy(h)


def test_loop_locks_and_derotates():
    rng = np.random.default_rng(5)
    pls, fr = frames_for("qpsk", 90, False, 4, rng, 14.0, 2e-5, 0)
    p = OrcPll(0.005, "qpsk", 90, False, pls, 0)
    for f in fr:
        y, st = p.process(f)
    pay = y[90:]
    # locked: the descrambled payload sits on the QPSK points
    ang = np.angle(pay * np.exp(-1j * np.pi / 4)) % (np.pi / 2)
    dev = np.minimum(ang, np.pi / 2 - ang)
    assert np.median(dev) < 0.15
    assert abs(st[1] - 2 * np.pi * 2e-5) < 5e-5     # the loop frequency found the carrier offset


@needs_ref
@pytest.mark.parametrize("case", [("qpsk", 90, False, 10.0, 3e-5, 0, 0), ("qpsk", 90, True, 6.0, -2e-5, 1, 1),
                                  ("8psk", 60, False, 12.0, 1e-5, 0, 2), ("16apsk", 45, True, 16.0, 2e-5, 7, 3),
                                  ("32apsk", 36, False, 20.0, 1e-5, 0, 4), ("qpsk", 360, True, 3.0, 4e-5, 0, 5)])
def test_pll_matches_reference(case):
    name, slots, pilots, esn0, cfo, codenum, seed = case
    rng = np.random.default_rng(100 + seed)
    pls, fr = frames_for(name, slots, pilots, 3, rng, esn0, cfo, codenum)
    a = OrcPll(0.004, name, slots, pilots, pls, codenum)
    b = RefPll(0.004, name, slots, pilots, pls, codenum)
    assert a.total == b.total
    for f in fr:
        ya, sa = a.process(f)
        yb, sb = b.process(f)
        assert len(ya) == len(yb) == a.total
        assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32))
        assert np.array_equal(sa.view(np.uint32), sb.view(np.uint32))


def test_phase_error_table_is_a_phase_detector():
    """the oracle's 256x256 table: finite, and inside a QPSK decision region the error is the angle to the point"""
    o = orclib.oracle()
    c = o.orc_const_create(1, 0, 0)
    lut = np.ctypeslib.as_array(o.orc_const_phase_lut(c), shape=(256, 256)).copy()
    assert np.isfinite(lut).all() and np.abs(lut).max() <= np.float32(np.pi)
    x = (200 - 128) / 256 * 1.5          # a sample at 30 degrees: 15 degrees short of the point at 45
    y = x * np.tan(np.pi / 6)
    e = o.orc_demod_phase_error(c, x, y)
    assert abs(e - (-np.pi / 12)) < 0.02
    o.orc_const_destroy(c)
