"""Row 8(f)-3 oracle (oracle/oracle_plsync.c) against the reference's own S2PLSyncBlock / S2PLHDRDemod /
dvbs2_pilot_coarse_fed, compiled unmodified into oracle/_ref (SDR++ core and VOLK replaced by oracle/shim/)."""
import ctypes as C

import numpy as np
import pytest

import orclib
import plstream

needs_ref = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_plsync_create"),
                               reason="oracle/_ref/libdvbs2_ref.so (with PL sync) not built")


def f32(x):
    return np.ascontiguousarray(x).view(np.float32)


class OrcSync:
    def __init__(self, slots, pilots):
        self.o = orclib.oracle()
        self.h = self.o.orc_plsync_create(slots, int(pilots))
        self.rfs = self.o.orc_raw_frame_size(slots, int(pilots))

    def process(self, x):
        out = np.zeros(2 * (len(x) + 2 * self.rfs), np.float32)
        n = self.o.orc_plsync_process(self.h, len(x), f32(x), out)
        return out[:2 * n].view(np.complex64).copy()

    def stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_double()
        self.o.orc_plsync_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class RefSync(OrcSync):
    def __init__(self, slots, pilots):
        self.o = orclib.ref()
        self.h = self.o.ref_plsync_create(slots, int(pilots))
        self.rfs = orclib.oracle().orc_raw_frame_size(slots, int(pilots))

    def process(self, x):
        out = np.zeros(2 * (len(x) + 2 * self.rfs), np.float32)
        n = self.o.ref_plsync_process(self.h, len(x), f32(x), out)
        return out[:2 * n].view(np.complex64).copy()

    def stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_double()
        self.o.ref_plsync_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


def test_tables_and_frame_sizes():
    o = orclib.oracle()
    assert o.orc_raw_frame_size(360, 0) == 32490 and o.orc_raw_frame_size(360, 1) == 32490 + 22 * 36
    assert o.orc_raw_frame_size(90, 1) == 8190 + 5 * 36 and o.orc_raw_frame_size(36, 1) == 3330 + 2 * 36
    h = plstream.plheader(4 << 2)
    assert np.allclose(np.abs(h), 1.0, atol=1e-6)
    # pi/2-BPSK: consecutive header symbols are 90 degrees apart
    assert np.allclose(np.abs(np.angle(h[1:] * np.conj(h[:-1]))), np.pi / 2, atol=1e-5)


def test_sync_locks_and_delivers_aligned_frames():
    rng = np.random.default_rng(3)
    slots, lead = 90, 1234
    x = plstream.stream(4 << 2, slots, False, 6, rng, lead=lead)
    s = OrcSync(slots, False)
    y = s.process(x)
    rfs = s.rfs
    assert len(y) % rfs == 0 and len(y) >= 4 * rfs
    fr = y.reshape(-1, rfs)
    # once locked every delivered frame is a verbatim slice of the input starting at a PLHEADER
    assert np.array_equal(fr[1], x[lead + rfs: lead + 2 * rfs]) or np.array_equal(fr[1], x[lead: lead + rfs])
    assert s.stats()[2] > 0.6


@needs_ref
@pytest.mark.parametrize("case", [(90, False, 16.0, 0), (90, True, 8.0, 1), (36, False, 3.0, 2), (144, True, 1.0, 3), (360, False, 6.0, 4)])
def test_sync_matches_reference(case):
    slots, pilots, esn0, seed = case
    rng = np.random.default_rng(40 + seed)
    x = plstream.stream(11 << 2 | pilots, slots, pilots, 5, rng, esn0_db=esn0, lead=int(rng.integers(1, 3000)), cfo=2e-4 * seed)
    o, r = OrcSync(slots, pilots), RefSync(slots, pilots)
    cuts = sorted(set(int(c) for c in rng.integers(0, len(x), 7)) | {0, len(x)})
    for a, b in zip(cuts[:-1], cuts[1:]):
        ya, yb = o.process(x[a:b]), r.process(x[a:b])
        assert len(ya) == len(yb) and np.array_equal(ya.view(np.uint32), yb.view(np.uint32))
        sa, sb = o.stats(), r.stats()
        assert sa[:2] == sb[:2] and sa[2] == sb[2]


@needs_ref
def test_sync_on_noise_matches_reference():
    """no PLHEADER anywhere: positions are picked on noise maxima; the oracle follows the reference through every
    realignment"""
    rng = np.random.default_rng(9)
    x = ((rng.normal(size=40000) + 1j * rng.normal(size=40000)) / np.sqrt(2)).astype(np.complex64)
    o, r = OrcSync(36, False), RefSync(36, False)
    for a in range(0, len(x), 5000):
        ya, yb = o.process(x[a:a + 5000]), r.process(x[a:a + 5000])
        assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32))
        assert o.stats() == r.stats()


@needs_ref
@pytest.mark.parametrize("pls", [4 << 2, (13 << 2) | 1, (28 << 2) | 2, (6 << 2) | 3])
def test_plhdr_demod_matches_reference(pls):
    rng = np.random.default_rng(pls)
    slots, pilots = 90, bool(pls & 1)
    x = plstream.stream(pls, slots, pilots, 8, rng, esn0_db=7.0, lead=0, cfo=3e-5, phase=0.4)
    rfs = orclib.oracle().orc_raw_frame_size(slots, int(pilots))
    o, r = orclib.oracle(), orclib.ref()
    ho, hr = o.orc_plhdr_create(0.004), r.ref_plhdr_create(0.004)
    for k in range(8):
        fr = f32(x[k * rfs:(k + 1) * rfs])
        oa, ob = np.zeros(180, np.float32), np.zeros(180, np.float32)
        ra, rb = np.zeros(3, np.int32), np.zeros(3, np.int32)
        la, lb = np.zeros(2, np.float32), np.zeros(2, np.float32)
        o.orc_plhdr_process(ho, rfs, fr, oa, ra, la)
        r.ref_plhdr_process(hr, rfs, fr, ob, rb, lb)
        assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
        assert list(ra) == list(rb) and np.array_equal(la.view(np.uint32), lb.view(np.uint32))
    o.orc_plhdr_destroy(ho)


@pytest.mark.parametrize("pls", [4 << 2, (13 << 2) | 1, (28 << 2) | 2, (6 << 2) | 3])
def test_plhdr_decodes_the_pls_code_on_a_clean_carrier(pls):
    """no carrier offset: MODCOD, frame size and pilots come back (with an offset the reference's loop may settle
    half a turn away, which inverts the frame-size bit -- reference behaviour, covered by the parity test above)"""
    rng = np.random.default_rng(pls)
    pilots = bool(pls & 1)
    x = plstream.stream(pls, 90, pilots, 4, rng, esn0_db=12.0, lead=0, cfo=0.0, phase=0.0)
    o = orclib.oracle()
    rfs = o.orc_raw_frame_size(90, int(pilots))
    h = o.orc_plhdr_create(0.004)
    res, loop, out = np.zeros(3, np.int32), np.zeros(2, np.float32), np.zeros(180, np.float32)
    for k in range(4):
        o.orc_plhdr_process(h, rfs, f32(x[k * rfs:(k + 1) * rfs]), out, res, loop)
        assert list(res) == [pls >> 2, (pls >> 1) & 1, pls & 1]
    o.orc_plhdr_destroy(h)


@needs_ref
@pytest.mark.parametrize("case", [(90, False, 0), (90, True, 1), (360, True, 2), (36, True, 3)])
def test_coarse_fed_matches_reference(case):
    slots, pilots, seed = case
    rng = np.random.default_rng(70 + seed)
    pls = (5 << 2) | int(pilots)
    codenum = [0, 1, 7, 262141][seed]
    cfo = 1e-3 * (seed + 1)
    x = plstream.stream(pls, slots, pilots, 3, rng, esn0_db=10.0, lead=0, cfo=cfo, codenum=codenum)
    rfs = orclib.oracle().orc_raw_frame_size(slots, int(pilots))
    rn = plstream.pl_rn(codenum)
    for k in range(3):
        fr = f32(x[k * rfs:(k + 1) * rfs])
        a = orclib.oracle().orc_coarse_fed(fr, rfs, int(pilots), pls, rn)
        b = orclib.ref().ref_coarse_fed(fr, rfs, int(pilots), pls, codenum)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        assert a > 0   # a positive carrier offset reads as a positive error
