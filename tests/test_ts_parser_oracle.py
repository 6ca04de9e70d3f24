"""Row 8(f)-1 (BBFRAME -> TS / GSE): the C restatement oracle/oracle_ts.c against the compiled reference
parser (dvbs2/bbframe_ts_parser.cpp in oracle/_ref), plus properties that hold without the reference."""
import ctypes as C

import numpy as np
import pytest

import bbstream
import orclib

KBCH = {"n1/2": 32208, "s1/4": 3072, "n9/10": 58192, "s8/9": 14232}


class OrcParser:
    def __init__(self, kbch):
        self.o = orclib.oracle()
        self.h = self.o.orc_ts_create(kbch)

    def work(self, frames, cap=65536 * 10):
        out = np.zeros(cap + 4096, np.uint8)
        n = self.o.orc_ts_work(self.h, np.ascontiguousarray(frames), len(frames), out, cap)
        return out[:n].copy()

    def stats(self):
        hdr = np.zeros(10, np.uint8)
        v = [C.c_int() for _ in range(6)]
        self.o.orc_ts_stats(self.h, hdr, *[C.byref(x) for x in v])
        have, cnt, proc, gse_err, synched, pending = [x.value for x in v]
        return dict(hdr=hdr, have=have, cnt=cnt, proc=proc, gse_err=gse_err, synched=synched, pending=pending)

    def close(self):
        self.o.orc_ts_destroy(self.h)


class RefParser:
    def __init__(self, kbch):
        self.r = orclib.ref()
        self.h = self.r.ref_ts_create(kbch)

    def work(self, frames, cap=65536 * 10):
        out = np.zeros(cap + 4096, np.uint8)
        n = self.r.ref_ts_work(self.h, np.ascontiguousarray(frames).copy(), len(frames), out, cap)
        return out[:n].copy()

    def stats(self):
        f = np.zeros(11, np.int32)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.r.ref_ts_stats(self.h, f, C.byref(a), C.byref(b), C.byref(c))
        return dict(fields=f, cnt=a.value, proc=b.value, gse_err=c.value)

    def close(self):
        pass  # keep alive like the module does


def header_fields(h):
    h = [int(x) for x in h]
    sis = (h[0] >> 5) & 1
    return [h[0] >> 6, sis, (h[0] >> 4) & 1, (h[0] >> 3) & 1, (h[0] >> 2) & 1, h[0] & 3, h[1] if sis == 0 else 0,
            (h[2] << 8) | h[3], (h[4] << 8) | h[5], h[6], (h[7] << 8) | h[8]]


def test_ts_packets_come_back_in_order_with_sync_restored():
    rng = np.random.default_rng(1)
    kbch = KBCH["n1/2"]
    pk = bbstream.ts_packets(400, rng)
    frames, used = bbstream.ts_bbframes(kbch, pk, first_byte=77)
    p = OrcParser(kbch)
    out = p.work(frames)
    assert len(out) % 188 == 0 and len(out) > 0
    got = out.reshape(-1, 188)
    # resync enters just past the first sync byte: packet 1 is the first complete one after stream byte 77
    assert np.array_equal(got, pk[1:1 + len(got)])
    st = p.stats()
    assert st["proc"] == st["cnt"] == len(frames) and st["synched"] == 1
    assert len(got) == (used - 188 - 1) // 188 or len(got) == (used - 189) // 188 + 0
    p.close()


def test_ts_state_carries_across_calls_and_resyncs_after_a_bad_header():
    rng = np.random.default_rng(2)
    kbch = KBCH["s1/4"]
    pk = bbstream.ts_packets(300, rng)
    frames, _ = bbstream.ts_bbframes(kbch, pk)
    one = OrcParser(kbch)
    whole = one.work(frames)
    split = OrcParser(kbch)
    parts = [split.work(frames[a:b]) for a, b in ((0, 5), (5, 6), (6, 40), (40, len(frames)))]
    assert np.array_equal(np.concatenate(parts), whole)
    # break frame 9's header: its packets and the partial ones around it are lost, order is kept
    bad = frames.copy()
    bad[9, 3] ^= 0x10
    out = OrcParser(kbch).work(bad).reshape(-1, 188)
    ids = [int(np.flatnonzero((pk == row).all(axis=1))[0]) for row in out]
    assert ids == sorted(ids) and len(set(ids)) == len(ids)
    assert len(ids) < len(whole) // 188
    one.close(); split.close()


needs_ref = pytest.mark.skipif(not orclib.have_ref() or not hasattr(orclib.ref(), "ref_ts_create"),
                               reason="oracle/_ref/libdvbs2_ref.so (with the TS parser) not built")


def _compare(kbch, batches, cap=65536 * 10):
    o, r = OrcParser(kbch), RefParser(kbch)
    for frames in batches:
        a, b = o.work(frames, cap), r.work(frames, cap)
        assert np.array_equal(a, b)
        so, sr = o.stats(), r.stats()
        assert (so["cnt"], so["proc"], so["gse_err"]) == (sr["cnt"], sr["proc"], sr["gse_err"])
        if so["have"]:
            assert header_fields(so["hdr"]) == [int(x) for x in sr["fields"]]
    o.close()


@needs_ref
@pytest.mark.parametrize("name", list(KBCH))
def test_ts_matches_reference_parser(name):
    kbch = KBCH[name]
    rng = np.random.default_rng(hash(name) % 1000)
    pk = bbstream.ts_packets(max(300, 12 * kbch // 1504), rng)
    frames, _ = bbstream.ts_bbframes(kbch, pk, first_byte=int(rng.integers(0, 188)))
    cuts = sorted(set(int(x) for x in rng.integers(1, len(frames), 4)))
    batches = [frames[a:b] for a, b in zip([0] + cuts, cuts + [len(frames)])]
    _compare(kbch, batches)


@needs_ref
def test_ts_odd_streams_match_reference_parser():
    """short data fields (partial unit never completed), header faults of every kind, non-TS frames in between"""
    kbch = KBCH["s1/4"]
    f = bbstream.odd_ts_scenario(np.random.default_rng(5), kbch)
    _compare(kbch, [f[:60], f[60:61], f[61:]])


@needs_ref
def test_ts_output_room_rule_matches_reference_parser():
    """fewer than 189 bytes of room left: the call stops early and drops the rest (first call only compared --
    the reference overruns its 188-byte carry buffer at that point, its later state is undefined)"""
    rng = np.random.default_rng(6)
    kbch = KBCH["n1/2"]
    pk = bbstream.ts_packets(200, rng)
    frames, _ = bbstream.ts_bbframes(kbch, pk)
    for cap in (188, 189, 190, 188 * 7 + 5, 188 * 21, 188 * 21 + 1, 188 * 22):
        _compare(kbch, [frames[:3]], cap=cap)


def test_gse_pdus_are_reassembled_and_wrapped():
    rng = np.random.default_rng(8)
    kbch = KBCH["n1/2"]
    frames = bbstream.gse_bbframes(kbch, bbstream.gse_scenario(rng))
    p = OrcParser(kbch)
    out = p.work(frames)
    # 4 complete PDUs + 3 good reassemblies; IPv4/IPv6 carry a 4-byte GRE header, others 2 bytes
    sizes = [40, 64, 1200, 10, 2000, 3000, 500]
    heads = [4, 4, 4, 2, 2, 4, 4]
    assert len(out) == sum(sizes) + sum(heads)
    assert p.stats()["proc"] == 5
    p.close()


@needs_ref
def test_gse_matches_reference_parser():
    rng = np.random.default_rng(8)
    for name in ("n1/2", "n9/10"):
        kbch = KBCH[name]
        frames = bbstream.gse_bbframes(kbch, bbstream.gse_scenario(rng))
        _compare(kbch, [frames[:2], frames[2:]])
        _compare(kbch, [frames])


@needs_ref
@pytest.mark.parametrize("seed", range(10))
def test_gse_random_streams_match_reference_parser(seed):
    """random GSE traffic (more FragIDs in flight than slots, restarts, CRC failures), alone and mixed with TS
    frames and sync losses, cut into calls at random places"""
    rng = np.random.default_rng(2000 + seed)
    kbch = [7032, 14232, 32208, 58192][seed % 4]
    f = bbstream.random_gse_scenario(rng, kbch, nframes=30, ts_every=0 if seed % 2 else 4)
    cuts = sorted(set(int(x) for x in rng.integers(0, len(f) + 1, 5)) | {0, len(f)})
    _compare(kbch, [f[a:b] for a, b in zip(cuts[:-1], cuts[1:])], cap=65536 * 16)


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_ts_random_streams_match_reference_parser(seed):
    rng = np.random.default_rng(1000 + seed)
    kbch = [3072, 7032, 14232, 32208][seed % 4]
    f = bbstream.random_ts_scenario(rng, kbch)
    cuts = sorted(set(int(x) for x in rng.integers(0, len(f) + 1, 6)) | {0, len(f)})
    _compare(kbch, [f[a:b] for a, b in zip(cuts[:-1], cuts[1:])])
