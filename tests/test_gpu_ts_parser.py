"""Row 8(f)-1 on the GPU: dvbs2fec_ts_work (BBFRAME -> TS packets) against the CPU oracle of
BBFrameTSParser::work, against golden outputs of the reference parser, and through properties at full size."""
import ctypes as C
import os

import numpy as np
import pytest

import bbstream
import orclib
from fec import pkg
from test_ts_parser_oracle import OrcParser, header_fields, KBCH

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu_header_fields(h):
    return [h.ts_gs, h.sis_mis, h.ccm_acm, h.issyi, h.npd, h.ro, h.isi, h.upl, h.dfl, h.sync, h.syncd]


def compare(kbch, batches, cap=65536 * 10):
    o = OrcParser(kbch)
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    for k, frames in enumerate(batches):
        c = cap[k] if isinstance(cap, (list, tuple)) else cap
        want = o.work(frames, c)
        got = g.work(frames, len(frames), c)
        assert np.array_equal(got, want)
        st = o.stats()
        assert (g.last_bb_cnt, g.last_bb_proc) == (st["cnt"], st["proc"])
        assert g.last_gse_crc_err == st["gse_err"]
        assert g.have_header == bool(st["have"])
        if st["have"]:
            assert gpu_header_fields(g.last_header) == header_fields(st["hdr"])
    o.close()
    g.close()


@pytest.mark.parametrize("name", list(KBCH))
def test_ts_matches_oracle(name):
    kbch = KBCH[name]
    rng = np.random.default_rng(len(name) + kbch)
    pk = bbstream.ts_packets(max(300, 40 * kbch // 1504), rng)
    frames, _ = bbstream.ts_bbframes(kbch, pk, first_byte=int(rng.integers(0, 188)))
    cuts = sorted(set(int(x) for x in rng.integers(1, len(frames), 5)))
    compare(kbch, [frames[a:b] for a, b in zip([0] + cuts, cuts + [len(frames)])])


def test_ts_odd_streams_match_oracle():
    kbch = KBCH["s1/4"]
    f = bbstream.odd_ts_scenario(np.random.default_rng(5), kbch)
    compare(kbch, [f[:60], f[60:61], f[61:]])
    compare(kbch, [f[i:i + 1] for i in range(len(f))])          # one frame per call: all state crosses calls
    compare(kbch, [f[:0], f])                                   # empty call first


def test_ts_output_room_rule_matches_oracle():
    rng = np.random.default_rng(6)
    kbch = KBCH["n1/2"]
    frames, _ = bbstream.ts_bbframes(kbch, bbstream.ts_packets(200, rng))
    for cap in (0, 188, 189, 190, 188 * 7 + 5, 188 * 21, 188 * 21 + 1, 188 * 22, 188 * 43):
        compare(kbch, [frames[:3]], cap=cap)


def test_ts_parser_recovers_after_running_out_of_room():
    """a call that stops on the room rule with >= 188 bytes of its frame unconsumed (the reference overruns its
    reassembly buffer there): the parser drops sync instead of carrying an impossible byte count, and the
    following calls -- ample room, and out of room again -- stay exact and in bounds"""
    rng = np.random.default_rng(16)
    for kbch in (KBCH["n1/2"], KBCH["s1/4"]):
        frames, _ = bbstream.ts_bbframes(kbch, bbstream.ts_packets(400, rng), first_byte=7)
        for cap0 in (0, 188, 189, 188 * 3 + 1, 188 * 9):
            compare(kbch, [frames[:4], frames[4:9], frames[9:12], frames[12:]], cap=[cap0, 65536 * 10, 188 * 2 + 5, 65536 * 10])


def test_gse_frames_can_be_counted_without_unpacking():
    rng = np.random.default_rng(8)
    kbch = KBCH["n1/2"]
    frames = bbstream.gse_bbframes(kbch, bbstream.gse_scenario(rng))
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    g.set_gse(-1)
    out = g.work(frames)
    assert len(out) == 0 and g.last_bb_proc == len(frames) and g.gse_frames == len(frames)
    assert g.last_header.ts_gs == 1
    g.close()


def test_gse_pdus_match_oracle():
    """the scenario of the oracle's own test (complete PDUs, three interleaved reassemblies, a CRC-32 failure), as one
    call, frame by frame (every reassembly crosses calls: heads carried in the device-side buffers), and in between"""
    rng = np.random.default_rng(8)
    for name in ("n1/2", "n9/10"):
        kbch = KBCH[name]
        frames = bbstream.gse_bbframes(kbch, bbstream.gse_scenario(rng))
        compare(kbch, [frames])
        compare(kbch, [frames[i:i + 1] for i in range(len(frames))])
        compare(kbch, [frames[:2], frames[2:3], frames[3:]])


@pytest.mark.parametrize("seed", range(10))
def test_gse_random_streams_match_oracle(seed):
    """random GSE traffic -- more FragIDs in flight than reassembly slots, FragIDs restarted, CRC failures, PDUs that
    never end -- alone and mixed with TS frames and sync losses (a frame entered out of sync is walked one byte
    late, :157-168), cut into calls at random places"""
    rng = np.random.default_rng(2000 + seed)
    kbch = [7032, 14232, 32208, 58192][seed % 4]
    f = bbstream.random_gse_scenario(rng, kbch, nframes=30 if seed < 8 else 300, ts_every=0 if seed % 2 else 4)
    cuts = sorted(set(int(x) for x in rng.integers(0, len(f) + 1, 5)) | {0, len(f)})
    compare(kbch, [f[a:b] for a, b in zip(cuts[:-1], cuts[1:])], cap=65536 * 64)
    compare(kbch, [f], cap=65536 * 64)


def test_gse_counters_and_refused_calls():
    rng = np.random.default_rng(77)
    kbch = 7032   # s1/2
    f = bbstream.random_gse_scenario(rng, kbch, nframes=40)
    want = OrcParser(kbch).work(f, 65536 * 16)
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    # too little room for the GSE output: refused as a whole (the reference would overrun the buffer)
    with pytest.raises(pkg.DVBS2FecError) as e:
        g.work(f, len(f), len(want) // 2)
    assert e.value.code == pkg.ENOSPC
    g.close()
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    got = g.work(f, len(f), len(want) + 189)
    assert np.array_equal(got, want)
    c = g.gse_counters
    assert c["pdus"] > 20 and c["crc_errors"] >= 1 and c["malformed"] == 0
    # a packet whose length leads outside the input: reported, in bounds, nothing else disturbed
    bad = f[:3].copy()
    bad[2, 10] = 0xC0 | 0x30 | 0x0F   # complete PDU, no label, 12-bit length 0xFFF.. inside the LAST frame of the call
    bad[2, 11] = 0xFF
    g.work(bad)
    assert g.gse_counters["malformed"] == 1
    g.close()


def test_gse_device_buffers_async_pool():
    """dvbs2fec_ts_work_device: the GSE pass is enqueued blindly behind the TS pass with a descriptor pool sized in
    advance; results equal the host-buffer call, a pool that is too small refuses the call"""
    import torch
    rng = np.random.default_rng(5150)
    kbch = KBCH["n1/2"]
    f = bbstream.random_gse_scenario(rng, kbch, nframes=60, ts_every=5)
    want = OrcParser(kbch).work(f, 65536 * 64)
    dev = torch.device("cuda", 0)
    d_bb = torch.from_numpy(f).to(dev)
    d_out = torch.zeros(f.size + 3 * 70000, dtype=torch.uint8, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    got, at = [], 0
    for n in (1, 7, len(f) - 8):
        g.work_device(d_bb[at:at + n].data_ptr(), n, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        got.append(d_out[:int(d_n.item())].cpu().numpy().copy())
        at += n
    assert np.array_equal(np.concatenate(got), want)
    g.close()
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    g.set_gse(3)
    g.work_device(d_bb.data_ptr(), len(f), d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    assert int(d_n.item()) == pkg.ENOSPC
    g.close()


@pytest.mark.parametrize("name", ["tsparse_ts_n12", "tsparse_odd_s14", "tsparse_gse_n12", "tsparse_gsemix_s12"])
def test_cuda_reproduces_reference_ts_parser(name):
    gold = dict(np.load(os.path.join(GOLD, name + ".npz")))
    g = pkg.BBFrameTSParser()
    g.setFrameSize(int(gold["kbch"]))
    at = 0
    for k, (a, b) in enumerate(zip(gold["cuts"][:-1], gold["cuts"][1:])):
        out = g.work(gold["frames"][a:b], int(b - a))
        n = int(gold["out_len"][k])
        assert len(out) == n and np.array_equal(out, gold["out"][at:at + n])
        at += n
        assert [g.last_bb_cnt, g.last_bb_proc, g.last_gse_crc_err] == [int(x) for x in gold["stats"][k][11:14]]
        if g.have_header:
            assert gpu_header_fields(g.last_header) == [int(x) for x in gold["stats"][k][:11]]
    g.close()


def test_full_size_stream_round_trip_on_device_buffers():
    """4096 normal-frame BBFRAMEs (16 MB) through the device-pointer entry point, fed in uneven calls: every
    TS packet after the first sync point comes back exactly once, in order, with 0x47 restored"""
    import torch
    rng = np.random.default_rng(9)
    kbch = KBCH["n1/2"]
    kb = kbch // 8
    nframes = 4096
    npk = nframes * (kb - 10) // 188 + 2
    pk = bbstream.ts_packets(npk, rng)
    stream = pk.copy()
    stream[:, 0] = 0xEE                                       # stands in for the CRC-8 slot (never checked)
    stream = stream.reshape(-1)
    df = kb - 10
    frames = np.zeros((nframes, kb), np.uint8)
    hdrs = {}
    for f in range(nframes):
        pos = f * df
        syncd = ((-pos) % 188) * 8
        if syncd not in hdrs:
            hdrs[syncd] = bbstream.bbheader(0xF0, 1504, df * 8, 0x47, syncd)
        frames[f, :10] = hdrs[syncd]
        frames[f, 10:] = stream[pos:pos + df]
    dev = torch.device("cuda", 0)
    d_bb = torch.from_numpy(frames).to(dev)
    d_out = torch.zeros(nframes * kb + 188, dtype=torch.uint8, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    st = torch.cuda.current_stream()
    got, at = [], 0
    for n in (1, 2, 1000, 3, 3090):
        g.work_device(d_bb[at:at + n].data_ptr(), n, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        got.append(d_out[:int(d_n.item())].cpu().numpy().copy())
        at += n
    assert at == nframes
    out = np.concatenate(got).reshape(-1, 188)
    assert len(out) == (nframes * df - 1) // 188
    assert np.array_equal(out, pk[:len(out)])
    g.close()


@pytest.mark.parametrize("seed", range(12))
def test_ts_random_streams_match_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    kbch = [3072, 7032, 14232, 32208][seed % 4]
    f = bbstream.random_ts_scenario(rng, kbch, nframes=80 if seed % 3 else 700)   # 700: several frames per plan thread
    cuts = sorted(set(int(x) for x in rng.integers(0, len(f) + 1, 6)) | {0, len(f)})
    compare(kbch, [f[a:b] for a, b in zip(cuts[:-1], cuts[1:])])
    if seed % 4 == 0:   # output room that runs out somewhere in the middle: the sequential fallback
        total = len(OrcParser(kbch).work(f))
        compare(kbch, [f], cap=max(189, total // 2))


def _llr_of_bbframes(modcod, short, frames):
    """BBFRAME bytes (already carrying a BBHEADER) -> clean LLRs of their FECFRAMEs"""
    return np.stack([np.where(pkg.encode_fecframe(modcod, short, f) > 0, -40, 40).astype(np.int8) for f in frames])


@pytest.mark.parametrize("modcod,scenario", [(1, "odd"), (4, "plain")])
def test_queue_delivers_ts_packets_fused_behind_the_decoder(modcod, scenario):
    """dvbs2fec_set_ts_output: BBFRAMEs never leave the device; what collect_ts returns equals BBFrameTSParser::work
    (oracle) over the same BBFRAMEs, across many small batches (parser state carried on the device)"""
    short = True
    info = pkg.modcod_info(modcod, short)
    kbch = info["kbch"]
    rng = np.random.default_rng(31 + modcod)
    if scenario == "odd":
        frames = bbstream.odd_ts_scenario(rng, kbch)
    else:
        frames, _ = bbstream.ts_bbframes(kbch, bbstream.ts_packets(300, rng), first_byte=33)
    want = OrcParser(kbch).work(frames)
    llr = _llr_of_bbframes(modcod, short, frames)
    d = pkg.DVBS2Decoder(max_batch=8, max_latency_us=300, max_trials=10)
    try:
        d.set_ts_output(True)
        d.setDemodParams(modcod, short, False)
        got, tags = [], []
        for i in range(len(llr)):
            while True:
                rc = pkg.lib().dvbs2fec_submit_llr(d._h, llr[i].ctypes.data_as(C.c_void_p), i)
                if rc != pkg.EAGAIN:
                    break
                ts, res = d.collect_ts(cap=188 * 7, timeout_us=100_000)       # small bites: batches split over calls
                got.append(ts.copy()); tags += list(res["tag"])
            assert rc == 0
        d.flush()
        while len(tags) < len(llr):
            ts, res = d.collect_ts(cap=188 * 11, timeout_us=2_000_000)
            got.append(ts.copy()); tags += list(res["tag"])
        out = np.concatenate(got)
        assert tags == list(range(len(llr)))
        assert len(out) == len(want) and np.array_equal(out, want)
        # switching the MODCOD starts the parser over (setFrameSize): a fresh oracle parser gives the same
        d.setDemodParams(modcod, short, False)
        for i in range(6):
            d.submit_llr(llr[i], 100 + i)
        d.flush()
        ts, res = d.collect_ts(timeout_us=2_000_000)
        assert np.array_equal(ts, OrcParser(kbch).work(frames[:6])) and len(res) == 6
        # fewer result records per call than a batch holds: records trickle out, nothing is stuck
        d.setDemodParams(modcod, short, False)
        for i in range(8):
            d.submit_llr(llr[i], 200 + i)
        d.flush()
        got, tags = [], []
        for _ in range(20):
            ts, res = d.collect_ts(max_results=3, timeout_us=500_000)
            got.append(ts.copy()); tags += list(res["tag"])
            if len(tags) == 8:
                break
        assert tags == list(range(200, 208))
        assert np.array_equal(np.concatenate(got), OrcParser(kbch).work(frames[:8]))
    finally:
        d.close()
