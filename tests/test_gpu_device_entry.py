"""dvbs2fec_decode_batch_device -- the entry point whose time bench.py reports as `value` -- at the bench's own shape:
thousands of QPSK 1/2 normal frames at Es/N0 2.2 dB resident in device memory, one launch with every CTA of the
persistent grid handing itself several frame pairs.  Results (LDPC iteration counts, BCH correction counts, BBFRAME
bytes) are compared with the CPU oracle on the slowest frames and on random ones; the rest is covered by the
encode -> decode round trip."""
import ctypes as C

import numpy as np
import pytest

import orclib
from fec import pkg

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _pool(modcod, short, n, esn0, seed, ncodes=32):
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(seed)
    info = pkg.modcod_info(modcod, short)
    payloads = rng.integers(0, 256, (ncodes, info["kbch"] // 8), dtype=np.uint8)
    codes = np.stack([pkg.encode_fecframe(modcod, short, p) for p in payloads])
    a, sigma2 = 1 / np.sqrt(2.0), 1.0 / (2.0 * 10 ** (esn0 / 10.0))
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cw = torch.from_numpy(codes).to(dev)
    idx = torch.arange(n, device=dev) % ncodes
    pool = torch.empty((n, info["nldpc"]), dtype=torch.int8, device=dev)
    for f0 in range(0, n, 512):
        sl = idx[f0:f0 + 512]
        y = (1.0 - 2.0 * cw[sl].float()) * a
        y += torch.randn(y.shape, generator=g, device=dev) * float(np.sqrt(sigma2))
        pool[f0:f0 + 512] = torch.clamp(torch.round(4.0 * 2.0 * a * y / sigma2), -127, 127).to(torch.int8)
    return pool, payloads, idx.cpu().numpy(), info


def _oracle_check(short, rate, llr_rows, bb_rows, res_rows, kbch):
    o = orclib.oracle()
    want = np.zeros(kbch // 8, np.uint8)
    for k in range(llr_rows.shape[0]):
        it, co = C.c_int(), C.c_int()
        o.orc_decode_frame(short, rate, llr_rows[k].copy(), 25, want, C.byref(it), C.byref(co))
        assert (int(res_rows["ldpc_iters"][k]), int(res_rows["bch_corr"][k])) == (it.value, co.value), k
        assert np.array_equal(bb_rows[k], want), k


def test_decode_batch_device_at_the_bench_shape():
    n = 4096
    pool, payloads, idx, info = _pool(4, False, n, 2.2, 11)
    dev = pool.device
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=n, max_trials=25)
    dec.setDemodParams(4, False, False, 25)
    d_bb = torch.empty((n, info["kbch"] // 8), dtype=torch.uint8, device=dev)
    d_res = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    dec.decode_batch_device(pool.data_ptr(), n, d_bb.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
    bb = d_bb.cpu().numpy()
    it = res["ldpc_iters"].astype(np.int32)
    # every frame: round trip (a decoded frame is the payload that was encoded)
    ok = res["bch_corr"] >= 0
    assert ok.mean() > 0.995
    assert np.array_equal(bb[ok], payloads[idx][ok])
    assert 6.5 < np.where(it < 0, 25, it).mean() < 8.5          # the bench's operating point
    # oracle: the 24 slowest frames, the 8 fastest, 40 random ones
    order = np.argsort(np.where(it < 0, 26, it), kind="stable")
    pick = sorted(set(order[-24:].tolist()) | set(order[:8].tolist()) |
                  set(np.random.default_rng(1).choice(n, 40, replace=False).tolist()))
    rows = torch.tensor(pick, device=dev)
    _oracle_check(0, 3, pool[rows].cpu().numpy(), bb[pick], res[pick], info["kbch"])
    # the same frames again, odd count (last pair has one frame) and a different grid occupancy: identical results
    m = 1001
    dec.decode_batch_device(pool.data_ptr(), m, d_bb.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    res2 = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)[:m]
    assert np.array_equal(res2["ldpc_iters"], res["ldpc_iters"][:m])
    assert np.array_equal(res2["bch_corr"], res["bch_corr"][:m])
    assert np.array_equal(d_bb.cpu().numpy()[:m], bb[:m])
    dec.close()


def test_decode_batch_device_high_iteration_regime():
    """8PSK 3/5 normal (code n3/5) near threshold: most frames need 15-25 iterations, some fail (config 2's regime)."""
    n = 600
    pool, payloads, idx, info = _pool(5, False, n, 2.5, 12)    # QPSK-style LLRs on the n3/5 code, just above its threshold
    dev = pool.device
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=n, max_trials=25)
    dec.setDemodParams(5, False, False, 25)
    d_bb = torch.empty((n, info["kbch"] // 8), dtype=torch.uint8, device=dev)
    d_res = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    dec.decode_batch_device(pool.data_ptr(), n, d_bb.data_ptr(), d_res.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
    it = res["ldpc_iters"].astype(np.int32)
    assert np.where(it < 0, 25, it).mean() > 9
    pick = sorted(np.random.default_rng(2).choice(n, 24, replace=False).tolist())
    rows = torch.tensor(pick, device=dev)
    _oracle_check(0, 4, pool[rows].cpu().numpy(), d_bb.cpu().numpy()[pick], res[pick], info["kbch"])
    dec.close()


def test_device_entry_on_two_streams_and_between_sync_calls():
    """a device call on one stream, another on a second stream and a synchronous host call issued while
    they may still be running must not share scratch buffers: every result equals the serial one."""
    n = 512
    pool, payloads, idx, info = _pool(4, True, n, 2.6, 13)
    dev = pool.device
    dec = pkg.DVBS2Decoder(devices=[0], max_batch=n, max_trials=25)
    dec.setDemodParams(4, True, False, 25)
    kb = info["kbch"] // 8
    host_llr = pool.cpu().numpy()
    want_bb, want_res = dec.decode_batch(host_llr)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(3):
        a_bb = torch.zeros((n, kb), dtype=torch.uint8, device=dev)
        a_res = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
        b_bb = torch.zeros((n, kb), dtype=torch.uint8, device=dev)
        b_res = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        dec.decode_batch_device(pool.data_ptr(), n, a_bb.data_ptr(), a_res.data_ptr(), s1.cuda_stream)
        dec.decode_batch_device(pool.data_ptr(), n - 7, b_bb.data_ptr(), b_res.data_ptr(), s2.cuda_stream)
        c_bb, c_res = dec.decode_batch(host_llr[: 100 + rep])      # synchronous entry point in between
        torch.cuda.synchronize()
        outs.append((a_bb.cpu().numpy(), a_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1),
                     b_bb.cpu().numpy(), b_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1), c_bb, c_res))
    for a_bb, a_res, b_bb, b_res, c_bb, c_res in outs:
        assert np.array_equal(a_bb, want_bb) and np.array_equal(a_res["ldpc_iters"], want_res["ldpc_iters"])
        assert np.array_equal(b_bb[: n - 7], want_bb[: n - 7]) and np.array_equal(b_res["bch_corr"][: n - 7], want_res["bch_corr"][: n - 7])
        assert np.array_equal(c_bb, want_bb[: len(c_bb)]) and np.array_equal(c_res["ldpc_iters"], want_res["ldpc_iters"][: len(c_bb)])
    dec.close()


def test_symbols_stay_on_the_device_from_pl_sync_to_the_bbframe():
    """PL sync (K7) -> payload phase loop (K8) -> demapper, LDPC, BCH, descrambler (dvbs2fec_decode_plframes_device), every
    buffer a device buffer and one stream: the BBFRAMEs are the transmitted ones (after the loop has pulled in) and equal
    to what the host-buffer entry points give for the same symbols"""
    import plstream
    rng = np.random.default_rng(77)
    modcod, short, n, codenum = 4, True, 6, 3
    dec = pkg.DVBS2Decoder(max_batch=16)
    dec.setDemodParams(modcod, short, False, 25)
    pay = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
    pls = modcod << 2 | 2
    rn = plstream.pl_rn(codenum)
    frames = []
    for i in range(n):
        sym = pkg.modulate(modcod, short, False, pkg.encode_fecframe(modcod, short, pay[i])).view(np.complex64)
        frames.append(np.concatenate([plstream.plheader(pls), sym[90:] * np.array([1, 1j, -1, -1j])[rn[:len(sym) - 90]]]))
    x = np.concatenate([0.3 * (rng.normal(size=1500) + 1j * rng.normal(size=1500))] + frames + [frames[0][:200]]).astype(np.complex64)
    x = x * np.exp(1j * (0.4 + 2 * np.pi * 1.5e-5 * np.arange(len(x))))
    sigma = 0.667 * np.sqrt(0.5 / 10 ** 1.4)
    x = (x + sigma * (rng.normal(size=len(x)) + 1j * rng.normal(size=len(x)))).astype(np.complex64)
    # host-buffer entry points, stage by stage
    h = pkg.S2PLSyncBlock(90, False)
    h.pll_set_params(0.004, modcod, short, False, codenum)
    fr = h.process(x).reshape(-1, h.raw_frame_size)
    out, _ = h.pll_process(fr)
    bb_host, res_host = dec.decode_plframes(out.view(np.float32).reshape(len(fr), -1))
    h.close()
    # the same on device buffers
    g = pkg.S2PLSyncBlock(90, False)
    g.pll_set_params(0.004, modcod, short, False, codenum)
    rfs = g.raw_frame_size
    assert rfs == dec.plframe_symbols
    st = torch.cuda.Stream()
    d_x = torch.from_numpy(x.view(np.float32).copy()).cuda()
    d_fr = torch.zeros((n + 2) * rfs * 2, dtype=torch.float32, device="cuda")
    d_pl = torch.zeros_like(d_fr)
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    g.process_device(d_x.data_ptr(), len(x), d_fr.data_ptr(), n + 2, d_n.data_ptr(), st.cuda_stream)
    st.synchronize()
    nfr = int(d_n.item())
    assert nfr == len(fr) >= n - 1
    d_bb = torch.zeros((nfr, dec.kbch // 8), dtype=torch.uint8, device="cuda")
    d_res = torch.zeros((nfr, 16), dtype=torch.uint8, device="cuda")
    g.pll_process_device(d_fr.data_ptr(), nfr, rfs, d_pl.data_ptr(), 0, st.cuda_stream)
    dec.decode_plframes_device(d_pl.data_ptr(), nfr, d_bb.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    st.synchronize()
    bb = d_bb.cpu().numpy()
    res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
    assert np.array_equal(bb, bb_host) and np.array_equal(res["ldpc_iters"], res_host["ldpc_iters"])
    assert np.array_equal(res["bch_corr"], res_host["bch_corr"])
    good = [i for i in range(nfr) if any(np.array_equal(bb[i], p) for p in pay)]
    assert len(good) >= nfr - 1 and (res["bch_corr"][1:] >= 0).all()
    g.close()
    dec.close()


@pytest.mark.parametrize("case", [(4, True, False, 3, 0), (4, True, True, 0, 1), (12, True, False, 7, 2)])
def test_s2_demod_stage_in_one_call(case):
    """dvbs2fec_s2_demod_process (symbols -> BBFRAMEs, the module's process() behind its sample-domain front end) over two
    calls cut inside a frame: the same frames, BBFRAME bytes, results, coarse frequency errors and PLHEADER fields as the
    stages called one after the other through their host-buffer entry points; and the transmitted payloads come back"""
    import plstream
    modcod, short, pilots, codenum, seed = case
    rng = np.random.default_rng(500 + seed)
    n = 6
    dec = pkg.DVBS2Decoder(max_batch=16)
    dec.setDemodParams(modcod, short, pilots, 25)
    info = pkg.modcod_info(modcod, short, pilots)
    pay = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
    pls = modcod << 2 | int(short) << 1 | int(pilots)
    rn = plstream.pl_rn(codenum)
    frames = []
    for i in range(n):
        sym = pkg.modulate(modcod, short, pilots, pkg.encode_fecframe(modcod, short, pay[i])).view(np.complex64).copy()
        body = sym[90:]
        if pilots:      # the in-tree modulator leaves the pilot blocks empty: unmodulated pilots (1 + j) / sqrt 2, scaled as the payload
            k = np.arange(len(body)) % (1440 + 36)
            body[k >= 1440] = (1 + 1j) / np.sqrt(2) * np.abs(body[0])
        frames.append(np.concatenate([plstream.plheader(pls) * np.abs(body[0]), body * np.array([1, 1j, -1, -1j])[rn[:len(body)]]]))
    x = np.concatenate([0.3 * (rng.normal(size=1500) + 1j * rng.normal(size=1500))] + frames + [frames[0][:300]]).astype(np.complex64)
    x = x * np.exp(1j * (0.4 + 2 * np.pi * 1.5e-5 * np.arange(len(x))))
    sigma = np.abs(frames[0][100]) * np.sqrt(0.5 / 10 ** (1.4 if modcod == 4 else 1.8))
    x = (x + sigma * (rng.normal(size=len(x)) + 1j * rng.normal(size=len(x)))).astype(np.complex64)
    cut = len(x) // 2 + 1234
    # stage by stage, host buffers
    slots = info["nldpc"] // info["bits"] // 90
    h = pkg.S2PLSyncBlock(slots, pilots)
    h.plhdr_set_params(0.004)
    h.pll_set_params(0.004, modcod, short, pilots, codenum)
    want = []
    for seg in (x[:cut], x[cut:]):
        fr = h.process(seg).reshape(-1, h.raw_frame_size)
        if not len(fr):
            want.append(None)
            continue
        fed = h.coarse_fed(fr, pilots, pls, codenum)
        out, _ = h.pll_process(fr)
        _, hres, _ = h.plhdr_process(fr)
        bb, res = dec.decode_plframes(out.view(np.float32).reshape(len(fr), -1))
        want.append((bb, res, np.asarray(fed, np.float32), hres))
    h.close()
    g = pkg.DVBS2DemodStage(max_batch=16)
    g.setDemodParams(modcod, short, pilots, 25, 0.004, 0.004, codenum)
    got_bb = []
    for seg, w in zip((x[:cut], x[cut:]), want):
        bb, res, fed, hres = g.process(seg)
        if w is None:
            assert len(bb) == 0
            continue
        assert np.array_equal(bb, w[0]) and np.array_equal(res["ldpc_iters"], w[1]["ldpc_iters"]) and np.array_equal(res["bch_corr"], w[1]["bch_corr"])
        assert np.array_equal(fed.view(np.uint32), w[2].view(np.uint32)) and np.array_equal(hres, w[3])
        got_bb.append(bb)
    got_bb = np.concatenate(got_bb)
    assert len(got_bb) >= n - 1
    good = sum(any(np.array_equal(b, p) for p in pay) for b in got_bb)
    assert good >= len(got_bb) - 1      # (the first frame: the loop is still pulling in)
    assert len(g.process(np.zeros(0, np.complex64))[0]) == 0
    g.close()
    dec.close()


def test_s2_demod_stage_with_ts_output():
    """symbols -> TS packets in one call (the BBFRAME parser behind the stage, its state carried over calls): the same bytes
    as BBFrameTSParser.work on the BBFRAMEs the plain stage call delivers, and the transmitted packets come back"""
    import bbstream
    import plstream
    modcod, short, codenum = 4, True, 2
    rng = np.random.default_rng(606)
    info = pkg.modcod_info(modcod, short, False)
    packets = bbstream.ts_packets(200, rng)
    frames_bb, _ = bbstream.ts_bbframes(info["kbch"], packets)
    frames_bb = frames_bb[:7]
    pls = modcod << 2 | 2
    rn = plstream.pl_rn(codenum)
    frames = []
    for bbf in frames_bb:
        sym = pkg.modulate(modcod, short, False, pkg.encode_fecframe(modcod, short, bbf)).view(np.complex64).copy()
        body = sym[90:]
        frames.append(np.concatenate([plstream.plheader(pls) * np.abs(body[0]), body * np.array([1, 1j, -1, -1j])[rn[:len(body)]]]))
    x = np.concatenate([0.3 * (rng.normal(size=900) + 1j * rng.normal(size=900))] + frames + [frames[0][:300]]).astype(np.complex64)
    x = x * np.exp(1j * (0.1 + 2 * np.pi * 1e-5 * np.arange(len(x))))
    sigma = np.abs(frames[0][100]) * np.sqrt(0.5 / 10 ** 1.4)
    x = (x + sigma * (rng.normal(size=len(x)) + 1j * rng.normal(size=len(x)))).astype(np.complex64)
    cut = len(x) // 2 + 77
    a = pkg.DVBS2DemodStage(max_batch=16)
    a.setDemodParams(modcod, short, False, 25, 0.004, 0.004, codenum)
    b = pkg.DVBS2DemodStage(max_batch=16)
    b.setDemodParams(modcod, short, False, 25, 0.004, 0.004, codenum)
    parser = pkg.BBFrameTSParser()
    parser.setFrameSize(info["kbch"])
    got_all = []
    for seg in (x[:cut], x[cut:]):
        bb, _, _, _ = a.process(seg)
        want = parser.work(bb, len(bb)) if len(bb) else np.zeros(0, np.uint8)
        got, nfr = b.process_ts(seg)
        assert nfr == len(bb) and np.array_equal(got, want)
        got_all.append(got)
    ts = np.concatenate(got_all).reshape(-1, 188)
    sent = {p.tobytes() for p in packets}
    assert len(ts) >= 5 * (info["kbch"] // 8 - 10) // 188 and sum(p.tobytes() in sent for p in ts) >= len(ts) - 20
    for h in (a, b, parser):
        h.close()
