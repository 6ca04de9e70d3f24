#!/usr/bin/env python3
"""Small whole-stage run for compute-sanitizer (memcheck / racecheck): a few short and normal frames."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
rng = np.random.default_rng(0)
dec = pkg.DVBS2Decoder(max_batch=8, max_trials=6)
for modcod, short, n in ((4, True, 5), (13, True, 3), (4, False, 3)):
    dec.setDemodParams(modcod, short, False, 6)
    pay = rng.integers(0, 256, (n, dec.kbch // 8), dtype=np.uint8)
    pl = np.stack([pkg.modulate(modcod, short, False, pkg.encode_fecframe(modcod, short, pay[i])) for i in range(n)])
    noisy = pl.view(np.float32) + rng.normal(0, 0.08, (n, dec.plframe_symbols * 2)).astype(np.float32)
    bb, res = dec.decode_plframes(noisy)
    print(modcod, short, (bb == pay).all(), res["ldpc_iters"].tolist(), res["bch_corr"].tolist())
# LLR path near threshold: several LDPC iterations, level-scheduled layers, frozen-lane path, BCH corrections
for modcod, short, esn0 in ((4, True, 1.6), (5, True, 3.2), (4, False, 1.6)):
    dec.setDemodParams(modcod, short, False, 6)
    n = 5
    a = 1 / np.sqrt(2.0)
    sigma2 = 1.0 / (2.0 * 10 ** (esn0 / 10.0))
    codes = np.stack([pkg.encode_fecframe(modcod, short, rng.integers(0, 256, dec.kbch // 8, dtype=np.uint8)) for _ in range(n)])
    y = (1.0 - 2.0 * codes.astype(np.float32)) * a + rng.normal(0, np.sqrt(sigma2), codes.shape).astype(np.float32)
    llr = np.clip(np.rint(4.0 * 2.0 * a * y / sigma2), -127, 127).astype(np.int8)
    llr[0, :40] = -llr[0, :40]
    bb, res = dec.decode_batch(llr)
    print(modcod, short, res["ldpc_iters"].tolist(), res["bch_corr"].tolist())
dec.close()
# frame queue: small batches through the staging ring, partial collects
dec = pkg.DVBS2Decoder(max_batch=2, max_latency_us=300, max_trials=6)
dec.setDemodParams(4, True, False, 6)
codes = [pkg.encode_fecframe(4, True, rng.integers(0, 256, dec.kbch // 8, dtype=np.uint8)) for _ in range(7)]
got = 0
for i, c in enumerate(codes):
    dec.submit_llr(np.where(c > 0, -20, 20).astype(np.int8), i)
    got += len(dec.collect(1)[1])
dec.flush()
while got < len(codes):
    got += len(dec.collect(3, timeout_us=2_000_000)[1])
print("queue", got)
dec.close()
# BBFRAME -> TS: header faults, resyncs, carried units, tiny output room, calls of every size
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bbstream  # noqa: E402
ts = pkg.BBFrameTSParser()
ts.setFrameSize(3072)
f = bbstream.odd_ts_scenario(np.random.default_rng(5), 3072)
tot = 0
for a, b, cap in ((0, 60, 655360), (60, 61, 655360), (61, 90, 189), (90, 120, 655360)):
    tot += len(ts.work(f[a:b], b - a, cap))
ts.setFrameSize(32208)
pk = bbstream.ts_packets(300, np.random.default_rng(6))
fr, _ = bbstream.ts_bbframes(32208, pk, first_byte=50)
tot += len(ts.work(fr))
print("ts bytes", tot)
ts.close()
# fused TS output of the queue
dec = pkg.DVBS2Decoder(max_batch=4, max_latency_us=300, max_trials=6)
dec.set_ts_output(True)
dec.setDemodParams(1, True, False, 6)
frames = bbstream.odd_ts_scenario(np.random.default_rng(7), 3072)[:24]
n_ts, n_res = 0, 0
for i, f in enumerate(frames):
    llr = np.where(pkg.encode_fecframe(1, True, f) > 0, -30, 30).astype(np.int8)
    while True:
        try:
            dec.submit_llr(llr, i)
            break
        except pkg.DVBS2FecError as e:
            if e.code != pkg.EAGAIN:
                raise
            ts, res = dec.collect_ts(cap=188 * 5, timeout_us=200_000)
            n_ts += len(ts); n_res += len(res)
dec.flush()
while n_res < len(frames):
    ts, res = dec.collect_ts(timeout_us=2_000_000)
    n_ts += len(ts); n_res += len(res)
print("fused ts bytes", n_ts, "frames", n_res)
dec.close()

# GSE extraction: interleaved reassemblies, CRC failures, more FragIDs than slots, mixed with TS frames and sync losses;
# host-buffer calls of several sizes, a refused call (too little room), a device-buffer call with a small pool
ts = pkg.BBFrameTSParser()
ts.setFrameSize(7032)
g = bbstream.random_gse_scenario(np.random.default_rng(8), 7032, nframes=40, ts_every=4)
tot = 0
for a, b in ((0, 1), (1, 9), (9, 10), (10, 33), (33, len(g))):
    tot += len(ts.work(g[a:b], b - a, 65536 * 16))
try:
    ts.work(g, len(g), 1000)
except pkg.DVBS2FecError as e:
    print("refused:", e.code)
print("gse bytes", tot, ts.gse_counters)
ts.close()
# PL front end: sync on a noisy stream in uneven calls, PLHEADER demodulation, coarse FED with pilots
import plstream  # noqa: E402
x = plstream.stream((4 << 2) | 1, 36, True, 6, np.random.default_rng(9), esn0_db=5.0, lead=500, cfo=1e-4)
ps = pkg.S2PLSyncBlock(36, True)
frames = []
for a, b in ((0, 1), (1, 100), (100, 3400), (3400, 3401), (3401, 12000), (12000, len(x))):
    frames.append(ps.process(x[a:b]))
fr = np.concatenate(frames).reshape(-1, ps.raw_frame_size)
hdr, res, loop = ps.plhdr_process(fr)
err = ps.coarse_fed(fr, True, (4 << 2) | 1, 3)
print("plsync frames", len(fr), res[:, 0].tolist(), [round(float(e), 4) for e in err])
# payload phase loop on the delivered frames (pilots on: table cells and the pilot stretches), and a 32APSK stream
ps.close()
x = plstream.stream((4 << 2) | 3, 90, True, 2, np.random.default_rng(11), esn0_db=6.0, lead=0, cfo=2e-5, codenum=3)
ps = pkg.S2PLSyncBlock(90, True)
ps.pll_set_params(0.004, 4, True, True, 3)
po, pst = ps.pll_process(x.reshape(2, -1))
print("pll", po.shape, [round(float(v), 4) for v in pst[-1]], ps.pll_rounds())
ps.close()
x = plstream.stream(28 << 2, 36, False, 2, np.random.default_rng(10), esn0_db=18.0, lead=0, cfo=1e-5, bits=3)
ps = pkg.S2PLSyncBlock(36, False)
ps.pll_set_params(0.004, 28, True, False, 0)
po, pst = ps.pll_process(x.reshape(2, -1))
print("pll 32apsk", po.shape, ps.pll_rounds())
ps.close()
# quantised symbols path and zero-copy submit
dec = pkg.DVBS2Decoder(max_batch=4, max_latency_us=300, max_trials=6)
dec.setDemodParams(13, True, False, 6)
pay = rng.integers(0, 256, (3, dec.kbch // 8), dtype=np.uint8)
pl = np.stack([pkg.modulate(13, True, False, pkg.encode_fecframe(13, True, pay[i])) for i in range(3)])
noisy = pl.view(np.float32) + rng.normal(0, 0.05, (3, dec.plframe_symbols * 2)).astype(np.float32)
idx = dec.quantize_plframes(noisy)
bb, res = dec.decode_plframes_idx(idx)
print("idx path", (bb == pay).all())
dec.setDemodParams(4, True, False, 6)
for i in range(5):
    slot = dec.acquire_llr()
    slot[:] = np.where(pkg.encode_fecframe(4, True, rng.integers(0, 256, dec.kbch // 8, dtype=np.uint8)) > 0, -20, 20)
    dec.commit(i)
dec.flush()
print("zero-copy", len(dec.collect(8, timeout_us=2_000_000)[1]))
dec.close()
# DVB-S outer decoder: clean, correctable and hopeless packets, state over calls, the module's 204-byte frame stride
import dvbs_stream  # noqa: E402
ts_, ch = dvbs_stream.outer_stream(6, np.random.default_rng(12))
bad = dvbs_stream.add_errors(ch, np.random.default_rng(13), per_packet=(0, 14))
od = pkg.DVBSOuterDecoder()
o1, e1 = od.process(bad, 2)
o2, e2 = od.process(bad[2 * 1632:], 4)
o3, e3 = od.process(bad, 9, 204)
print("dvbs outer", o1.shape, o2.shape, o3.shape, int(e1.sum() + e2.sum()), int(e3.sum()))
od.close()
# DVB-S inner half: TS deframer (window over two calls), Viterbi search + lock + decode at a continuous-depuncturer rate,
# search on noise, the whole stage from symbols
frames_, bits_ = None, np.unpackbits(ch)
df = pkg.DVBSTSDeframer()
f1 = df.work(bits_[:20001]); f2 = df.work(bits_[20001:])
print("dvbs deframer", f1.shape, f2.shape, df.stats())
df.close()
vrng = np.random.default_rng(14)
vit = pkg.DVBSViterbi(0.15, 2)
sv = dvbs_stream.inner_softs(vrng.integers(0, 2, 5462 * 4, dtype=np.uint8), 1, vrng, sigma=8.0)[:3 * 8192]
b1 = vit.process(sv)
b2 = vit.process(np.clip(np.rint(vrng.normal(0, 40, 4 * 8192)), -127, 127).astype(np.int8))
print("dvbs viterbi", len(b1), len(b2), vit.stats(), vit.counters())
vit.close()
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_vit import dvbs_symbols  # noqa: E402
_, csyms = dvbs_symbols(30, 0, np.random.default_rng(400), lead=14)      # (a short stream would be spent on the reference's first, wrong lock)
dm = pkg.DVBSDemod(frame_stride=1632)
t1 = dm.process(csyms)
print("dvbs chain", t1.shape, dm.stats())
dm.close()
# the DVB-S2 stage in one call (K7 -> K8 -> K1..K3 -> K6 on one stream), two calls cut inside a frame
import bbstream  # noqa: E402
import plstream  # noqa: E402
srng = np.random.default_rng(15)
sinfo = pkg.modcod_info(4, True, False)
sbb, _ = bbstream.ts_bbframes(sinfo["kbch"], bbstream.ts_packets(60, srng))
sfr = []
for bbf in sbb[:3]:
    sym = pkg.modulate(4, True, False, pkg.encode_fecframe(4, True, bbf)).view(np.complex64).copy()
    sfr.append(np.concatenate([plstream.plheader(4 << 2 | 2) * np.abs(sym[90]), sym[90:] * np.array([1, 1j, -1, -1j])[plstream.pl_rn(1)[:len(sym) - 90]]]))
sx = np.concatenate([0.3 * (srng.normal(size=500) + 1j * srng.normal(size=500))] + sfr + [sfr[0][:200]]).astype(np.complex64)
sx = (sx + 0.03 * (srng.normal(size=len(sx)) + 1j * srng.normal(size=len(sx)))).astype(np.complex64)
stg = pkg.DVBS2DemodStage(max_batch=8)
stg.setDemodParams(4, True, False, 6, 0.004, 0.004, 1)
t_a, n_a = stg.process_ts(sx[:9000])
t_b, n_b = stg.process_ts(sx[9000:])
print("s2 stage", len(t_a), n_a, len(t_b), n_b)
stg.close()
