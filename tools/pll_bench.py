#!/usr/bin/env python3
"""K8 (payload phase loop) on device buffers against the oracle on one host core: frames per second of ONE stream
(the loop is a recurrence: frames of a stream cannot overlap), rounds per block of 32 symbols."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import torch
from test_pll_oracle import OrcPll, frames_for

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); a = ap.parse_args()
res = []
for name, modcod, slots, pilots, esn0 in (("qpsk", 4, 360, False, 8.0), ("qpsk", 4, 360, True, 8.0), ("8psk", 12, 240, False, 12.0),
                                          ("16apsk", 18, 180, False, 16.0), ("32apsk", 28, 144, False, 20.0)):
    rng = np.random.default_rng(1)
    nfr = 24
    pls, fr = frames_for(name, slots, pilots, nfr, rng, esn0, 2e-5, 0, modcod=modcod)
    g = pkg.S2PLSyncBlock(slots, pilots)
    g.pll_set_params(0.004, modcod, False, pilots, 0)
    rfs = fr.shape[1]
    d_in = torch.from_numpy(fr.view(np.float32).copy()).cuda()
    d_out = torch.zeros_like(d_in)
    st = torch.cuda.current_stream().cuda_stream
    g.pll_process_device(d_in.data_ptr(), nfr, rfs, d_out.data_ptr(), 0, st); torch.cuda.synchronize()
    g.pll_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.pll_process_device(d_in.data_ptr(), nfr, rfs, d_out.data_ptr(), 0, st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nfr
    rounds = g.pll_rounds() / (nfr * ((g.pll_frame_symbols + 31) // 32))
    o = OrcPll(0.004, name, slots, pilots, pls, 0)
    t0 = time.perf_counter()
    for f in fr[:8]: o.process(f)
    cpu_ms = (time.perf_counter() - t0) / 8 * 1e3
    got = d_out.cpu().numpy().view(np.complex64).reshape(nfr, rfs)
    o2 = OrcPll(0.004, name, slots, pilots, pls, 0)
    dev = max(np.abs(got[f, :o2.total] - o2.process(fr[f])[0]).max() for f in range(nfr))
    res.append(dict(constellation=name, pilots=pilots, symbols_per_frame=g.pll_frame_symbols, gpu_ms_per_frame=round(ms, 4),
                    gpu_msym_s=round(g.pll_frame_symbols / ms / 1e3, 2), rounds_per_block=round(rounds, 3),
                    oracle_ms_per_frame_1core=round(cpu_ms, 3), speedup=round(cpu_ms / ms, 2), max_abs_dev_vs_oracle=float(dev)))
    print(res[-1], flush=True)
    g.close()
# eight transponders at once (QPSK 1/2 normal): a warp each, one launch
blocks, ins, outs = [], [], []
rng = np.random.default_rng(2)
for k in range(8):
    pls, fr = frames_for("qpsk", 360, False, 8, rng, 8.0, 1e-5 * (k + 1), 0, modcod=4)
    g = pkg.S2PLSyncBlock(360, False); g.pll_set_params(0.004, 4, False, False, 0)
    blocks.append(g); ins.append(torch.from_numpy(fr.view(np.float32).copy()).cuda()); outs.append(torch.zeros_like(ins[-1]))
st = torch.cuda.current_stream().cuda_stream
args = (blocks, [t.data_ptr() for t in ins], 8, fr.shape[1], [t.data_ptr() for t in outs], st)
pkg.pll_process_multi_device(*args); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pkg.pll_process_multi_device(*args); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
res.append(dict(streams=8, frames_per_stream=8, ms=round(ms, 3), frames_per_s=round(64 / ms * 1e3, 1), msym_s=round(64 * 32490 / ms / 1e3, 1)))
print(res[-1], flush=True)
# the sequential walk (one device thread) for comparison
pls, fr = frames_for("qpsk", 360, False, 2, np.random.default_rng(3), 8.0, 2e-5, 0, modcod=4)
g = pkg.S2PLSyncBlock(360, False); g.pll_set_params(0.004, 4, False, False, 0); g.pll_set_sequential(1)
d_in = torch.from_numpy(fr.view(np.float32).copy()).cuda(); d_out = torch.zeros_like(d_in)
g.pll_process_device(d_in.data_ptr(), 1, fr.shape[1], d_out.data_ptr(), 0, st); torch.cuda.synchronize()
e0.record(); g.pll_process_device(d_in.data_ptr(), 2, fr.shape[1], d_out.data_ptr(), 0, st); e1.record(); torch.cuda.synchronize()
res.append(dict(sequential_device_thread_ms_per_frame=round(e0.elapsed_time(e1) / 2, 3)))
print(res[-1], flush=True)
# 128 streams
blocks, ins, outs = [], [], []
for k in range(128):
    g = pkg.S2PLSyncBlock(360, False); g.pll_set_params(0.004, 4, False, False, 0)
    blocks.append(g); ins.append(torch.from_numpy(np.roll(fr, k, axis=1).view(np.float32).copy()).cuda()); outs.append(torch.zeros_like(ins[-1]))
args = (blocks, [t.data_ptr() for t in ins], 2, fr.shape[1], [t.data_ptr() for t in outs], st)
pkg.pll_process_multi_device(*args); torch.cuda.synchronize()
e0.record(); pkg.pll_process_multi_device(*args); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
res.append(dict(streams=128, frames_per_stream=2, ms=round(ms, 3), frames_per_s=round(256 / ms * 1e3, 1), msym_s=round(256 * 32490 / ms / 1e3, 1)))
print(res[-1], flush=True)
if a.out: json.dump(res, open(a.out, "w"), indent=1)
