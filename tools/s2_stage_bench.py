#!/usr/bin/env python3
"""dvbs2fec_s2_demod_process (the DVB-S2 decode stage in one call: symbols in host memory -> BBFRAMEs in host memory) on one
stream of PLFRAMEs: frames / s and symbols / s, with the share of the payload phase loop (K8, one warp per stream: a
recurrence) measured beside it.  The other stages are batch kernels; K8 is what a single stream waits for."""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import torch
import plstream

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--frames", type=int, default=24); a = ap.parse_args()
res = []
for name, modcod, short, esn0 in (("QPSK 1/2 normal", 4, False, 6.0), ("8PSK 3/5 normal", 12, False, 10.0), ("QPSK 1/2 short", 4, True, 6.0)):
    rng = np.random.default_rng(modcod)
    info = pkg.modcod_info(modcod, short, False)
    kb = info["kbch"] // 8
    pay = rng.integers(0, 256, (4, kb), dtype=np.uint8)
    pls = modcod << 2 | int(short) << 1
    rn = plstream.pl_rn(0)
    fr = []
    for i in range(4):
        sym = pkg.modulate(modcod, short, False, pkg.encode_fecframe(modcod, short, pay[i])).view(np.complex64).copy()
        body = sym[90:]
        fr.append(np.concatenate([plstream.plheader(pls) * np.abs(body[0]), body * np.array([1, 1j, -1, -1j])[rn[:len(body)]]]))
    x = np.concatenate([fr[i % 4] for i in range(a.frames)] + [fr[0][:500]]).astype(np.complex64)
    x = x * np.exp(1j * (0.2 + 2 * np.pi * 1e-5 * np.arange(len(x))))
    sigma = np.abs(fr[0][100]) * np.sqrt(0.5 / 10 ** (esn0 / 10))
    x = (x + sigma * (rng.normal(size=len(x)) + 1j * rng.normal(size=len(x)))).astype(np.complex64)
    g = pkg.DVBS2DemodStage(max_batch=64)
    g.setDemodParams(modcod, short, False, 25, 0.004, 0.004, 0)
    bb, r, fed, hdr = g.process(x)
    ok = int(sum(any(np.array_equal(b, p) for p in pay) for b in bb))
    g.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        g.reset(); g.process(x)
    ms = (time.perf_counter() - t0) / reps * 1e3
    # the phase loop alone on the same frames (device buffers)
    slots = info["nldpc"] // info["bits"] // 90
    p = pkg.S2PLSyncBlock(slots, False); p.pll_set_params(0.004, modcod, short, False, 0)
    nfr = len(bb)
    d_in = torch.from_numpy(x[:nfr * p.raw_frame_size].view(np.float32).copy()).cuda(); d_out = torch.zeros_like(d_in)
    st = torch.cuda.current_stream().cuda_stream
    p.pll_process_device(d_in.data_ptr(), nfr, p.raw_frame_size, d_out.data_ptr(), 0, st); torch.cuda.synchronize(); t0 = time.perf_counter()
    p.pll_process_device(d_in.data_ptr(), nfr, p.raw_frame_size, d_out.data_ptr(), 0, st); torch.cuda.synchronize()
    pll_ms = (time.perf_counter() - t0) * 1e3
    res.append(dict(workload=name, frames=len(bb), frames_equal_to_sent=ok, symbols=len(x), ms_host_to_host=round(ms, 2), msym_s=round(len(x) / ms / 1e3, 2),
                    frames_per_s=round(len(bb) / ms * 1e3, 1), info_mbit_s=round(len(bb) * info["kbch"] / ms / 1e3, 1), phase_loop_ms=round(pll_ms, 2),
                    phase_loop_share=round(pll_ms / ms, 2), mean_ldpc_iters=float(np.where(r["ldpc_iters"] < 0, 25, r["ldpc_iters"]).mean())))
    print(res[-1], flush=True)
    g.close(); p.close()
# several transponders: a handle, a host thread and a stream each (ctypes releases the GIL for the call); the phase loops of
# different transponders run side by side on the device
import threading
multi = []
for nt in (1, 4, 16):
    hs = []
    for _ in range(nt):
        g = pkg.DVBS2DemodStage(max_batch=64); g.setDemodParams(4, True, False, 25, 0.004, 0.004, 0); g.process(x); g.reset(); hs.append(g)
    def work(g):
        for _ in range(3):
            g.reset(); g.process(x)
    th = [threading.Thread(target=work, args=(g,)) for g in hs]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    multi.append(dict(workload="QPSK 1/2 short, one handle + host thread per transponder", transponders=nt, ms_per_round=round(ms, 2),
                      msym_s_total=round(nt * len(x) / ms / 1e3, 1), frames_per_s_total=round(nt * len(bb) / ms * 1e3)))
    print(multi[-1], flush=True)
    for g in hs: g.close()
res.append(dict(multi_transponder=multi))
if a.out: json.dump(res, open(a.out, "w"), indent=1)
