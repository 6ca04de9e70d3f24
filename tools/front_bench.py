#!/usr/bin/env python3
"""Device time of the stages either side of the decode stage that round 2 added -- GSE extraction (K6, GSE pass) and
the PL front end (K7: PL sync, PLHEADER demodulation, coarse frequency error) -- on device-resident input, with the
reference's own code (oracle/_ref) timed on one host core beside it.  Run on the GPU box."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bbstream  # noqa: E402
import orclib  # noqa: E402
import plstream  # noqa: E402


def timed(fn, reps, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):   # buffers are sized on the first calls
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gse_case(pkg, torch, nframes):
    kbch = 32208
    kb = kbch // 8
    rng = np.random.default_rng(3)
    base = bbstream.random_gse_scenario(rng, kbch, nframes=256)
    reps_in = max(1, nframes // len(base))
    frames = np.concatenate([base] * reps_in)          # the same traffic repeated: reassembly state stays consistent
    n = len(frames)
    dev = torch.device("cuda", 0)
    d_bb = torch.from_numpy(frames).to(dev)
    d_out = torch.zeros(n * kb + 3 * 70000, dtype=torch.uint8, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    g = pkg.BBFrameTSParser()
    g.setFrameSize(kbch)
    st = torch.cuda.current_stream()
    ms = timed(lambda: g.work_device(d_bb.data_ptr(), n, d_out.data_ptr(), d_out.numel(), d_n.data_ptr(), st.cuda_stream), 10, torch)
    produced = int(d_n.item())
    g._stats()
    row = dict(stage="GSE extraction (BBFrameTSParser::work, ts_gs = 01)", frames=n, bytes_in=int(frames.nbytes), bytes_out=produced,
               gpu_ms=round(ms, 4), gpu_gb_s=round((frames.nbytes + produced) / ms / 1e6, 1), counters=g.gse_counters)
    g.close()
    r = orclib.ref() if orclib.have_ref() and hasattr(orclib.ref(), "ref_ts_create") else None
    if r is not None:
        h = r.ref_ts_create(kbch)
        out = np.zeros(n * kb + 3 * 70000, np.uint8)
        buf = np.ascontiguousarray(frames).copy()
        r.ref_ts_work(h, buf, n, out, len(out))
        t0 = time.perf_counter()
        for _ in range(3):
            got = r.ref_ts_work(h, buf, n, out, len(out))
        row["cpu_ms_reference_1_core"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
        row["cpu_bytes_out"] = int(got)
        row["speedup"] = round(row["cpu_ms_reference_1_core"] / ms, 1)
    return row


def plsync_case(pkg, torch, slots, pilots, nfr):
    rng = np.random.default_rng(4)
    pls = (4 << 2) | int(pilots)
    x = plstream.stream(pls, slots, pilots, nfr, rng, esn0_db=6.0, lead=1000, cfo=5e-5)
    dev = torch.device("cuda", 0)
    g = pkg.S2PLSyncBlock(slots, pilots)
    rfs = g.raw_frame_size
    d_x = torch.from_numpy(x.view(np.float32)).to(dev)
    d_out = torch.zeros((nfr + 2) * rfs * 2, dtype=torch.float32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def run():
        g.process_device(d_x.data_ptr(), len(x), d_out.data_ptr(), nfr + 2, d_n.data_ptr(), st.cuda_stream)

    ms = timed(run, 10, torch)
    n = int(d_n.item())
    row = dict(stage="PL sync (S2PLSyncBlock::process)", slots=slots, pilots=bool(pilots), symbols=len(x), frames_out=n,
               gpu_ms=round(ms, 4), gpu_msym_s=round(len(x) / ms / 1e3, 1),
               # per position: 57 differential products (shared) and 57 complex accumulations; bytes: 8 in, 4 metric out + in, frames out
               algorithmic_bytes=int(len(x) * 8 + len(x) * 8 + n * rfs * 8))
    row["gpu_gb_s"] = round(row["algorithmic_bytes"] / ms / 1e6, 1)
    # PLHEADER demodulation + coarse FED on the delivered frames (device buffers)
    d_hdr = torch.zeros(n * 90 * 2, dtype=torch.float32, device=dev)
    d_res = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    d_err = torch.zeros(n, dtype=torch.float32, device=dev)
    L = pkg.lib()
    ms_h = timed(lambda: L.dvbs2fec_plhdr_process_device(g._p, n, d_out.data_ptr(), d_hdr.data_ptr(), d_res.data_ptr(), st.cuda_stream), 5, torch)
    ms_f = timed(lambda: L.dvbs2fec_coarse_fed_device(g._p, n, d_out.data_ptr(), int(pilots), pls, 0, d_err.data_ptr(), st.cuda_stream), 20, torch)
    row["plhdr_gpu_ms"] = round(ms_h, 4)
    row["fed_gpu_ms"] = round(ms_f, 4)
    g.close()
    r = orclib.ref() if orclib.have_ref() and hasattr(orclib.ref(), "ref_plsync_create") else None
    if r is not None:
        h = r.ref_plsync_create(slots, int(pilots))
        out = np.zeros(2 * (len(x) + 2 * rfs), np.float32)
        xf = np.ascontiguousarray(x).view(np.float32)
        t0 = time.perf_counter()
        got = r.ref_plsync_process(h, len(x), xf, out)
        row["cpu_ms_reference_1_core"] = round((time.perf_counter() - t0) * 1e3, 2)
        row["speedup"] = round(row["cpu_ms_reference_1_core"] / ms, 1)
        frames = out[:2 * got].reshape(-1, 2 * rfs)
        hh = r.ref_plhdr_create(0.004)
        o90, res, loop = np.zeros(180, np.float32), np.zeros(3, np.int32), np.zeros(2, np.float32)
        t0 = time.perf_counter()
        for k in range(len(frames)):
            r.ref_plhdr_process(hh, rfs, np.ascontiguousarray(frames[k]), o90, res, loop)
        row["plhdr_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        t0 = time.perf_counter()
        for k in range(len(frames)):
            r.ref_coarse_fed(np.ascontiguousarray(frames[k]), rfs, int(pilots), pls, 0)
        row["fed_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "front_bench.json"))
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    rows = [gse_case(pkg, torch, 4096), plsync_case(pkg, torch, 360, False, 64), plsync_case(pkg, torch, 360, True, 64),
            plsync_case(pkg, torch, 90, False, 256)]
    out = dict(note="device-resident input, CUDA-event times averaged over repeated calls; CPU = the reference's own code "
                    "(oracle/_ref) on one host core, same input", rows=rows)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
