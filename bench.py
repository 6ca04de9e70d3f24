#!/usr/bin/env python3
"""Benchmark of the DVB-S2 decode stage (BASELINE.json metric: decoded information Gbit/s, QPSK 1/2 normal).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--esn0 DB] [--pool F]

One "step" = one pass of the hot path (LDPC + BCH + descramble, from int8 LLRs) over a pool of F synthetic
FECFRAMEs.  Per rank the pool is resident in HBM and larger than L2 (F*64800 B = 265 MB at F = 4096), so
consecutive steps never find their input in cache.  `value` is timed with CUDA events on the stream the
kernels run on; `e2e` goes through dvbs2fec_decode_batch with pinned HOST buffers (H2D + kernels + D2H).
Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank decodes its own pool, no
data-path collective; timing is barrier-bracketed and the max over ranks is used.

`--impl reference` times the reference's own CPU code (oracle/_ref, compiled from /root/reference
unmodified; else the oracle port) on the host cores with the same pool generator and metric.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODCOD, SHORT, RATE_ENUM = 4, False, 3  # QPSK 1/2 normal
MAX_TRIALS = 25
LLR_SCALE = 4.0  # "L4" generator of SURVEY.md 8d: rint(4 * true channel LLR), clamped to +-127
HBM_BYTES_PER_FRAME = 64800 + 32208 // 8 + 16  # algorithmic: LLRs in, BBFRAME + result record out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs"), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic_per_frame():
    """dram__bytes_read+write per frame of the LDPC kernel from the newest committed ncu capture (profiles/)."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ldpc_pair_kernel.json"))):
        try:
            d = json.load(open(f))
            if "dram_traffic_bytes_per_frame" in d:
                best = (d["dram_traffic_bytes_per_frame"], os.path.basename(f),
                        {k: d.get(k) for k in ("alu_pipe_pct", "issue_slots_busy_pct", "fma_pipe_pct", "lsu_pipe_pct", "dram_throughput_pct")})
        except Exception:
            pass
    return best


def make_codewords(pkg, n, seed):
    rng = np.random.default_rng(seed)
    info = pkg.modcod_info(MODCOD, SHORT)
    out = np.zeros((n, info["nldpc"]), np.uint8)
    for i in range(n):
        out[i] = pkg.encode_fecframe(MODCOD, SHORT, rng.integers(0, 256, info["kbch"] // 8, dtype=np.uint8))
    return out


def llr_params(esn0_db):
    a = 1.0 / np.sqrt(2.0)
    sigma2 = 1.0 / (2.0 * 10 ** (esn0_db / 10.0))
    return a, sigma2


def make_pool_numpy(code_bits, nframes, esn0_db, seed):
    rng = np.random.default_rng(seed)
    a, sigma2 = llr_params(esn0_db)
    out = np.zeros((nframes, code_bits.shape[1]), np.int8)
    for i in range(nframes):
        y = (1.0 - 2.0 * code_bits[i % len(code_bits)].astype(np.float32)) * a
        y = y + rng.normal(0.0, np.sqrt(sigma2), y.shape).astype(np.float32)
        out[i] = np.clip(np.rint(LLR_SCALE * 2.0 * a * y / sigma2), -127, 127).astype(np.int8)
    return out


def make_pool_torch(torch, code_bits, nframes, esn0_db, seed, device):
    a, sigma2 = llr_params(esn0_db)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cw = torch.from_numpy(code_bits).to(device)
    pool = torch.empty((nframes, cw.shape[1]), dtype=torch.int8, device=device)
    step = 256
    for f0 in range(0, nframes, step):
        m = min(step, nframes - f0)
        idx = (torch.arange(f0, f0 + m, device=device) % cw.shape[0])
        y = (1.0 - 2.0 * cw[idx].float()) * a
        y += torch.randn(y.shape, generator=g, device=device) * float(np.sqrt(sigma2))
        pool[f0:f0 + m] = torch.clamp(torch.round(LLR_SCALE * 2.0 * a * y / sigma2), -127, 127).to(torch.int8)
    return pool


class ClockSampler:
    """nvidia-smi in loop mode (one process, 50 ms period) while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def start(self):
        time.sleep(0.15)   # let the first samples arrive before the timed region

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        for ln in out.splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU reference arm
_cpu_state = {}


def _cpu_worker_init(kind):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    _cpu_state["orclib"] = orclib
    _cpu_state["kind"] = kind
    _cpu_state["lib"] = orclib.ref() if kind == "reference" else orclib.oracle()


def _cpu_worker(job):
    """job: (n, 64800) int8 LLRs -> (frames done, decoded-ok count).  'reference': the vendored library used
    as designed (one frame per SSE lane, blocks = lanes) + per-frame BBFrameBCH::decode + descrambler."""
    llr = job
    orclib, lib, kind = _cpu_state["orclib"], _cpu_state["lib"], _cpu_state["kind"]
    n = llr.shape[0]
    ok = 0
    if kind == "reference":
        lanes = 16
        for f0 in range(0, n, lanes):
            blk = np.ascontiguousarray(llr[f0:f0 + lanes])
            if blk.shape[0] < lanes:
                blk = np.concatenate([blk, np.repeat(blk[-1:], lanes - blk.shape[0], 0)])
            lib.ref_ldpc_decode_simd(0, RATE_ENUM, blk.reshape(-1), MAX_TRIALS)
            for k in range(min(lanes, n - f0)):
                packed = np.packbits((blk[k, :32400] < 0).astype(np.uint8))
                c = lib.ref_bch_decode(0, RATE_ENUM, packed)
                lib.ref_descramble(0, RATE_ENUM, packed)
                ok += c >= 0
    else:
        bb = np.zeros(32208 // 8, np.uint8)
        for k in range(n):
            it, co = C.c_int(), C.c_int()
            lib.orc_decode_frame(0, RATE_ENUM, llr[k].copy(), MAX_TRIALS, bb, C.byref(it), C.byref(co))
            ok += co.value >= 0
    return n, ok


def cpu_baseline(pool_host, budget_s, procs):
    """Bounded sample of the same workload on the host cores.  Returns dict for the JSON line."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    kind = "reference" if orclib.have_ref() else "port"
    per_job = 16
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_cpu_worker_init, initargs=(kind,)) as pool:
        # calibrate on one job per process, then size the sample to the budget
        jobs = [pool_host[(i * per_job) % len(pool_host):][:per_job] for i in range(procs)]
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
        rounds = max(1, min(64, int(budget_s / max(dt, 1e-3))))
        jobs = [pool_host[(i * per_job) % (len(pool_host) - per_job + 1):][:per_job] for i in range(procs * rounds)]
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    return {"value": frames * 32208 / dt / 1e9, "unit": "Gbit/s", "cores": procs, "kind": kind,
            "frames_per_s": frames / dt, "seconds": dt,
            "sample": "%d frames of the QPSK 1/2 pool, %d processes, LDPC 16 frames per SSE4.1 call (blocks=16) + "
                      "per-frame BCH + descramble" % (frames, procs) if kind == "reference" else
                      "%d frames of the QPSK 1/2 pool, %d processes, scalar oracle port" % (frames, procs)}


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--esn0", type=float, default=2.2)
    ap.add_argument("--pool", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
    info = pkg.modcod_info(MODCOD, SHORT)
    kbch, N = info["kbch"], info["nldpc"]
    workload = ("DVB-S2 QPSK 1/2 normal FECFRAME (64800) LDPC+BCH decode, synthetic AWGN int8 LLRs "
                "(rint(4*LLR)) at Es/N0 %.1f dB, %d max iterations" % (args.esn0, MAX_TRIALS))
    config = {"workload": workload, "frames_per_step_per_gpu": args.pool, "esn0_db": args.esn0, "llr_generator": "L4",
              "max_iters": MAX_TRIALS, "cache": "pool of %d frames = %.0f MB per GPU > 126 MB L2" % (args.pool, args.pool * N / 1e6),
              "parallelism": "frames sharded by GPU, no collective"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        codes = make_codewords(pkg, 32, 1)
        procs = os.cpu_count() or 1
        pool_host = make_pool_numpy(codes, 16 * min(procs, 64), args.esn0, 2)
        t_all, frames_all = 0.0, 0
        last = None
        for s in range(args.warmup + args.steps):
            last = cpu_baseline(pool_host, max(2.0, args.cpu_seconds / 2), procs)
            if s >= args.warmup:
                t_all += last["seconds"]
                frames_all += int(round(last["frames_per_s"] * last["seconds"]))
            if s >= args.warmup and t_all > 120:
                break
        v = frames_all * kbch / t_all / 1e9
        line = {"impl": "reference", "metric": "decoded_info_gbit_s", "value": v, "unit": "Gbit/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
                "config": config, "frames_per_s": frames_all / t_all,
                "cpu_baseline": {"value": v, "unit": "Gbit/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
                "e2e": {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode stage has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    dec = pkg.DVBS2Decoder(devices=[local_rank], max_batch=args.pool, max_trials=MAX_TRIALS)
    dec.setDemodParams(MODCOD, SHORT, False, MAX_TRIALS)
    codes = make_codewords(pkg, 64, 1 + rank)
    pool = make_pool_torch(torch, codes, args.pool, args.esn0, 100 + rank, dev)
    d_bb = torch.empty((args.pool, kbch // 8), dtype=torch.uint8, device=dev)
    d_res = torch.empty((args.pool, 16), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        dec.decode_batch_device(pool.data_ptr(), args.pool, d_bb.data_ptr(), d_res.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches_per_step = dec.last_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dec.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    demap_ms, ldpc_ms, bch_ms, nspans = dec.kernel_times()
    dec.set_profiling(False)
    res = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(-1)
    iters = res["ldpc_iters"].astype(np.int32)
    mean_it = float(np.where(iters < 0, MAX_TRIALS, iters).mean())
    fer = float((res["bch_corr"] < 0).mean())

    # ---- e2e: public host API, pinned host buffers, copies inside the timed region
    e2e_frames = 2 * args.pool   # 4 chunks of 2048 frames, double-buffered H2D / kernels / D2H
    L = pkg.lib()
    h_in = L.dvbs2fec_alloc_pinned(e2e_frames * N)
    h_bb = L.dvbs2fec_alloc_pinned(e2e_frames * (kbch // 8))
    h_res = L.dvbs2fec_alloc_pinned(e2e_frames * 16)
    host_copy = pool.cpu().numpy()
    C.memmove(h_in, host_copy.ctypes.data, args.pool * N)
    C.memmove(h_in + args.pool * N, host_copy.ctypes.data, args.pool * N)
    dec_e = pkg.DVBS2Decoder(devices=[local_rank], max_batch=max(256, args.pool // 2), max_trials=MAX_TRIALS)
    dec_e.setDemodParams(MODCOD, SHORT, False, MAX_TRIALS)
    for _ in range(2):
        dec_e.decode_batch_raw(h_in, e2e_frames, h_bb, h_res)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for _ in range(e2e_steps):
        dec_e.decode_batch_raw(h_in, e2e_frames, h_bb, h_res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop()

    # ---- per-frame latency vs batch size (device-resident, one call, median of 5)
    latency = {}
    for b in (2, 16, 128, 1024, args.pool):
        if b > args.pool:
            continue
        ts = []
        for _ in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            dec.decode_batch_device(pool.data_ptr(), b, d_bb.data_ptr(), d_res.data_ptr(), stream.cuda_stream)
            a1.record(stream)
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        latency[str(b)] = round(sorted(ts)[2], 4)

    # ---- reduce over ranks
    dispatch = importlib.import_module("sdrpp-dvbs-demodulator_b200.dispatch")
    ms_max, e2e_ms_max, ldpc_ms_max = dispatch.reduce_max([ms, e2e_s * 1e3, ldpc_ms], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    frames_total = args.pool * args.steps * world
    value = frames_total * kbch / (ms_max * 1e-3) / 1e9
    e2e_value = e2e_frames * e2e_steps * world * kbch / (e2e_ms_max * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    ldpc_launches = args.steps * ((args.pool + args.pool - 1) // args.pool)
    ldpc_avg_ms = ldpc_ms_max / max(1, ldpc_launches)
    achieved = args.pool * HBM_BYTES_PER_FRAME / (ldpc_avg_ms * 1e-3) / 1e9
    links = info["links_total"]
    tr = ncu_traffic_per_frame()
    line = {
        "metric": "decoded_info_gbit_s", "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": config,
        "frames_per_s": frames_total / (ms_max * 1e-3), "mean_ldpc_iters": mean_it, "fer": fer,
        "gpu_launches": launches_per_step * args.steps,
        "kernel_ms_per_step": {"ldpc_pair_kernel": ldpc_ms / args.steps, "bch_kernel": bch_ms / args.steps},
        "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": e2e_frames * N,
                "d2h_bytes_per_step": e2e_frames * (kbch // 8 + 16), "frames_per_step": e2e_frames, "steps": e2e_steps,
                "api": "dvbs2fec_decode_batch (pinned host buffers)"},
        "roofline": {"kernel": "ldpc_pair_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": (tr[0] * args.pool / 1e9) if tr else None,
                     "traffic_unit": "GB per launch (ncu dram__bytes_read+write, %s)" % (tr[1] if tr else "n/a"), "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": HBM_BYTES_PER_FRAME,
                     "edge_updates_per_s": args.pool / (ldpc_avg_ms * 1e-3) * links * mean_it,
                     "note": "the kernel is bound by integer issue and shared memory, not HBM (DESIGN.md); HBM fraction is reported as the contract asks",
                     "on_chip_utilisation_pct_ncu": (tr[2] if tr else None)},
        "latency_ms_per_batch": latency,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu and world == 1:   # the CPU reference is timed beside the N=1 run only
        codes_h = codes
        procs = os.cpu_count() or 1
        pool_host = make_pool_numpy(codes_h, 16 * min(procs, 64), args.esn0, 2)
        cb = cpu_baseline(pool_host, args.cpu_seconds, procs)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
