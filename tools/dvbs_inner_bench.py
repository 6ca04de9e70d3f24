#!/usr/bin/env python3
"""K10 / K11 and the whole DVB-S decode stage against the reference's own classes on one host core (oracle/_ref; the
oracle port where it is absent):
  viterbi   per rate: blocks of 8192 soft bits per second on device buffers (locked, steady state), decoded Mbit/s, the
            reference's time for the same blocks; the lock search on noise (52 candidate decodes per block)
  deframer  bits per second
  chain     symbols in host memory -> TS packets in host memory (dvbs2fec_dvbs_demod_process), symbols per second,
            against sts + viterbi + deframer + outer decoder of the reference run one after the other"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import torch
import dvbs_stream, orclib
from test_vit_oracle import OrcViterbi, RefViterbi
from test_dvbs_oracle import OrcDeframer, RefDeframer, OrcOuter, RefOuter

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--blocks", type=int, default=1056); a = ap.parse_args()
have_ref = orclib.have_ref() and hasattr(orclib.ref(), "ref_vit_create")
Vit, Def, Out = (RefViterbi, RefDeframer, RefOuter) if have_ref else (OrcViterbi, OrcDeframer, OrcOuter)
res = dict(cpu_kind="reference" if have_ref else "oracle", viterbi=[], search=None, deframer=None, chain=[])

def timed(fn, reps):
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.05:      # the host work between the sections lets the GPU clock down: warm up first
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

for rate in range(5):
    rng = np.random.default_rng(rate)
    per_block = [4096, 5462, 6144, 6827, 7168][rate]
    base = dvbs_stream.inner_softs(rng.integers(0, 2, per_block * 68, dtype=np.uint8), rate, rng, sigma=[14.0, 8.0, 8.0, 5.0, 5.0][rate])
    base = base[:66 * 8192]      # 66 blocks: the depuncturers' phase (periods of 3 and 6 soft bits) is the same at every joint
    s = np.tile(base, a.blocks // 66)      # (the joints are wrong code words: a few bad blocks, far below max_outsync)
    n = len(s)
    d_in = torch.from_numpy(s).cuda(); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    g = pkg.DVBSViterbi()
    nbits = g.process_device(d_in.data_ptr(), n, d_out.data_ptr())      # locks on the first block
    first = d_out.cpu().numpy()[:nbits].copy()
    ms = timed(lambda: g.process_device(d_in.data_ptr(), n, d_out.data_ptr()), 5)
    st = g.stats()
    cpu = Vit(); nc = 66 * 8192
    want = cpu.process(s[:nc]); t0 = time.perf_counter(); cpu.process(s[nc:2 * nc]); cpu_ms = (time.perf_counter() - t0) * 1e3 * n / nc
    res["viterbi"].append(dict(rate=dvbs_stream.RATES[rate], blocks=n // 8192, gpu_ms=round(ms, 3), blocks_per_s=round(n / 8192 / ms * 1e3),
                               decoded_mbit_s=round(nbits / ms / 1e3, 1), soft_bits_gb_s=round(n / ms / 1e6, 2), cpu_ms_1_core=round(cpu_ms, 1),
                               speedup=round(cpu_ms / ms, 1), locked=st[1] == 1 and st[2] == rate, equal_to_cpu_on_first_blocks=bool(np.array_equal(first[:len(want)], want)),
                               counters_tasks_repeated_passes=g.counters()))
    print(res["viterbi"][-1], flush=True)
    g.close()

# call size: a task is a warp, so the kernel needs a few thousand blocks in a call to fill the GPU
res["viterbi_call_size"] = []
rng = np.random.default_rng(0)
base = dvbs_stream.inner_softs(rng.integers(0, 2, 4096 * 68, dtype=np.uint8), 0, rng, sigma=14.0)[:66 * 8192]
for blocks in (66, 264, 528, 1056, 2112, 4224, 8184):
    s = np.tile(base, blocks // 66); n = len(s)
    d_in = torch.from_numpy(s).cuda(); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    g = pkg.DVBSViterbi(); g.process_device(d_in.data_ptr(), n, d_out.data_ptr())
    ms = timed(lambda: g.process_device(d_in.data_ptr(), n, d_out.data_ptr()), 5)
    res["viterbi_call_size"].append(dict(rate="1/2", blocks=blocks, gpu_ms=round(ms, 3), blocks_per_s=round(blocks / ms * 1e3), decoded_mbit_s=round(blocks * 4096 / ms / 1e3, 1)))
    print(res["viterbi_call_size"][-1], flush=True); g.close()

# the search on noise: nothing locks, every block runs the 52 candidates
rng = np.random.default_rng(9)
nb = 256
s = np.clip(np.rint(rng.normal(0, 40, nb * 8192)), -127, 127).astype(np.int8)
d_in = torch.from_numpy(s).cuda(); d_out = torch.zeros(len(s), dtype=torch.uint8, device="cuda")
g = pkg.DVBSViterbi(); g.process_device(d_in.data_ptr(), len(s), d_out.data_ptr())
ms = timed(lambda: g.process_device(d_in.data_ptr(), len(s), d_out.data_ptr()), 3)
cpu = Vit(); nc = 16 * 8192; t0 = time.perf_counter(); cpu.process(s[:nc]); cpu_ms = (time.perf_counter() - t0) * 1e3 * len(s) / nc
res["search"] = dict(blocks=nb, gpu_ms=round(ms, 2), blocks_per_s=round(nb / ms * 1e3), msym_s=round(nb * 4096 / ms / 1e3, 1), cpu_ms_1_core=round(cpu_ms, 1),
                     speedup=round(cpu_ms / ms, 1), still_searching=g.stats()[1] == 0)
print(res["search"], flush=True); g.close()

# deframer
from test_dvbs_oracle import deframer_bits
frames, bits = deframer_bits(600, np.random.default_rng(3), lead=4321)
d_bits = torch.from_numpy(bits).cuda(); d_fr = torch.zeros(700 * 1632, dtype=torch.uint8, device="cuda")
g = pkg.DVBSTSDeframer(); st = torch.cuda.current_stream().cuda_stream
ms = timed(lambda: g.work_device(d_bits.data_ptr(), len(bits), d_fr.data_ptr(), 700, 0, st), 10)
cpu = Def(); nc = 40 * 13056; t0 = time.perf_counter(); cpu.work(bits[:nc]); cpu_ms = (time.perf_counter() - t0) * 1e3 * len(bits) / nc
res["deframer"] = dict(bits=len(bits), gpu_ms=round(ms, 4), gbit_s=round(len(bits) / ms / 1e6, 2), cpu_ms_1_core=round(cpu_ms, 1), speedup=round(cpu_ms / ms, 1))
print(res["deframer"], flush=True); g.close()

# whole stage, host to host
from test_gpu_vit import dvbs_symbols
o = orclib.ref() if have_ref else orclib.oracle()
for rate in (0, 2, 4):
    rng = np.random.default_rng(20 + rate)
    ts, syms = dvbs_symbols(600, rate, rng)
    g = pkg.DVBSDemod(frame_stride=1632)
    got = g.process(syms); g.reset()
    ms = timed(lambda: (g.reset(), g.process(syms)), 3)
    # the reference's blocks one after the other on one core
    x = np.ascontiguousarray(syms).view(np.float32).reshape(-1); n = len(syms)
    sts = (o.ref_sts_create if have_ref else o.orc_sts_create)(); soft = np.zeros(2 * n + 8192, np.int8)
    vit, de, ou = Vit(), Def(), Out()
    t0 = time.perf_counter()
    k = (o.ref_sts_process if have_ref else o.orc_sts_process)(sts, n, x, soft)
    bits_ = vit.process(soft[:k]); fr, _ = de.work(bits_); want, _ = ou.process(fr.reshape(-1), len(fr), 1632)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    sent = {p.tobytes() for p in ts}
    res["chain"].append(dict(rate=dvbs_stream.RATES[rate], symbols=n, ts_packets=len(got), gpu_ms_host_to_host=round(ms, 2), msym_s=round(n / ms / 1e3, 1),
                             ts_mbit_s=round(len(got) * 1504 / ms / 1e3, 1), cpu_ms_1_core=round(cpu_ms, 1), speedup=round(cpu_ms / ms, 1),
                             equal_to_cpu=bool(got.shape == want.shape and np.array_equal(got, want)),
                             packets_recovered=int(sum(p.tobytes() in sent for p in got))))
    print(res["chain"][-1], flush=True); g.close()
if a.out: json.dump(res, open(a.out, "w"), indent=1)
