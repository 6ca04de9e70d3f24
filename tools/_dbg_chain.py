import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, importlib
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
import dvbs_stream, orclib
from test_gpu_vit import dvbs_symbols, OracleChain
rate, stride, seed = 4, 1632, 3
rng = np.random.default_rng(400 + seed)
ts, syms = dvbs_symbols(40, rate, rng, lead=2 * int(rng.integers(0, 50)))
cuts = sorted(set(int(c) for c in rng.integers(0, len(syms), 4)) | {0, len(syms)})
o, g = OracleChain(stride), pkg.DVBSDemod(frame_stride=stride)
for lo, hi in zip(cuts[:-1], cuts[1:]):
    want = o.process(syms[lo:hi]); got = g.process(syms[lo:hi])
    print(lo, hi, want.shape, got.shape, g.stats())
    m = min(len(want), len(got))
    print(" first m equal:", np.array_equal(want[:m], got[:m]), "tail equal:", np.array_equal(want[-m:], got[-m:]) if m else None)
