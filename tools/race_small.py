import importlib, os, sys
import numpy as np
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sdrpp-dvbs-demodulator_b200")
rng = np.random.default_rng(0)
dec = pkg.DVBS2Decoder(max_batch=8, max_trials=3)
dec.setDemodParams(4, True, False, 3)
n = 2
a = 1 / np.sqrt(2.0)
esn0 = 1.0
sigma2 = 1.0 / (2.0 * 10 ** (esn0 / 10.0))
codes = np.stack([pkg.encode_fecframe(4, True, rng.integers(0, 256, dec.kbch // 8, dtype=np.uint8)) for _ in range(n)])
y = (1.0 - 2.0 * codes.astype(np.float32)) * a + rng.normal(0, np.sqrt(sigma2), codes.shape).astype(np.float32)
llr = np.clip(np.rint(4.0 * 2.0 * a * y / sigma2), -127, 127).astype(np.int8)
bb, res = dec.decode_batch(llr)
print(res["ldpc_iters"].tolist())
dec.close()
