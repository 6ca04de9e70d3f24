"""Row 8(f)-3 on the GPU: dvbs2fec_plsync_* (PL sync, PLHEADER demodulation, coarse frequency error) against the CPU
oracle, which tests/test_plsync_oracle.py pins to the reference's own sources."""
import ctypes as C

import numpy as np
import pytest

import orclib
import plstream
from fec import pkg
from test_plsync_oracle import OrcSync, f32

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.mark.parametrize("case", [(90, False, 16.0, 0), (90, True, 8.0, 1), (36, False, 3.0, 2), (144, True, 1.0, 3), (360, False, 6.0, 4),
                                  (360, True, 2.0, 5)])
def test_sync_matches_oracle(case):
    """frames delivered, in what order and from which positions; current_position and best_match: bit for bit,
    with the symbol stream cut into calls at random places (the gathering state crosses calls)"""
    slots, pilots, esn0, seed = case
    rng = np.random.default_rng(40 + seed)
    x = plstream.stream(11 << 2 | pilots, slots, pilots, 6, rng, esn0_db=esn0, lead=int(rng.integers(1, 3000)), cfo=2e-4 * seed)
    o = OrcSync(slots, pilots)
    g = pkg.S2PLSyncBlock(slots, pilots)
    assert g.raw_frame_size == o.rfs
    cuts = sorted(set(int(c) for c in rng.integers(0, len(x), 9)) | {0, len(x)})
    for a, b in zip(cuts[:-1], cuts[1:]):
        want, got = o.process(x[a:b]), g.process(x[a:b])
        assert len(want) == len(got) and np.array_equal(bits(want), bits(got))
        rfs, cur, best = o.stats()
        assert (g.current_position, g.best_match) == (cur, best)
    g.close()


def test_sync_on_noise_and_tiny_calls_matches_oracle():
    rng = np.random.default_rng(9)
    x = ((rng.normal(size=30000) + 1j * rng.normal(size=30000)) / np.sqrt(2)).astype(np.complex64)
    o, g = OrcSync(36, False), pkg.S2PLSyncBlock(36, False)
    at = 0
    for n in [1, 0, 2, 88, 89, 90, 91, 3329, 3330, 3331, 1, 7000, 5000, 4000]:
        want, got = o.process(x[at:at + n]), g.process(x[at:at + n])
        assert np.array_equal(bits(want), bits(got))
        assert (g.current_position, g.best_match) == o.stats()[1:]
        at += n
    g.reset()
    o2 = OrcSync(36, False)
    assert np.array_equal(bits(o2.process(x[:9000])), bits(g.process(x[:9000])))
    g.close()


def test_sync_full_size_block_on_device_buffers():
    """2 M symbols (64 normal frames, pilots on) in three device-buffer calls: every frame behind the first lock is
    a verbatim PLFRAME of the input, consecutive and in order; equal to the host-buffer call"""
    import torch
    rng = np.random.default_rng(12)
    slots, pilots, nfr = 360, True, 64
    x = plstream.stream((4 << 2) | 1, slots, pilots, nfr, rng, esn0_db=5.0, lead=4321, cfo=1e-4)
    g = pkg.S2PLSyncBlock(slots, pilots)
    rfs = g.raw_frame_size
    dev = torch.device("cuda", 0)
    d_x = torch.from_numpy(x.view(np.float32)).to(dev)
    d_out = torch.zeros((nfr + 2) * rfs * 2, dtype=torch.float32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    got = []
    cuts = [0, 50001, 50002, 1000000, len(x)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g.process_device(d_x[2 * a:].data_ptr(), b - a, d_out.data_ptr(), nfr + 2, d_n.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        n = int(d_n.item())
        got.append(d_out[:2 * n * rfs].cpu().numpy().view(np.complex64).copy())
    y = np.concatenate(got).reshape(-1, rfs)
    h = pkg.S2PLSyncBlock(slots, pilots)
    assert np.array_equal(bits(h.process(x)), bits(y.reshape(-1)))
    # the first window finds the PLHEADER 4321 symbols in and is realigned onto it: every delivered frame is a PLFRAME
    assert len(y) >= nfr - 1
    for i in range(len(y)):
        assert np.array_equal(y[i], x[4321 + i * rfs: 4321 + (i + 1) * rfs])
    g.close(); h.close()


@pytest.mark.parametrize("pls", [4 << 2, (13 << 2) | 1, (28 << 2) | 2, (6 << 2) | 3])
def test_plhdr_demod_matches_oracle(pls):
    """header symbols and loop state to float rounding (device sin/cos), PLS fields exactly"""
    rng = np.random.default_rng(pls)
    slots, pilots = 90, bool(pls & 1)
    x = plstream.stream(pls, slots, pilots, 12, rng, esn0_db=7.0, lead=0, cfo=3e-5, phase=0.4)
    o = orclib.oracle()
    g = pkg.S2PLSyncBlock(slots, pilots, loop_bw=0.004)
    rfs = g.raw_frame_size
    ho = o.orc_plhdr_create(0.004)
    frames = x[:12 * rfs].reshape(12, rfs)
    want_h, want_r, want_l = np.zeros((12, 90), np.complex64), np.zeros((12, 3), np.int32), np.zeros((12, 2), np.float32)
    for k in range(12):
        o.orc_plhdr_process(ho, rfs, f32(frames[k]), want_h[k].view(np.float32), want_r[k], want_l[k])
    # in two calls: the loop state crosses calls
    h1, r1, _ = g.plhdr_process(frames[:5])
    h2, r2, loop = g.plhdr_process(frames[5:])
    got_h, got_r = np.concatenate([h1, h2]), np.concatenate([r1, r2])
    assert np.allclose(got_h, want_h, atol=1e-4)
    assert np.allclose(loop, want_l[-1], atol=1e-4)
    assert np.array_equal(got_r[:, :3], want_r)
    assert np.array_equal(got_r[:, 3], (want_r[:, 0] << 2) | (want_r[:, 1] << 1) | want_r[:, 2])
    o.orc_plhdr_destroy(ho)
    g.close()


@pytest.mark.parametrize("case", [(90, False, 0), (90, True, 1), (360, True, 2), (36, True, 3)])
def test_coarse_fed_matches_oracle(case):
    slots, pilots, seed = case
    rng = np.random.default_rng(70 + seed)
    pls = (5 << 2) | int(pilots)
    codenum = [0, 1, 7, 262141][seed]
    x = plstream.stream(pls, slots, pilots, 9, rng, esn0_db=10.0, lead=0, cfo=1e-3 * (seed + 1), codenum=codenum)
    g = pkg.S2PLSyncBlock(slots, pilots)
    rfs = g.raw_frame_size
    rn = plstream.pl_rn(codenum)
    frames = x[:9 * rfs].reshape(9, rfs)
    want = np.array([orclib.oracle().orc_coarse_fed(f32(frames[k]), rfs, int(pilots), pls, rn) for k in range(9)], np.float32)
    got = g.coarse_fed(frames, pilots, pls, codenum)
    assert np.array_equal(bits(want), bits(got))
    g.close()
