#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own code compiled unmodified (oracle/_ref).

The reference ships no tests or vectors for its decode stage; these fixtures are
outputs of the reference itself, run in the build container:

    make -C oracle ref && python tools/make_golden.py

Each fixture holds seeded inputs and what the reference produced for them.  tests/test_golden.py checks
the C oracle against them on the CPU and the CUDA library against them on the GPU box (where
/root/reference does not exist)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
import bbstream  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SNR = {0: -2.0, 1: -1.0, 2: 0.0, 3: 1.3, 4: 2.6, 5: 3.4, 6: 4.3, 7: 5.0, 8: 5.5, 10: 6.5, 11: 6.7}


def ldpc_chain_fixture(short, rate, nframes, seed):
    r = orclib.ref()
    p = orclib.code_params(short, rate)
    rng = np.random.default_rng(seed)
    llr = np.zeros((nframes, p["N"]), np.int8)
    for i in range(nframes):
        _, code = orclib.encode_frame(short, rate, rng)
        llr[i] = orclib.awgn_llr(code, SNR[rate] + (0.4 if short else 0.0) + 0.9 * (i / max(1, nframes - 1) - 0.4), rng)
    llr[0, ::131] = 0
    post = llr.copy()
    iters = np.zeros(nframes, np.int16)
    corr = np.zeros(nframes, np.int16)
    bb = np.zeros((nframes, p["kbch"] // 8), np.uint8)
    for i in range(nframes):
        iters[i] = r.ref_ldpc_decode(short, rate, post[i], 25)
        packed = np.packbits((post[i, : p["K"]] < 0).astype(np.uint8))
        corr[i] = r.ref_bch_decode(short, rate, packed)
        r.ref_descramble(short, rate, packed)
        bb[i] = packed[: p["kbch"] // 8]
    return dict(short=short, rate=rate, max_trials=25, llr=llr, post_sha=np.frombuffer(
        b"".join(hashlib.sha256(post[i].tobytes()).digest() for i in range(nframes)), np.uint8).reshape(nframes, 32),
        iters=iters, corr=corr, bb=bb)


def bch_fixture(short, rate, seed):
    r = orclib.ref()
    p = orclib.code_params(short, rate)
    rng = np.random.default_rng(seed)
    K, kbch, t = p["K"], p["kbch"], p["t"]
    nerrs = [0, 1, 2, 3, t - 1, t, t + 1, t + 4, 60]
    frames = np.zeros((len(nerrs), K // 8), np.uint8)
    for i, ne in enumerate(nerrs):
        frames[i, : kbch // 8] = rng.integers(0, 256, kbch // 8, dtype=np.uint8)
        r.ref_bch_encode(short, rate, frames[i])
        for e in rng.choice(K, ne, replace=False):
            frames[i, e >> 3] ^= 0x80 >> (e & 7)
    out = frames.copy()
    corr = np.array([r.ref_bch_decode(short, rate, out[i]) for i in range(len(nerrs))], np.int16)
    return dict(short=short, rate=rate, frames=frames, corrected=out, corr=corr)


def demap_fixture(ctype, const, short, rate, g1, g2, seed):
    r = orclib.ref()
    bits = {1: 2, 3: 3, 4: 4, 5: 5}[ctype]
    n = 16200 if short else 64800
    nsym = n // bits
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, 1 << bits, nsym).astype(np.uint8)
    pts = np.zeros(2 * nsym, np.float32)
    r.ref_mod(ctype, g1, g2, labels, nsym, pts)
    sym = pts + rng.normal(0, 0.08, pts.shape).astype(np.float32)
    soft = np.zeros(nsym * bits, np.int8)
    if ctype == 5:
        r.ref_demap_calc(ctype, g1, g2, sym, nsym, soft)
    else:
        r.ref_demap(ctype, g1, g2, sym, nsym, soft)
    out = np.zeros(n, np.int8)
    r.ref_deinterleave(const, short, rate, soft.copy(), out)
    pl = np.zeros(2 * (90 + nsym), np.float32)
    pl[180:] = sym
    return dict(ctype=ctype, const=const, short=short, rate=rate, g1=g1, g2=g2, plframe=pl, llr=out)


def ts_parser_fixture(kind, kbch, seed):
    """BBFrameTSParser::work over a stream cut into three calls; outputs concatenated, lengths + stats per call"""
    import ctypes as C
    rng = np.random.default_rng(seed)
    if kind == "ts_odd":
        frames = bbstream.odd_ts_scenario(rng, kbch)
        cuts = [0, 60, 61, len(frames)]
    elif kind == "ts":
        pk = bbstream.ts_packets(260, rng)
        frames, _ = bbstream.ts_bbframes(kbch, pk, first_byte=101)
        cuts = [0, 3, 4, len(frames)]
    elif kind == "gse_mixed":   # random GSE traffic between TS frames and sync losses
        frames = bbstream.random_gse_scenario(rng, kbch, nframes=24, ts_every=4)
        cuts = [0, 5, 6, 19, len(frames)]
    else:
        frames = bbstream.gse_bbframes(kbch, bbstream.gse_scenario(rng))
        cuts = [0, 2, 3, len(frames)]
    r = orclib.ref()
    h = r.ref_ts_create(kbch)
    outs, lens, stats = [], [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        out = np.zeros(65536 * 10 + 4096, np.uint8)
        n = r.ref_ts_work(h, np.ascontiguousarray(frames[a:b]).copy(), b - a, out, 65536 * 10)
        f = np.zeros(11, np.int32)
        x, y, z = C.c_int(), C.c_int(), C.c_int()
        r.ref_ts_stats(h, f, C.byref(x), C.byref(y), C.byref(z))
        outs.append(out[:n].copy())
        lens.append(n)
        stats.append(list(f) + [x.value, y.value, z.value])
    return dict(kbch=kbch, frames=frames, cuts=np.array(cuts), out=np.concatenate(outs), out_len=np.array(lens),
                stats=np.array(stats, np.int32))


def plfront_fixture(seed):
    """the reference's PL front end (dvbs2_pl_sync.cpp, dvbs2_plhdr_demod.cpp, dvbs2_fed.h, dvbs2_pll.cpp) on a stream of
    five short 32APSK 8/9 PLFRAMEs (36 slots) with pilots behind 777 symbols of junk, Gold code 5, a carrier offset: frames as PL sync
    delivers them over two calls, and per frame the PLHEADER demodulator's output, the coarse frequency error and the
    payload phase loop's output and state"""
    import plstream
    from test_plsync_oracle import RefSync, f32
    import test_pll_oracle
    from test_pll_oracle import RefPll
    rng = np.random.default_rng(seed)
    slots, pilots, codenum, modcod = 36, True, 5, 27
    pls = (modcod << 2) | 2 | 1
    x = plstream.stream(pls, slots, pilots, 5, rng, esn0_db=16.0, lead=777, cfo=1e-4, phase=0.3, codenum=codenum, bits=3)
    test_pll_oracle.CONST["32apsk89"] = (5, 5, 2.54, 4.33)      # MODCOD 27's ring ratios (modcod_to_cfg.cpp:124-125)
    r = orclib.ref()
    s = RefSync(slots, pilots)
    cut = 9000
    y1 = s.process(x[:cut]); st1 = s.stats()
    y2 = s.process(x[cut:]); st2 = s.stats()
    rfs = s.rfs
    fr = np.concatenate([y1, y2]).reshape(-1, rfs)
    n = len(fr)
    hh = r.ref_plhdr_create(0.004)
    hdr, res, loop, fed = np.zeros((n, 180), np.float32), np.zeros((n, 3), np.int32), np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    pll = RefPll(0.004, "32apsk89", slots, pilots, pls, codenum)
    pout, pst = np.zeros((n, pll.total), np.complex64), np.zeros((n, 3), np.float32)
    for k in range(n):
        r.ref_plhdr_process(hh, rfs, f32(fr[k]), hdr[k], res[k], loop[k])
        fed[k] = r.ref_coarse_fed(f32(fr[k]), rfs, int(pilots), pls, codenum)
        pout[k], pst[k] = pll.process(fr[k])
    return dict(slots=slots, pilots=int(pilots), codenum=codenum, pls=pls, modcod=modcod, x=x, cut=cut, nsym=np.array([len(y1), len(y2)]),
                sync_stats=np.array([st1, st2], np.float64), nframes=n,
                frames_sha=np.frombuffer(hashlib.sha256(np.ascontiguousarray(fr).tobytes()).digest(), np.uint8), hdr=hdr, hdr_res=res, hdr_loop=loop, fed=fed, pll_out=pout, pll_state=pst)


def dvbs_vit_fixture(rate, phase, lead, seed):
    """the reference's Viterbi_DVBS (viterbi_all.cpp) on two blocks of noise, then a signal: calls of 2, 1, 3, 1 and 5 blocks;
    decoded bits (packed) and ber / state / rate / phase / shift / invalid after every call"""
    import dvbs_stream
    from test_vit_oracle import RefViterbi
    rng = np.random.default_rng(seed)
    per = [4096, 5462, 6144, 6827, 7168][rate]
    sig = dvbs_stream.inner_softs(rng.integers(0, 2, per * 11, dtype=np.uint8), rate, rng, sigma=[12.0, 9.0, 9.0, 6.0, 5.0][rate], phase=phase, lead=lead)
    softs = np.concatenate([np.clip(np.rint(rng.normal(0, 40, 2 * 8192)), -127, 127).astype(np.int8), sig[:10 * 8192]])
    v = RefViterbi()
    calls, bits, nbits, stats = [2, 1, 3, 1, 5], [], [], []
    k = 0
    for n in calls:
        o = v.process(softs[k * 8192:(k + n) * 8192], fill=1)
        k += n
        bits.append(o)
        nbits.append(len(o))
        st = v.stats()
        stats.append([np.float32(st[0]).view(np.int32)] + list(st[1:]))
    return dict(rate=rate, softs=softs, calls=np.array(calls), nbits=np.array(nbits), bits=np.packbits(np.concatenate(bits)),
                stats=np.array(stats, np.int64))


def dvbs_outer_fixture(seed):
    """the reference's DVBS_TS_Deframer and its deinterleaver / DVBSReedSolomon (libcorrect) / descrambler on 24 frames with
    byte errors (some packets beyond the code) behind 1234 bits of junk; frame stride 1632 and the module's own 204"""
    import dvbs_stream
    from test_dvbs_oracle import RefDeframer, RefOuter
    rng = np.random.default_rng(seed)
    ts, ch = dvbs_stream.outer_stream(24, rng)
    bad = dvbs_stream.add_errors(ch, rng, per_packet=(0, 11))
    bits = np.concatenate([rng.integers(0, 2, 1234, dtype=np.uint8), np.unpackbits(bad), rng.integers(0, 2, 6, dtype=np.uint8)])
    d = RefDeframer()
    f1, s1 = d.work(bits[:100001])
    f2, s2 = d.work(bits[100001:])
    frames = np.concatenate([f1, f2])
    o1, e1 = RefOuter().process(frames.reshape(-1), len(frames), 1632)
    n204 = (frames.size - 1632) // 204 + 1
    o2, e2 = RefOuter().process(frames.reshape(-1), n204, 204)
    return dict(bits=np.packbits(bits), nbits=len(bits), cut=100001, nframes=np.array([len(f1), len(f2)]), deframer_stats=np.array([s1, s2]),
                frames=frames, ts_1632=o1, err_1632=e1, ts_204=o2, err_204=e2)


def main():
    if not orclib.have_ref():
        raise SystemExit("oracle/_ref/libdvbs2_ref.so is missing: run `make -C oracle ref` in the build container")
    os.makedirs(OUT, exist_ok=True)
    for name, (short, rate, n) in {"chain_s1_2": (1, 3, 6), "chain_s8_9": (1, 10, 4), "chain_s1_4": (1, 0, 4),
                                   "chain_n1_2": (0, 3, 2)}.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **ldpc_chain_fixture(short, rate, n, 4242 + rate))
    np.savez_compressed(os.path.join(OUT, "bch_n12.npz"), **bch_fixture(0, 3, 1))
    np.savez_compressed(os.path.join(OUT, "bch_n10.npz"), **bch_fixture(0, 5, 2))
    np.savez_compressed(os.path.join(OUT, "bch_n8.npz"), **bch_fixture(0, 10, 3))
    np.savez_compressed(os.path.join(OUT, "bch_s12.npz"), **bch_fixture(1, 3, 4))
    np.savez_compressed(os.path.join(OUT, "demap_qpsk.npz"), **demap_fixture(1, 0, 1, 3, 0.0, 0.0, 5))
    np.savez_compressed(os.path.join(OUT, "demap_8psk35.npz"), **demap_fixture(3, 1, 1, 4, 0.0, 0.0, 6))
    np.savez_compressed(os.path.join(OUT, "demap_16apsk.npz"), **demap_fixture(4, 2, 1, 5, 3.15, 0.0, 7))
    np.savez_compressed(os.path.join(OUT, "demap_32apsk.npz"), **demap_fixture(5, 3, 1, 6, 2.84, 5.27, 8))
    np.savez_compressed(os.path.join(OUT, "tsparse_ts_n12.npz"), **ts_parser_fixture("ts", 32208, 11))
    np.savez_compressed(os.path.join(OUT, "tsparse_odd_s14.npz"), **ts_parser_fixture("ts_odd", 3072, 5))
    np.savez_compressed(os.path.join(OUT, "tsparse_gse_n12.npz"), **ts_parser_fixture("gse", 32208, 8))
    np.savez_compressed(os.path.join(OUT, "tsparse_gsemix_s12.npz"), **ts_parser_fixture("gse_mixed", 7032, 21))
    np.savez_compressed(os.path.join(OUT, "plfront_s36p.npz"), **plfront_fixture(35))
    np.savez_compressed(os.path.join(OUT, "dvbs_vit_r12_p90.npz"), **dvbs_vit_fixture(0, 1, 1, 31))
    np.savez_compressed(os.path.join(OUT, "dvbs_vit_r23.npz"), **dvbs_vit_fixture(1, 0, 4, 32))
    np.savez_compressed(os.path.join(OUT, "dvbs_vit_r56.npz"), **dvbs_vit_fixture(3, 0, 7, 33))
    np.savez_compressed(os.path.join(OUT, "dvbs_outer.npz"), **dvbs_outer_fixture(34))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
