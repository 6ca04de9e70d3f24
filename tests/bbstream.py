"""Transmit-side generators for BBFRAME streams (EN 302 307 5.1: mode adaptation), test infrastructure.

TS: 188-byte packets, sync byte replaced by the CRC-8 of the previous packet, sliced into data fields with
SYNCD pointing at the first sync/CRC byte.  GSE (TS 102 606): complete PDUs and Start/Continuation/End
fragments with CRC-32, padded with zero bytes.  Headers are sealed with the oracle's CRC-8."""
import numpy as np

import orclib


def bbheader(matype1, upl_bits, dfl_bits, sync, syncd_bits, matype2=0):
    h = np.array([matype1, matype2, upl_bits >> 8, upl_bits & 255, dfl_bits >> 8, dfl_bits & 255, sync,
                  syncd_bits >> 8, syncd_bits & 255, 0], np.uint8)
    orclib.oracle().orc_bbheader_seal(h)
    return h


def ts_packets(n, rng):
    pk = rng.integers(0, 256, (n, 188), dtype=np.uint8)
    pk[:, 0] = 0x47
    pk[:, 1] &= 0x7F  # transport_error_indicator clear
    return pk


def ts_stream_bytes(packets):
    """continuous user-packet stream: [CRC-8 of previous packet][187 bytes] per packet"""
    o = orclib.oracle()
    out = packets.copy()
    prev = 0
    for k in range(len(packets)):
        out[k, 0] = prev
        prev = o.orc_up_crc8(np.ascontiguousarray(packets[k, 1:]))
    return out.reshape(-1)


def ts_bbframes(kbch, packets, dfl_bytes=None, first_byte=0, pad=0):
    """slice the user-packet stream, starting at stream byte `first_byte`, into full BBFRAMEs of kbch/8 bytes.
    dfl_bytes: int or per-frame list (default: as large as fits); returns (frames, n_stream_bytes_used)"""
    stream = ts_stream_bytes(packets)
    kb = kbch // 8
    max_df = kb - 10
    frames, pos, f = [], first_byte, 0
    while True:
        df = max_df if dfl_bytes is None else (dfl_bytes[f] if hasattr(dfl_bytes, "__len__") else dfl_bytes)
        if pos + df > len(stream):
            break
        first_sync = (-pos) % 188  # bytes until the next sync/CRC slot
        fr = np.full(kb, pad, np.uint8)
        fr[:10] = bbheader(0xF0, 188 * 8, df * 8, 0x47, first_sync * 8)
        fr[10:10 + df] = stream[pos:pos + df]
        frames.append(fr)
        pos += df
        f += 1
        if hasattr(dfl_bytes, "__len__") and f >= len(dfl_bytes):
            break
    return np.stack(frames), pos


_CRC32_TAB = None


def crc32_mpeg(data, crc=0xFFFFFFFF):
    global _CRC32_TAB
    if _CRC32_TAB is None:
        tab = []
        for i in range(256):
            c = i << 24
            for _ in range(8):
                c = ((c << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if c & 0x80000000 else (c << 1) & 0xFFFFFFFF
            tab.append(c)
        _CRC32_TAB = tab
    for b in bytes(data):
        crc = ((crc << 8) & 0xFFFFFFFF) ^ _CRC32_TAB[((crc >> 24) ^ b) & 0xFF]
    return crc


def gse_complete(pdu, proto, label=None):
    """S=1 E=1; label: 6 bytes (LT=00) or None (LT=11, label re-use/none)"""
    body = bytes([proto >> 8, proto & 255]) + (bytes(label) if label is not None else b"") + bytes(pdu)
    lt = 0 if label is not None else 3
    return bytes([0xC0 | (lt << 4) | (len(body) >> 8), len(body) & 255]) + body


def gse_fragments(pdu, proto, frag_id, cuts, label=None):
    """Start / Continuation... / End packets for one PDU cut at byte offsets `cuts`; CRC-32 over
    total_length, protocol type, label and PDU"""
    pdu = bytes(pdu)
    lab = bytes(label) if label is not None else b""
    lt = 0 if label is not None else 3
    total = len(pdu) + 2 + len(lab)
    tl, pt = bytes([total >> 8, total & 255]), bytes([proto >> 8, proto & 255])
    crc = crc32_mpeg(tl + pt + lab + pdu)
    parts = [pdu[a:b] for a, b in zip([0] + list(cuts), list(cuts) + [len(pdu)])]
    out = []
    for k, part in enumerate(parts):
        if k == 0:
            body = bytes([frag_id]) + tl + pt + lab + part
            flags = 0x80
        elif k == len(parts) - 1:
            body = bytes([frag_id]) + part + crc.to_bytes(4, "big")
            flags = 0x40
        else:
            body = bytes([frag_id]) + part
            flags = 0x00
        # continuation/end packets carry LT=11 so that they are not mistaken for padding
        ltk = lt if k == 0 else 3
        out.append(bytes([flags | (ltk << 4) | (len(body) >> 8), len(body) & 255]) + body)
    return out


def gse_bbframes(kbch, fields):
    """fields: list of lists of GSE packets (bytes) per frame; zero padding to DFL = everything that fits"""
    kb = kbch // 8
    frames = []
    for pk in fields:
        df = b"".join(pk)
        assert len(df) <= kb - 10
        fr = np.zeros(kb, np.uint8)
        fr[:10] = bbheader(0x70, 0, (kb - 10) * 8, 0, 0)
        fr[10:10 + len(df)] = np.frombuffer(df, np.uint8)
        frames.append(fr)
    return np.stack(frames)


def odd_ts_scenario(rng, kbch):
    """120 TS frames with short data fields (partial unit never completed), header faults of every kind and
    non-TS frames in between"""
    pk = ts_packets(400, rng)
    dfl = [int(x) for x in rng.choice([374, 373, 200, 187, 188, 189, 100, 16, 8, 360], 120)]
    frames, _ = ts_bbframes(kbch, pk, dfl_bytes=dfl, pad=0xA5)
    f = frames.copy()
    f[7, 9] ^= 1                                                # CRC-8 failure
    f[15, :10] = bbheader(0xF0, 1504, kbch - 80 + 8, 0x47, 0)   # DFL too long
    f[22, :10] = bbheader(0xF0, 1504, 800, 0x47, 800 - 8)       # SYNCD >= DFL-8
    f[23, :10] = bbheader(0xF0, 1504, 4, 0x47, 0)               # DFL-8 negative
    f[30, :10] = bbheader(0xF0, 1504, 803, 0x47, 16)            # DFL not a byte multiple
    f[41, :10] = bbheader(0x30, 0, 2000, 0, 0)                  # ts_gs = 00 (generic packetized): ignored
    f[42, :10] = bbheader(0xB0, 0, 2000, 0, 0)                  # ts_gs = 10 (reserved): ignored
    f[50, :10] = bbheader(0x74, 0, 2000, 0, 0)                  # GSE with NPD set: field skipped
    return f


def gse_scenario(rng):
    """five data fields: padding-only resync frame, complete PDUs with/without label, three interleaved
    fragmented PDUs (one with a protocol type the parser does not echo) and one whose CRC-32 fails"""
    lab = bytes(range(1, 7))
    pdus = [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in (40, 1200, 900, 64, 3000, 500, 10, 2000)]
    f0 = []                                                        # resync frame: all padding
    a = gse_fragments(pdus[1], 0x0800, 7, [500], label=lab)
    b = gse_fragments(pdus[4], 0x86DD, 9, [700, 1900])
    c = gse_fragments(pdus[7], 0x88B5, 3, [100, 1100], label=lab)   # protocol without EtherType echo
    broken = gse_fragments(pdus[2], 0x0800, 5, [450])
    broken[1] = broken[1][:-1] + bytes([broken[1][-1] ^ 1])        # CRC-32 mismatch on reassembly
    f1 = [gse_complete(pdus[0], 0x0800, label=lab), a[0], gse_complete(pdus[3], 0x86DD)]
    f2 = [b[0], a[1], c[0], gse_complete(pdus[6], 0x1234)]
    f3 = [b[1], c[1], broken[0]]
    f4 = [c[2], b[2], broken[1], gse_complete(pdus[5], 0x0800, label=lab)]
    return [f0, f1, f2, f3, f4]


def random_ts_scenario(rng, kbch, nframes=80):
    """random mix: data-field sizes from tiny to full, occasional header faults, non-TS frames, a restart of the
    packet stream at an arbitrary byte (as after a MODCOD change)"""
    kb = kbch // 8
    max_df = kb - 10
    choices = [max_df, max_df, max_df, max_df - 1, max(8, max_df // 2), 188, 187, 189, 376, 100, 16]
    choices = [c for c in choices if 8 <= c <= max_df]
    dfl = [int(x) for x in rng.choice(choices, nframes)]
    pk = ts_packets(sum(dfl) // 188 + 4, rng)
    frames, _ = ts_bbframes(kbch, pk, dfl_bytes=dfl, first_byte=int(rng.integers(0, 188)), pad=int(rng.integers(0, 256)))
    f = frames.copy()
    for i in rng.choice(len(f), max(1, len(f) // 12), replace=False):
        kind = int(rng.integers(0, 6))
        if kind == 0:
            f[i, int(rng.integers(0, 10))] ^= 1 << int(rng.integers(0, 8))          # CRC-8 failure
        elif kind == 1:
            f[i, :10] = bbheader(0xF0, 1504, kbch - 80 + 8 * int(rng.integers(1, 5)), 0x47, 0)   # DFL too long
        elif kind == 2:
            d = 8 * int(rng.integers(2, 40))
            f[i, :10] = bbheader(0xF0, 1504, d, 0x47, d - 8 + 8 * int(rng.integers(0, 3)))       # SYNCD >= DFL-8
        elif kind == 3:
            f[i, :10] = bbheader(0xF0, 1504, 8 * int(rng.integers(2, 40)) + int(rng.integers(1, 8)), 0x47, 0)   # DFL % 8
        elif kind == 4:
            f[i, :10] = bbheader(int(rng.choice([0x30, 0xB0, 0x74])), 0, 8 * int(rng.integers(2, 40)), 0, 0)    # not TS
        else:
            f[i, :10] = bbheader(0xF0, 1504, 8 * int(rng.integers(1, 3)), 0x47, 0)     # DFL of 8 or 16 bits
    return f


def random_gse_scenario(rng, kbch, nframes=40, ts_every=0):
    """random GSE stream: complete PDUs (with and without label, known and unknown protocol types) and fragmented
    PDUs whose pieces are spread over the following frames, up to five FragIDs in flight at once (the parser has
    three slots: the others are dropped), FragIDs restarted before their last fragment, CRC-32 failures, tiny
    fragments.  ts_every > 0: every ts_every-th frame is a TS frame, and now and then a frame with a broken header
    followed by the all-padding frame a resynchronising parser needs (it enters the next field one byte late).
    Returns frames [n][kbch/8]."""
    kb = kbch // 8
    room_full = kb - 10
    protos = [0x0800, 0x86DD, 0x88B5, 0x0806]
    lab = bytes(range(0x10, 0x16))
    pending = []            # queues of fragments still to be sent, one list per PDU in flight
    fields = [[]]           # the parser starts out of sync: an all-padding frame first
    for f in range(nframes):
        room = room_full - int(rng.integers(0, 40))
        pk = []
        # fragments of PDUs in flight first, in random order, at most one piece of each per frame
        order = list(rng.permutation(len(pending)))
        for k in order:
            q = pending[k]
            if q and len(q[0]) <= room and rng.random() < 0.8:
                pk.append(q.pop(0))
                room -= len(pk[-1])
        pending = [q for q in pending if q]
        while room > 40 and rng.random() < 0.85:
            kind = rng.random()
            label = lab if rng.random() < 0.5 else None
            proto = int(rng.choice(protos))
            if kind < 0.55 or len(pending) >= 5:
                n = int(rng.integers(1, min(1500, room - 14)))
                pk.append(gse_complete(rng.integers(0, 256, n, dtype=np.uint8).tobytes(), proto, label=label))
            else:
                n = int(rng.integers(30, 6000))
                pdu = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
                nparts = int(rng.integers(2, 6))
                cuts = sorted(set(int(x) for x in rng.integers(1, n, nparts - 1)))
                # every piece must fit a data field together with its header
                while any(b - a > room_full - 60 for a, b in zip([0] + cuts, cuts + [n])):
                    cuts = sorted(set(cuts + [int(x) for x in rng.integers(1, n, 2)]))
                frag_id = int(rng.choice([3, 7, 9, 200, 255, 0]))
                parts = gse_fragments(pdu, proto, frag_id, cuts, label=label)
                r = rng.random()
                if r < 0.15:      # CRC-32 failure
                    parts[-1] = parts[-1][:-1] + bytes([parts[-1][-1] ^ 0x40])
                elif r < 0.25:    # the last fragment never comes
                    parts = parts[:-1]
                if len(parts[0]) > room:
                    break
                pk.append(parts.pop(0))
                pending.append(parts)
            room -= len(pk[-1])
        fields.append(pk)
    frames = gse_bbframes(kbch, fields)
    if ts_every:
        tsf, _ = ts_bbframes(kbch, ts_packets(nframes * (kb // 188 + 2), rng), first_byte=int(rng.integers(0, 188)))
        out, t = [], 0
        for f in range(len(frames)):
            if f % ts_every == ts_every - 1:
                out.append(tsf[t]); t += 1
                if rng.random() < 0.3:                          # sync loss, then the padding frame
                    bad = tsf[t].copy(); t += 1
                    bad[9] ^= 0x55
                    out.append(bad)
                    out.append(gse_bbframes(kbch, [[]])[0])
            out.append(frames[f])
        frames = np.stack(out)
    return frames
