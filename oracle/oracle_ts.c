/* TEST INFRASTRUCTURE -- CPU oracle, downstream row 8(f)-1: BBFRAME -> MPEG-TS packets / GSE PDUs.
 *
 * Restates BBFrameTSParser (dvbs2/bbframe_ts_parser.cpp, bbframe_ts_parser.h) in plain C, including the
 * behaviour that only shows on odd input (kept because parity is defined by what the reference does):
 *   - a frame is dropped, and the parser falls out of sync, on BBHEADER CRC-8 failure, DFL > kbch-80,
 *     SYNCD >= DFL-8 (signed comparison) or DFL not a multiple of 8            (:122-150)
 *   - after a resync the data field is entered SYNCD/8+1 bytes in, i.e. just past the first sync/CRC byte,
 *     also for GSE frames                                                       (:157-168)
 *   - TS mode cuts the byte stream into 188-byte units that END with the next packet's sync/CRC byte; the
 *     unit is emitted as 0x47 followed by its first 187 bytes; the CRC byte is never checked   (:173-201)
 *   - a partial unit carried over from the previous frame is only completed when the new data field still
 *     holds >= 188 bytes, otherwise it is overwritten by the new remainder      (:176-207)
 *   - output stops (and the remaining frames of the call are skipped) once fewer than 189 bytes of room are
 *     left                                                                      (:176,208-211)
 *   - GSE: the label-type test compares (h1 & 0x30) >> 2 with 0 and 2, so only "6-byte label" (0) and
 *     "no label" (anything else) exist; PDUs go out behind a 2-byte zero GRE header plus the protocol type
 *     when it is IPv4/IPv6; fragments are reassembled in up to three buffers keyed by FragID and checked
 *     with CRC-32 (poly 0x04C11DB7, init all-ones, no final XOR)                (:215-384)
 * Domain: GSE input must be well formed (lengths inside the data field); the reference does not bound-check
 * and neither result is defined otherwise.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

struct orc_ts_parser {
    unsigned kbch, max_dfl;
    unsigned count;          /* bytes of an unfinished 188-byte unit held in unit[] */
    int synched;
    uint8_t unit[188];
    /* GSE reassembly, three concurrent FragIDs (bbframe_ts_parser.h:91-98) */
    uint8_t frag_buf[3][65536];
    int frag_on[3], frag_len[3], frag_id[3];
    uint16_t frag_proto[3];
    uint32_t frag_crc[3];
    uint32_t crc32_tab[256];
    /* observable members */
    uint8_t last_header[10];
    int have_header;
    int last_gse_crc_err, last_bb_cnt, last_bb_proc;
};

/* crc32_init / crc32_checksum (:82-98): MSB-first CRC-32, one table entry = the register after clocking in a byte */
static void crc32_build(uint32_t* tab)
{
    for (unsigned i = 0; i < 256; ++i) {
        uint32_t c = (uint32_t)i << 24;
        for (int b = 0; b < 8; ++b) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        tab[i] = c;
    }
}
static uint32_t crc32_run(const uint32_t* tab, const uint8_t* p, unsigned n, uint32_t crc)
{
    while (n--) crc = (crc << 8) ^ tab[((crc >> 24) ^ *p++) & 0xFFu];
    return crc;
}

orc_ts_parser* orc_ts_create(int kbch_bits) /* setFrameSize (:31-43) */
{
    orc_ts_parser* p = (orc_ts_parser*)calloc(1, sizeof *p);
    if (!p) return NULL;
    p->kbch = (unsigned)kbch_bits;
    p->max_dfl = p->kbch - 80;
    crc32_build(p->crc32_tab);
    return p;
}
void orc_ts_destroy(orc_ts_parser* p) { free(p); }

/* 2-byte GRE header, then the EtherType when it is one of the two the reference knows (:265-274, :349-358) */
static int put_gre(uint8_t* out, int o, uint16_t proto)
{
    out[o++] = 0;
    out[o++] = 0;
    if (proto == 0x0800 || proto == 0x86DD) {
        out[o++] = (uint8_t)(proto >> 8);
        out[o++] = (uint8_t)proto;
    }
    return o;
}

static int gse_field(orc_ts_parser* p, const uint8_t* df, unsigned df_bytes, int plain, uint8_t* out, int o)
{
    unsigned at = 0;
    while (at < df_bytes) {
        if (!plain) break; /* ISSY, NPD or UPL != 0: skip the field (:386-388) */
        const uint8_t* g = df + at;
        const int start = g[0] >> 7, end = (g[0] >> 6) & 1;
        const int label6 = ((g[0] & 0x30) >> 2) == 0;
        if (!start && !end && label6) break; /* padding (:222-224) */
        uint16_t len = (uint16_t)(((g[0] & 0x0F) << 8) | g[1]);
        if (start && end) { /* complete PDU (:231-279) */
            uint16_t proto = (uint16_t)((g[2] << 8) | g[3]);
            unsigned skip = 4;
            len = (uint16_t)(len - 2);
            if (label6) {
                skip += 6;
                len = (uint16_t)(len - 6);
            }
            o = put_gre(out, o, proto);
            memcpy(out + o, g + skip, len);
            o += len;
            at += skip + len;
        } else if (start) { /* first fragment (:281-334) */
            const int id = g[2];
            uint16_t proto = (uint16_t)((g[5] << 8) | g[6]);
            unsigned skip = 7;
            len = (uint16_t)(len - 5);
            if (label6) {
                skip += 6;
                len = (uint16_t)(len - 6);
            }
            for (int r = 0; r < 3; ++r) {
                if (p->frag_on[r] && p->frag_id[r] != id) continue;
                p->frag_on[r] = 1;
                p->frag_id[r] = id;
                p->frag_proto[r] = proto;
                memcpy(p->frag_buf[r], g + skip, len);
                p->frag_len[r] = len;
                uint32_t c = 0xFFFFFFFFu; /* over total length, protocol type, label, data */
                c = crc32_run(p->crc32_tab, g + 3, 2, c);
                c = crc32_run(p->crc32_tab, g + 5, 2, c);
                if (label6) c = crc32_run(p->crc32_tab, g + 7, 6, c);
                p->frag_crc[r] = crc32_run(p->crc32_tab, g + skip, len, c);
                break;
            }
            at += skip + len;
        } else { /* continuation (:363-377) or last fragment (:335-362) */
            const int id = g[2];
            len = (uint16_t)(len - 1);
            for (int r = 0; r < 3; ++r) {
                if (!p->frag_on[r] || p->frag_id[r] != id) continue;
                memcpy(p->frag_buf[r] + p->frag_len[r], g + 3, len);
                if (!end) {
                    p->frag_len[r] += len;
                    p->frag_crc[r] = crc32_run(p->crc32_tab, g + 3, len, p->frag_crc[r]);
                    break;
                }
                p->frag_on[r] = 0;
                p->frag_len[r] += len - 4;
                uint32_t c = crc32_run(p->crc32_tab, g + 3, (unsigned)len - 4, p->frag_crc[r]);
                p->frag_crc[r] = c;
                const uint8_t* t = g + 3 + len - 4;
                uint32_t rx = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
                if (c != rx) {
                    p->last_gse_crc_err = 1;
                } else {
                    p->last_gse_crc_err = 0;
                    o = put_gre(out, o, p->frag_proto[r]);
                    memcpy(out + o, p->frag_buf[r], (size_t)p->frag_len[r]);
                    o += p->frag_len[r];
                }
                break;
            }
            at += 3 + len;
        }
    }
    return o;
}

int orc_ts_work(orc_ts_parser* p, const uint8_t* bbframes, int cnt, uint8_t* out, int out_cap) /* work (:100-392) */
{
    int o = 0, processed = 0;
    const unsigned stride = p->kbch / 8;
    for (int f = 0; f < cnt; ++f) {
        const uint8_t* bb = bbframes + (size_t)stride * f;
        if (orc_bbheader_crc8(bb) != 0) {
            p->synched = 0;
            continue;
        }
        const int dfl = (bb[4] << 8) | bb[5], syncd = (bb[7] << 8) | bb[8];
        if (dfl > (int)p->max_dfl || syncd >= dfl - 8 || dfl % 8 != 0) {
            p->synched = 0;
            continue;
        }
        unsigned left = (unsigned)dfl / 8;
        const uint8_t* df = bb + 10;
        if (!p->synched) {
            df += syncd / 8 + 1;
            left -= (unsigned)syncd / 8 + 1;
            p->count = 0;
            p->synched = 1;
        }
        memcpy(p->last_header, bb, 10);
        p->have_header = 1;
        ++processed;
        const int ts_gs = bb[0] >> 6;
        if (ts_gs == 3) {
            while (left >= 188 && out_cap - o > 188) {
                const uint8_t* unit;
                if (p->count > 0) {
                    const unsigned need = 188 - p->count;
                    memcpy(p->unit + p->count, df, need);
                    df += need;
                    left -= need;
                    unit = p->unit;
                    p->count = 0;
                } else {
                    unit = df;
                    df += 188;
                    left -= 188;
                }
                out[o] = 0x47;
                memcpy(out + o + 1, unit, 187);
                o += 188;
            }
            if (left >= 188) {
                /* only after running out of room.  The reference overruns packet_reassembly here and carries
                 * count >= 188 into the next call (a negative memcpy length): undefined.  Domain completion shared
                 * with the device parser: drop out of sync, pick up again at the next frame's SYNCD. */
                p->synched = 0;
                p->count = 0;
            } else if (left > 0) {
                p->count = left;
                memcpy(p->unit, df, left);
            }
            if (out_cap - o <= 188) break;
        } else if (ts_gs == 1) {
            const int plain = !((bb[0] >> 3) & 1) && !((bb[0] >> 2) & 1) && ((bb[2] << 8) | bb[3]) == 0;
            o = gse_field(p, df, (unsigned)dfl / 8, plain, out, o);
        }
    }
    p->last_bb_cnt = cnt;
    p->last_bb_proc = processed;
    return o;
}

void orc_ts_stats(const orc_ts_parser* p, uint8_t last_header[10], int* have_header, int* last_bb_cnt, int* last_bb_proc,
                  int* last_gse_crc_err, int* synched, int* pending)
{
    if (last_header) memcpy(last_header, p->last_header, 10);
    if (have_header) *have_header = p->have_header;
    if (last_bb_cnt) *last_bb_cnt = p->last_bb_cnt;
    if (last_bb_proc) *last_bb_proc = p->last_bb_proc;
    if (last_gse_crc_err) *last_gse_crc_err = p->last_gse_crc_err;
    if (synched) *synched = p->synched;
    if (pending) *pending = (int)p->count;
}

/* Transmit side for tests: fill in the CRC-8 so that check_crc8 over all 80 bits gives 0 (EN 302 307 5.1.6). */
void orc_bbheader_seal(uint8_t* hdr10)
{
    for (int v = 0; v < 256; ++v) {
        hdr10[9] = (uint8_t)v;
        if (orc_bbheader_crc8(hdr10) == 0) return;
    }
}
/* CRC-8 of a user packet's 187 bytes after the sync byte, as the mode adapter inserts it in place of the next
 * sync byte (EN 302 307 5.1.4): the value c for which check_crc8(payload || c) == 0. */
uint8_t orc_up_crc8(const uint8_t* payload187)
{
    unsigned crc = 0;
    for (int n = 0; n < 187 * 8; ++n) {
        unsigned b = ((payload187[n / 8] >> (7 - (n % 8))) & 1u) ^ (crc & 1u);
        crc >>= 1;
        if (b) crc ^= 0xABu;
    }
    /* register is bit-reflected relative to transmission order: the byte that drives it to 0 is its mirror */
    unsigned m = 0;
    for (int i = 0; i < 8; ++i) m |= ((crc >> i) & 1u) << (7 - i);
    return (uint8_t)m;
}
