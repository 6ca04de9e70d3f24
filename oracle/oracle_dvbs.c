/* TEST INFRASTRUCTURE -- see oracle.h.  Restatement of the outer decoder of the reference's DVB-S chain (SURVEY.md
 * 8(f) rank 4, the byte-domain half): the body of the frame loop of DVBSDemod::process, dvbs/module_dvbs_demod.cpp:91-106.
 *
 *   convolutional deinterleaver   dvbs/dvbs_interleaving.h:30-40,57-70      (I = 12, M = 17; FIFO j has 17 (11 - j) cells)
 *   RS(204,188) wrapper           dvbs/dvbs_reedsolomon.h:18-48             (51 zeros in front; NOTE :33-36 below)
 *   Reed-Solomon decoder          common/correct/reed-solomon/decode.c:12-27 (syndromes), :31-124 (Berlekamp-Massey),
 *                                 :128-145 (root search), :147-198 (Forney), :200-228 (locations), :303-376 (decode);
 *                                 field: reed-solomon/field.h (GF(256), x^8+x^4+x^3+x^2+1, log[1] = 255)
 *   energy-dispersal descrambler  dvbs/dvbs_scrambling.h:16-45
 *
 * Observable quirks that are part of the behaviour and therefore restated:
 *   * DVBSReedSolomon::decode tests libcorrect's return value against 1, not -1 (:33-36): a packet the decoder gives up on
 *     is NOT reported; libcorrect has not touched the output buffer then, so the packet is replaced by whatever the
 *     previous call left there -- the last packet that decoded -- and the "error count" is the number of bytes in which
 *     the received packet differs from that.
 *   * an error located at coefficient 0 (the last parity byte) has log(1) = 255 as its position (field.h: log[1] is
 *     overwritten at i = 255): the correction goes to coefficient 255, which is not part of the word; the message
 *     bytes are unaffected either way.
 *   * the descrambler's register starts at 0 and only an inverted sync byte (0xB8) loads it: until then nothing is
 *     descrambled.
 * Pinned bit for bit against the reference's own headers + vendored libcorrect (tests/test_dvbs_oracle.py). */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- GF(256) as libcorrect builds it (field.h:24-45) ---- */
static uint8_t g_exp[512], g_log[256];
static int g_init;
static void gf_init(void)
{
    if (g_init) return;
    unsigned e = 1;
    g_exp[0] = 1;
    g_log[0] = 0;
    for (unsigned i = 1; i < 512; ++i) {
        e *= 2;
        if (e > 255) e ^= 0x11d;
        g_exp[i] = (uint8_t)e;
        if (i < 256) g_log[e] = (uint8_t)i;
    }
    g_init = 1;
}
static uint8_t gmul(uint8_t l, uint8_t r) { return (!l || !r) ? 0 : g_exp[g_log[l] + g_log[r]]; }
static uint8_t gdiv(uint8_t l, uint8_t r) { return (!l || !r) ? 0 : g_exp[255 + g_log[l] - g_log[r]]; }
/* polynomial_eval_lut: value of sum c[i] x^i at the element v (v = 0: c[0]); the powers of v are counted as
 * polynomial_build_exp_lut counts them: logs i * log(v) kept in 1..255 (log[1] = 255, never 0) */
static uint8_t peval(const uint8_t* c, unsigned order, uint8_t v)
{
    if (v == 0) return c[0];
    uint8_t res = 0;
    unsigned acc = 255;                      /* log[1] */
    const unsigned vlog = g_log[v];
    for (unsigned i = 0; i <= order; ++i) {
        if (c[i]) res ^= g_exp[g_log[c[i]] + acc];
        acc += vlog;
        if (acc > 255) acc -= 255;
    }
    return res;
}

/* correct_reed_solomon_decode(rs, encoded, 255, msg) for RS(255,239), first root alpha^0, gap 1.
 * Returns 239 and writes msg[0..239), or -1 with msg untouched. */
static int rs_decode(const uint8_t* encoded, uint8_t* msg)
{
    enum { N = 255, D = 16 };
    uint8_t r[N + 1];                                   /* received polynomial, coefficient i = encoded[254 - i]; (+1: the stray write) */
    for (int i = 0; i < N; ++i) r[i] = encoded[N - (i + 1)];
    r[N] = 0;
    uint8_t syn[D];
    int all_zero = 1;
    for (int i = 0; i < D; ++i) {
        syn[i] = peval(r, N - 1, g_exp[i % 255]);        /* generator root i: alpha^i */
        if (syn[i]) all_zero = 0;
    }
    if (all_zero) {
        for (int i = 0; i < N - D; ++i) msg[i] = r[N - (i + 1)];
        return N - D;
    }
    /* Berlekamp-Massey (decode.c:31-124) */
    uint8_t loc[2 * D + 2], last[2 * D + 2];
    memset(loc, 0, sizeof loc);
    memset(last, 0, sizeof last);
    loc[0] = last[0] = 1;
    unsigned loc_order = 0, last_order = 0, numerrors = 0, delay = 1;
    uint8_t last_disc = 1;
    for (unsigned i = 0; i < D; ++i) {
        uint8_t disc = syn[i];
        for (unsigned j = 1; j <= numerrors; ++j) disc ^= gmul(loc[j], syn[i - j]);
        if (!disc) {
            delay++;
            continue;
        }
        if (2 * numerrors <= i) {
            for (int j = (int)last_order; j >= 0; --j) last[j + delay] = gdiv(gmul(last[j], disc), last_disc);
            for (int j = (int)delay - 1; j >= 0; --j) last[j] = 0;
            for (unsigned j = 0; j <= last_order + delay; ++j) {
                uint8_t t = loc[j];
                loc[j] ^= last[j];
                last[j] = t;
            }
            unsigned t = loc_order;
            loc_order = last_order + delay;
            last_order = t;
            numerrors = i + 1 - numerrors;
            last_disc = disc;
            delay = 1;
            continue;
        }
        for (int j = (int)last_order; j >= 0; --j) loc[j + delay] ^= gdiv(gmul(last[j], disc), last_disc);
        if (last_order + delay > loc_order) loc_order = last_order + delay;
        delay++;
    }
    /* roots of the locator among all 256 elements, ascending (decode.c:128-145) */
    uint8_t roots[2 * D];
    unsigned nroots = 0;
    for (unsigned v = 0; v < 256; ++v)
        if (!peval(loc, loc_order, (uint8_t)v)) {
            if (nroots < 2 * D) roots[nroots] = (uint8_t)v;
            nroots++;
        }
    if (nroots != loc_order) return -1;
    /* evaluator = locator * S(x) mod x^16 (:147-160, order 15), derivative (:74-87) */
    uint8_t omega[D], der[2 * D + 2];
    memset(omega, 0, sizeof omega);
    for (unsigned i = 0; i <= loc_order && i < D; ++i)
        for (unsigned j = 0; j + i < D; ++j) omega[i + j] ^= gmul(loc[i], syn[j]);
    unsigned der_order = loc_order - 1;
    for (unsigned i = 0; i <= der_order; ++i) der[i] = ((i + 1) & 1) ? loc[i + 1] : 0;
    for (unsigned i = 0; i < loc_order; ++i) {
        const uint8_t root = roots[i];                   /* never 0: the locator's constant term is 1 */
        const uint8_t inv = g_exp[(255 - g_log[root]) % 255];   /* field_pow(root, c - 1), c = 0 */
        /* location (:200-228): log of 1 / root, with log[1] = 255 */
        const uint8_t x = gdiv(1, root);
        const unsigned where = g_log[x];
        /* value (:162-198): root^(c - 1) * omega(root) / lambda'(root), c = 0 */
        const uint8_t val = gmul(inv, gdiv(peval(omega, D - 1, root), peval(der, der_order, root)));
        r[where] ^= val;
    }
    for (int i = 0; i < N - D; ++i) msg[i] = r[N - (i + 1)];
    return N - D;
}

struct orc_dvbs_outer {
    uint8_t fifo[12][17 * 11 + 1];
    int head[12];
    uint8_t obuffer[255];       /* DVBSReedSolomon::obuffer: survives from call to call */
    int reg;                    /* DVBSScrambling::reg */
};

orc_dvbs_outer* orc_dvbs_outer_create(void)
{
    gf_init();
    return (orc_dvbs_outer*)calloc(1, sizeof(orc_dvbs_outer));
}
void orc_dvbs_outer_destroy(orc_dvbs_outer* p) { free(p); }

static int prbs8(orc_dvbs_outer* p)
{
    int res = 0;
    for (int i = 0; i < 8; ++i) {
        int fb = ((p->reg >> 13) ^ (p->reg >> 14)) & 1;
        p->reg = ((p->reg << 1) | fb) & 0x7fff;
        res = (res << 1) | fb;
    }
    return res;
}

void orc_dvbs_outer_frame(orc_dvbs_outer* p, const uint8_t* frame, uint8_t* out, int* errors)
{
    uint8_t d[1632];
    /* deinterleave (dvbs_interleaving.h:57-70): byte n goes through FIFO n % 12 of 17 (11 - n % 12) cells */
    for (int n = 0; n < 1632; ++n) {
        const int j = n % 12, len = 17 * (11 - j);
        if (len == 0) {
            d[n] = frame[n];
        } else {
            d[n] = p->fifo[j][p->head[j]];
            p->fifo[j][p->head[j]] = frame[n];
            p->head[j] = (p->head[j] + 1) % len;
        }
    }
    for (int i = 0; i < 8; ++i) {   /* DVBSReedSolomon::decode (dvbs_reedsolomon.h:27-47) */
        uint8_t buffer[255];
        memset(buffer, 0, 51);
        memcpy(buffer + 51, d + 204 * i, 188);
        memcpy(buffer + 239, d + 204 * i + 188, 16);
        (void)rs_decode(buffer, p->obuffer);             /* (a failure leaves obuffer as it was) */
        int err = 0;
        for (int k = 51; k < 239; ++k) err += (buffer[k] ^ p->obuffer[k]) != 0;
        memcpy(d + 204 * i, p->obuffer + 51, 188);
        errors[i] = err;
    }
    for (int pkt = 0; pkt < 8; ++pkt) {   /* DVBSScrambling::descramble (dvbs_scrambling.h:30-43) */
        uint8_t* f = d + pkt * 204;
        if (f[0] == 0xB8) p->reg = 0xa9;
        else prbs8(p);
        f[0] = 0x47;
        for (int k = 1; k < 188; ++k) f[k] ^= (uint8_t)prbs8(p);
        memcpy(out + 188 * pkt, f, 188);
    }
}

/* the frame loop: frame k starts at frames + k * stride (the module uses 204: module_dvbs_demod.cpp:91) */
void orc_dvbs_outer_process(orc_dvbs_outer* p, const uint8_t* frames, int nframes, int stride, uint8_t* out, int* errors)
{
    for (int k = 0; k < nframes; ++k) orc_dvbs_outer_frame(p, frames + (size_t)k * stride, out + (size_t)k * 1504, errors + 8 * k);
}

/* transmit side for the tests: systematic RS(204,188) parity (16 bytes) of a 188-byte packet, generator
 * prod (x - alpha^i), i = 0..15, as correct_reed_solomon_encode with 51 leading zeros produces it */
void orc_rs204_parity(const uint8_t* msg188, uint8_t* parity16)
{
    gf_init();
    uint8_t g[17];
    memset(g, 0, sizeof g);
    g[0] = 1;
    for (int i = 0; i < 16; ++i) {        /* g(x) *= (x + alpha^i); g[k] = coefficient of x^k */
        const uint8_t a = g_exp[i];
        for (int k = i + 1; k >= 1; --k) g[k] = g[k - 1] ^ gmul(g[k], a);
        g[0] = gmul(g[0], a);
    }
    uint8_t rem[16];
    memset(rem, 0, sizeof rem);
    for (int i = 0; i < 188; ++i) {       /* long division, highest coefficient first */
        const uint8_t fb = msg188[i] ^ rem[15];
        for (int k = 15; k >= 1; --k) rem[k] = rem[k - 1] ^ gmul(fb, g[k]);
        rem[0] = gmul(fb, g[0]);
    }
    for (int i = 0; i < 16; ++i) parity16[i] = rem[15 - i];
}

/* ---- DVBS_TS_Deframer::work (dvbs/dvbs_ts_deframer.cpp:44-101): a window of 8 x 204 x 8 bits slides over the unpacked
 * bit stream (one bit per byte, as the Viterbi decoder delivers it); wherever its eight sync bytes differ from
 * B8 47 47 47 47 47 47 47 in at most 8 bits the window goes out as a frame of 1632 bytes, wherever they differ from the
 * inverted pattern in at most 8 bits it goes out inverted.  The reference shifts an array by memmove for every bit;
 * here the last 13056 bits sit in a ring.  (The reference leaves its array uninitialised; the harness and this
 * restatement start from zeros.) ---- */
#define DEF_BITS 13056
struct orc_dvbs_deframer {
    uint8_t ring[DEF_BITS];
    int pos;                 /* where the next bit goes = the oldest bit */
    int errors_nor, errors_inv;
};
orc_dvbs_deframer* orc_dvbs_deframer_create(void) { return (orc_dvbs_deframer*)calloc(1, sizeof(orc_dvbs_deframer)); }
void orc_dvbs_deframer_destroy(orc_dvbs_deframer* p) { free(p); }
void orc_dvbs_deframer_stats(const orc_dvbs_deframer* p, int* errors_nor, int* errors_inv)
{
    *errors_nor = p->errors_nor;
    *errors_inv = p->errors_inv;
}
int orc_dvbs_deframer_work(orc_dvbs_deframer* p, const uint8_t* input, int size, uint8_t* output)
{
    int frame_count = 0;
    for (int ibit = 0; ibit < size; ++ibit) {
        p->ring[p->pos] = input[ibit];
        p->pos = (p->pos + 1) % DEF_BITS;          /* window bit i = ring[(pos + i) % DEF_BITS] */
        int t_nor = 0, t_inv = 0;
        for (int i = 0; i < 8; ++i) {
            unsigned b = 0;
            for (int k = 0; k < 8; ++k) b |= (unsigned)p->ring[(p->pos + 204 * 8 * i + k) % DEF_BITS] << (7 - k);   /* pack_8: plain ORs of shifted bytes */
            b &= 0xFF;
            t_nor += __builtin_popcount(b ^ (i == 0 ? 0xB8u : 0x47u));
            t_inv += __builtin_popcount(b ^ (i == 0 ? 0x47u : 0xB8u));
        }
        if (t_nor <= 8) {
            uint8_t* f = output + (size_t)frame_count * 1632;
            memset(f, 0, 1632);
            for (int i = 0; i < DEF_BITS; ++i) f[i / 8] = (uint8_t)(f[i / 8] << 1 | p->ring[(p->pos + i) % DEF_BITS]);
            frame_count++;
            p->errors_nor = t_nor;
            p->errors_inv = 0;
        }
        if (t_inv <= 8) {
            uint8_t* f = output + (size_t)frame_count * 1632;
            memset(f, 0, 1632);
            for (int i = 0; i < DEF_BITS; ++i) f[i / 8] = (uint8_t)(f[i / 8] << 1 | !p->ring[(p->pos + i) % DEF_BITS]);
            frame_count++;
            p->errors_nor = 0;
            p->errors_inv = t_inv;
        }
    }
    return frame_count;
}
