/* TEST INFRASTRUCTURE -- CPU oracle, upstream row 8(f)-3: PL frame synchronisation, PLHEADER demodulation / PLS
 * decoding and the coarse frequency error detector.
 *
 * Restates, in plain C and in the reference's own order of float operations,
 *   S2PLSyncBlock::process / internal_process / correlate_*_diff   dvbs2/dvbs2_pl_sync.cpp:80-193
 *   S2PLHDRDemod::init / process / checkSyncMarker                 dvbs2/dvbs2_plhdr_demod.cpp:5-12,33-79
 *   dvbs2_pilot_coarse_fed                                         dvbs2/dvbs2_fed.h:7-48
 *   s2_sof, s2_plscodes                                            dvbs2/s2_defs.h:16-87
 * Pinned by tests/test_plsync_oracle.py against the unmodified reference sources compiled into oracle/_ref with
 * stand-ins for SDR++ core (complex_t, Processor, PhaseControlLoop, phasor, step) and for the two VOLK kernels
 * (oracle/shim/): the control flow and the arithmetic order are the reference's own, the primitives underneath are
 * unpinned (documented in DESIGN.md).
 * Compile without floating-point contraction (-ffp-contract=off): every product and sum is rounded on its own.
 */
#define _GNU_SOURCE   /* M_PI under -std=c11 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct { float re, im; } cf;
static cf cmul(cf a, cf b) { cf r = {(a.re * b.re) - (a.im * b.im), (a.im * b.re) + (a.re * b.im)}; return r; }   /* complex_t::operator* */
static cf cconj(cf a) { cf r = {a.re, -a.im}; return r; }
static float camp(cf a) { return sqrtf((a.re * a.re) + (a.im * a.im)); }

#define SOF_VALUE 0x18d2e82u
#define SOF_LEN 26
#define PLS_LEN 64
#define PLS_SCRAMBLING 0x719d83c953422dfaull

static cf g_sof[SOF_LEN];
static cf g_pls[128][PLS_LEN];
static uint64_t g_code[128];
static int g_tables;

static void tables(void) /* s2_defs.h:23-30,41-83 */
{
    if (g_tables) return;
    for (int s = 0; s < SOF_LEN; ++s) {
        int bit = (SOF_VALUE >> (SOF_LEN - 1 - s)) & 1;
        int angle = bit * 2 + (s & 1);
        g_sof[s].re = 1 * cosf((float)(M_PI / 4 + 2 * M_PI * angle / 4));
        g_sof[s].im = 1 * sinf((float)(M_PI / 4 + 2 * M_PI * angle / 4));
    }
    static const uint32_t G[6] = {0x55555555u, 0x33333333u, 0x0f0f0f0fu, 0x00ff00ffu, 0x0000ffffu, 0xffffffffu};
    for (int index = 0; index < 128; ++index) {
        uint32_t y = 0;
        for (int row = 0; row < 6; ++row)
            if ((index >> (6 - row)) & 1) y ^= G[row];
        uint64_t code = 0;
        for (int bit = 31; bit >= 0; --bit) {
            uint64_t yi = (y >> bit) & 1;
            code = (index & 1) ? ((code << 2) | (yi << 1) | (yi ^ 1)) : ((code << 2) | (yi << 1) | yi);
        }
        code ^= PLS_SCRAMBLING;
        g_code[index] = code;
        for (int i = 0; i < PLS_LEN; ++i) {
            int yi = (int)((code >> (PLS_LEN - 1 - i)) & 1);
            int nyi = yi ^ (i & 1);
            g_pls[index][i].re = (float)(1 * (1 - 2 * nyi)) / sqrtf(2);
            g_pls[index][i].im = (float)(1 * (1 - 2 * yi)) / sqrtf(2);
        }
    }
    g_tables = 1;
}

void orc_plheader_symbols(int pls_code, float* out90)
{
    tables();
    for (int i = 0; i < SOF_LEN; ++i) { out90[2 * i] = g_sof[i].re; out90[2 * i + 1] = g_sof[i].im; }
    for (int i = 0; i < PLS_LEN; ++i) { out90[52 + 2 * i] = g_pls[pls_code & 127][i].re; out90[52 + 2 * i + 1] = g_pls[pls_code & 127][i].im; }
}
uint64_t orc_pls_codeword(int pls_code) { tables(); return g_code[pls_code & 127]; }

/* raw_frame_size of init / setParams (dvbs2_pl_sync.cpp:15-30): (slots + 1) * 90, plus 36 per pilot block */
int orc_raw_frame_size(int slot_num, int pilots)
{
    int rfs = (slot_num + 1) * 90;
    if (pilots) {
        int raw = (rfs - 90) / 90, cnt = 1;
        raw -= 16;
        while (raw > 16) {
            raw -= 16;
            ++cnt;
        }
        rfs += cnt * 36;
    }
    return rfs;
}

/* ------------------------------------------------------------------ S2PLSyncBlock */
struct orc_plsync {
    int rfs;
    cf* corr;      /* correlation_buffer[raw_frame_size] */
    cf* inbuf;     /* in_buffer: never holds more than raw_frame_size symbols */
    int ptr, lim, state, best_pos;
    int current_position;
    double best_match;
};

orc_plsync* orc_plsync_create(int slot_num, int pilots)
{
    tables();
    orc_plsync* p = (orc_plsync*)calloc(1, sizeof *p);
    p->rfs = orc_raw_frame_size(slot_num, pilots);
    p->corr = (cf*)calloc((size_t)p->rfs, sizeof(cf));
    p->inbuf = (cf*)calloc((size_t)p->rfs, sizeof(cf));
    p->lim = p->rfs;
    p->current_position = -1;
    return p;
}
void orc_plsync_destroy(orc_plsync* p)
{
    if (!p) return;
    free(p->corr);
    free(p->inbuf);
    free(p);
}

/* correlate_sof_diff / correlate_plscode_diff (:167-193) on diffs[k] = conj(x[k-1]) x[k], diffs[0] = 0 */
static cf plheader_metric(const cf* x) /* x: 90 symbols from the candidate position; returns d (:112-122) */
{
    cf diffs[SOF_LEN + PLS_LEN];
    /* volk_32fc_conjugate_32fc + volk_32fc_x2_multiply_32fc (:111-113); element 0 is (0 + 0i) x[0]: a zero of either
     * sign, which no later sum can tell apart */
    diffs[0].re = 0.f;
    diffs[0].im = 0.f;
    for (int k = 1; k < SOF_LEN + PLS_LEN; ++k) {
        const float ar = x[k - 1].re, ai = -x[k - 1].im, br = x[k].re, bi = x[k].im;
        diffs[k].re = ar * br - ai * bi;
        diffs[k].im = ar * bi + ai * br;
    }
    cf csof = {0.f, 0.f}, cpls = {0.f, 0.f};
    const uint32_t dsof = SOF_VALUE ^ (SOF_VALUE >> 1);
    for (int i = 0; i < SOF_LEN; ++i) {
        if (((dsof >> (SOF_LEN - 1 - i)) ^ (uint32_t)i) & 1) { csof.re += diffs[i].re; csof.im += diffs[i].im; }
        else { csof.re -= diffs[i].re; csof.im -= diffs[i].im; }
    }
    const uint64_t dscr = PLS_SCRAMBLING ^ (PLS_SCRAMBLING >> 1);
    for (int i = 1; i < PLS_LEN; i += 2) {
        if ((dscr >> (PLS_LEN - 1 - i)) & 1) { cpls.re -= diffs[SOF_LEN + i].re; cpls.im -= diffs[SOF_LEN + i].im; }
        else { cpls.re += diffs[SOF_LEN + i].re; cpls.im += diffs[SOF_LEN + i].im; }
    }
    cf c0 = {csof.re + cpls.re, csof.im + cpls.im}, c1 = {csof.re - cpls.re, csof.im - cpls.im};
    cf c = camp(c0) > camp(c1) ? c0 : c1;
    const float k = 1.0f / (26 - 1 + 64 / 2);
    cf d = {c.re * k, c.im * k};
    return d;
}

static int plsync_internal(orc_plsync* p, cf* out) /* internal_process (:102-165) */
{
    const int rfs = p->rfs;
    if (p->state == 0) {
        memcpy(p->corr, p->inbuf, (size_t)rfs * sizeof(cf));
        p->best_pos = 0;
        p->best_match = 0;
        for (int ss = 0; ss < rfs - SOF_LEN - PLS_LEN; ++ss) {
            cf d = plheader_metric(p->corr + ss);
            double difference = camp(d);
            if (difference > p->best_match && d.im > 0) {
                p->best_match = difference;
                p->best_pos = ss;
                p->current_position = ss;
            }
        }
        if (p->best_pos != 0 && p->best_pos < rfs) {
            p->lim = p->best_pos;
            p->state = 1;
            return 0;
        }
    } else {
        if (p->best_pos != 0 && p->best_pos < rfs) {
            const int pos = p->best_pos;
            memmove(p->corr, p->corr + pos, (size_t)(rfs - pos) * sizeof(cf));
            memcpy(p->corr + rfs - pos, p->inbuf, (size_t)pos * sizeof(cf));
            p->best_pos = 0;
        }
        p->lim = rfs;
        p->state = 0;
    }
    memcpy(out, p->corr, (size_t)rfs * sizeof(cf));
    return rfs;
}

int orc_plsync_process(orc_plsync* p, int count, const float* in, float* out) /* process (:80-100) */
{
    int outcnt = 0;
    const cf* x = (const cf*)in;
    for (int i = 0; i < count; ++i) {
        p->inbuf[p->ptr++] = x[i];
        if (p->ptr >= p->lim) {
            outcnt += plsync_internal(p, (cf*)out + outcnt);
            p->ptr = 0;
        }
    }
    return outcnt;
}
void orc_plsync_stats(const orc_plsync* p, int* raw_frame_size, int* current_position, double* best_match)
{
    if (raw_frame_size) *raw_frame_size = p->rfs;
    if (current_position) *current_position = p->current_position;
    if (best_match) *best_match = p->best_match;
}

/* ------------------------------------------------------------------ S2PLHDRDemod */
struct orc_plhdr {
    float alpha, beta, phase, freq;
};
#define FL_PI 3.1415926535f
static void pcl_advance(orc_plhdr* p, float error) /* PhaseControlLoop<float>::advance, limits of init (:10) */
{
    p->freq += p->beta * error;
    if (p->freq > 1.0f * FL_PI) p->freq = 1.0f * FL_PI;
    else if (p->freq < -1.0f * FL_PI) p->freq = -1.0f * FL_PI;
    p->phase += p->freq + (p->alpha * error);
    const float delta = FL_PI - (-FL_PI);
    while (p->phase > FL_PI) p->phase -= delta;
    while (p->phase < -FL_PI) p->phase += delta;
}
orc_plhdr* orc_plhdr_create(float loop_bw) /* init (:5-12): criticallyDamped(loop_bw * 0.03f) */
{
    tables();
    orc_plhdr* p = (orc_plhdr*)calloc(1, sizeof *p);
    const float bw = loop_bw * 0.03f;
    const float damp = (float)(sqrt(2.0) / 2.0);
    const float den = (float)(1.0 + 2.0 * damp * bw + bw * bw);
    p->alpha = (4 * damp * bw) / den;
    p->beta = (4 * bw * bw) / den;
    return p;
}
void orc_plhdr_destroy(orc_plhdr* p) { free(p); }

int orc_plhdr_process(orc_plhdr* p, int count, const float* in, float* out90, int* res, float* loop) /* process (:33-67) */
{
    const cf* x = (const cf*)in;
    cf* out = (cf*)out90;
    for (int i = 0; i < 90; ++i) {
        cf ph = {cosf(-p->phase), sinf(-p->phase)};
        cf t = cmul(x[i], ph);
        float error = (((t.re > 0.0) ? 1.0f : -1.0f) * t.im) - (((t.im > 0.0) ? 1.0f : -1.0f) * t.re);
        if (i & 1) { out[i].re = -t.re; out[i].im = t.im; }
        else { out[i].re = t.im; out[i].im = t.re; }
        pcl_advance(p, error);
    }
    p->phase += p->freq * (count - 91);
    pcl_advance(p, 0);
    uint64_t plheader = 0;
    const cf rot = {(float)cos(-M_PI / 4), (float)sin(-M_PI / 4)};
    for (int y = 0; y < 64; ++y) {
        int value = cmul(out[26 + y], rot).re > 0;
        plheader = plheader << 1 | (uint64_t)!value;
    }
    int best = 0, diffs = 64;
    for (int c = 0; c < 128; ++c) {
        int d = 0;
        for (int i = 59; i >= 0; --i) d += ((g_code[c] >> i) & 1) != ((plheader >> i) & 1);   /* checkSyncMarker (:69-79): bits 59..0 */
        if (d < diffs) { best = c; diffs = d; }
    }
    res[0] = (best >> 2) & 31;
    res[1] = (best & 2) >> 1;
    res[2] = best & 1;
    loop[0] = p->phase;
    loop[1] = p->freq;
    return count;
}

/* ------------------------------------------------------------------ dvbs2_pilot_coarse_fed (dvbs2_fed.h:7-48) */
float orc_coarse_fed(const float* frame_f, int raw_frame_size, int pilots, int pls_code, const uint8_t* rn)
{
    tables();
    const cf* frame = (const cf*)frame_f;
    const cf* pl = g_pls[pls_code & 127];
    float err = 0, symcnt = 90 - 2, sym_err;
    for (int i = 0; i < SOF_LEN - 2; ++i) {
        sym_err = cmul(cmul(cmul(frame[i + 2], cconj(g_sof[i + 2])), cconj(frame[i])), g_sof[i]).im;
        err += sym_err;
    }
    sym_err = cmul(cmul(cmul(frame[24 + 2], cconj(pl[24 - SOF_LEN + 2])), cconj(frame[24])), g_sof[24]).im;
    err += sym_err;
    sym_err = cmul(cmul(cmul(frame[25 + 2], cconj(pl[25 - SOF_LEN + 2])), cconj(frame[25])), g_sof[25]).im;
    err += sym_err;
    for (int i = SOF_LEN; i < 90 - 2; ++i) {
        sym_err = cmul(cmul(cmul(frame[i + 2], cconj(pl[i - SOF_LEN + 2])), cconj(frame[i])), pl[i - SOF_LEN]).im;
        err += sym_err;
    }
    if (pilots) {
        cf t1 = {0, 0}, t2 = {0, 0};
        const cf ref = {0.707f, 0.707f};
        for (int blk = 0; blk < ((raw_frame_size / 90 - 1) / 16) - 1; ++blk) {
            const int startsym = 90 * 17 + blk * (90 * 16 + 37);
            const cf* pst = frame + startsym;
            int pos = startsym - 90;
            for (int i = 0; i < 36; ++i) {
                cf p = pst[i], descr;
                switch (rn[pos++]) {
                case 3: descr.re = -p.im; descr.im = p.re; break;
                case 2: descr.re = -p.re; descr.im = -p.im; break;
                case 1: descr.re = p.im; descr.im = -p.re; break;
                default: descr = p; break;
                }
                if (i >= 2) {
                    sym_err = cmul(cmul(cmul(descr, cconj(ref)), cconj(t2)), ref).im;
                    err += sym_err;
                }
                t2 = t1;
                t1 = descr;
            }
            symcnt += 36 - 2;
        }
    }
    return err / symcnt;
}
