/* TEST INFRASTRUCTURE -- CPU restatement of the inner half of the reference's DVB-S chain:
 *   DVBSymToSoftBlock::process        dvbs/dvbs_syms_to_soft.cpp:26-42
 *   DVBSVitBlock::process             dvbs/dvbs_vit.cpp:6-13
 *   viterbi::Viterbi_DVBS             dvbs/viterbi_all.cpp:9-318, viterbi_all.h:33-163 (self-locking punctured K = 7 decoder)
 *   viterbi::CCDecoder / CCEncoder    dvbs/viterbi/cc_decoder.cpp:166-316, cc_encoder.cpp:92-118
 *   the bundled generic butterfly     dvbs/viterbi/volk_k7_r2_generic_fixed.h:25-115
 *   Depunc23 / Depunc56               dvbs/depunc.h:8-190
 *   rotate_soft, signed_soft_to_unsigned   common/codings/rotation.cpp:4-62, common/utils.cpp:11-20
 * Pinned bit for bit by tests/test_vit_oracle.py to those sources compiled unmodified (oracle/_ref; VOLK stand-in that
 * offers no "spiral" kernel, so it is the BUNDLED GENERIC kernel that is pinned -- a build of the reference against a
 * VOLK that has the spiral kernel runs that instead, cc_decoder.cpp:61-94, and is unpinned).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file; the product never does.
 *
 * What the reference does that a textbook decoder does not, all of it reproduced:
 *  - every call decodes ONE block of 8192 soft bits with freshly biased metrics (63 everywhere, 0 at the state the previous
 *    block's traceback reached at its step frame_size); six more trellis steps are read behind the block (erasures, or
 *    whatever an earlier call left there);
 *  - the decoders always run their constructed length whatever `size` says, so rate 5/6 (constructed with 1.66 for 5/3)
 *    leaves 27 of every 6826 output bits unwritten and restarts the trellis in the wrong place; rate 2/3 decodes one
 *    trellis step of stale data every third call;
 *  - the lock search reads past what its depuncturers wrote (stale bytes of earlier candidates) and, for rate 1/2, 13
 *    bytes past ber_soft_buffer into the member behind it; get_ber reads one re-encoded bit rate 5/6 never writes.
 *    The members are modelled as one array, and start from zeros (the reference leaves them indeterminate; the harness
 *    zeroes them). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define VIT_BUF 8192
#define TEST_BITS 2048

/* ---- CCDecoder ---- */
typedef struct {
    int fs, veclen;          /* d_frame_size, d_veclen = fs + 6 */
    int start_state;         /* d_start_state_chaining */
    uint8_t metrics[64];     /* old_metrics as the next call finds them */
    uint64_t* dec;           /* one 64-bit decision word per trellis step */
} ccdec;

static uint8_t g_branch[64];   /* Branchtab: [0..31] first polynomial (79), [32..63] second (109) */
static void branch_init(void)
{
    static const int polys[2] = {79, 109};
    for (int st = 0; st < 32; ++st)
        for (int j = 0; j < 2; ++j) g_branch[j * 32 + st] = __builtin_parity((2 * st) & polys[j]) ? 255 : 0;   /* cc_decoder.cpp:127-134 */
}
static void ccdec_init(ccdec* d, int fs)
{
    d->fs = fs;
    d->veclen = fs + 6;
    d->start_state = 0;
    memset(d->metrics, 31, 64);      /* init_viterbi_unbiased (:189-201): the very first block starts without a bias */
    d->dec = (uint64_t*)calloc((size_t)d->veclen, sizeof(uint64_t));
}
/* CCDecoder::work (:295-316): update_viterbi_blk over veclen steps, find_endstate, chainback, init_viterbi */
static void ccdec_work(ccdec* d, const uint8_t* in, uint8_t* out)
{
    uint8_t A[64], B[64];
    uint8_t *x = A, *y = B;
    memcpy(A, d->metrics, 64);
    for (int s = 0; s < d->veclen; ++s) {
        uint64_t w = 0;
        for (int i = 0; i < 32; ++i) {      /* BFLY (volk_k7_r2_generic_fixed.h:42-82) */
            unsigned sum = 1u + (g_branch[i] ^ in[2 * s]) + (g_branch[32 + i] ^ in[2 * s + 1]);
            uint8_t metric = (uint8_t)((sum >> 1) >> 2);
            uint8_t m0 = (uint8_t)(x[i] + metric), m1 = (uint8_t)(x[i + 32] + (63 - metric));
            uint8_t m2 = (uint8_t)(x[i] + (63 - metric)), m3 = (uint8_t)(x[i + 32] + metric);
            unsigned d0 = m0 >= m1, d1 = m2 >= m3;
            y[2 * i] = d0 ? m1 : m0;
            y[2 * i + 1] = d1 ? m3 : m2;
            w |= (uint64_t)(d0 | d1 << 1) << (2 * i);
        }
        uint8_t mn = y[0];
        for (int i = 1; i < 64; ++i) if (y[i] < mn) mn = y[i];      /* renormalize (:25-38) */
        for (int i = 0; i < 64; ++i) y[i] -= mn;
        d->dec[s] = w;
        uint8_t* t = x; x = y; y = t;
    }
    int state = 0;                                                  /* find_endstate (:203-220): first minimum */
    for (int i = 1; i < 64; ++i) if (x[i] < x[state]) state = i;
    int retval = 0;                                                 /* chainback_viterbi (:242-283) with tailsize 6 */
    for (int n = d->fs - 1; n >= 0; --n) {
        int k = (int)(d->dec[n + 6] >> state) & 1;
        state = (state >> 1) | (k << 5);
        out[n] = (uint8_t)k;
        if (n == d->fs - 6) retval = state;
    }
    d->start_state = retval;
    memset(d->metrics, 63, 64);                                     /* init_viterbi (:170-187) */
    d->metrics[retval & 63] = 0;
}

/* ---- CCEncoder::work (cc_encoder.cpp:92-104) ---- */
typedef struct { unsigned state; int fs; } ccenc;
static void ccenc_work(ccenc* e, const uint8_t* in, uint8_t* out)
{
    unsigned st = e->state;
    for (int i = 0; i < e->fs; ++i) {
        st = (st << 1) | (in[i] & 1u);
        out[2 * i] = (uint8_t)__builtin_parity(st & 79u);
        out[2 * i + 1] = (uint8_t)__builtin_parity(st & 109u);
    }
    e->state = st;
}

/* ---- Depunc23 / Depunc56 (depunc.h): which inputs are followed / preceded by an erasure ---- */
typedef struct { int period; int is_first, changing_shift, got_extra; uint8_t buf; } depunc;
static int depunc_emit(int period, int ph, uint8_t v, uint8_t* out)
{
    if (period == 3) {
        if (ph == 1) { out[0] = v; out[1] = 128; return 2; }
        out[0] = v; return 1;
    }
    switch (ph) {
    case 0: case 2: out[0] = v; return 1;
    case 4: out[0] = 128; out[1] = v; return 2;
    default: out[0] = v; out[1] = 128; return 2;   /* 1, 3, 5 */
    }
}
static int depunc_static(int period, const uint8_t* in, uint8_t* out, int size, int shift)
{
    int oo = 0, a = shift % period;
    if (shift > period - 1) out[oo++] = 128;
    for (int i = 0; i < size; ++i) oo += depunc_emit(period, (i + a) % period, in[i], out + oo);
    return oo;
}
static void depunc_set_shift(depunc* d, int shift)
{
    d->changing_shift = shift;
    d->is_first = shift > d->period - 1;
}
static int depunc_cont(depunc* d, const uint8_t* in, uint8_t* out, int size)
{
    int oo = 0;
    if (d->is_first || d->got_extra) {
        out[oo++] = d->buf;
        d->is_first = 0;
        d->got_extra = 0;
    }
    d->changing_shift %= d->period;
    for (int i = 0; i < size; ++i) {
        oo += depunc_emit(d->period, d->changing_shift % d->period, in[i], out + oo);
        d->changing_shift++;
    }
    if (oo % 2 == 1) {
        d->buf = out[oo - 1];
        oo -= 1;
        d->got_extra = 1;
    }
    return oo;
}
/* Viterbi_DVBS::depuncture_34 / depuncture_78 (viterbi_all.h:93-151) */
static int depuncture_34(const uint8_t* in, uint8_t* out, int size, int shift)
{
    int oo = 0;
    for (int i = 0; i < size / 2; ++i) {
        if ((shift != 0) ^ (i % 2 == 0)) { out[oo++] = in[2 * i]; out[oo++] = in[2 * i + 1]; }
        else { out[oo++] = 128; out[oo++] = in[2 * i]; out[oo++] = in[2 * i + 1]; out[oo++] = 128; }
    }
    return oo;
}
static int depuncture_78(const uint8_t* in, uint8_t* out, int size, int shift)
{
    int oo = 0;
    for (int i = 0; i < size / 2; ++i) {
        switch ((i + shift) % 4) {
        case 0: out[oo++] = in[2 * i]; out[oo++] = in[2 * i + 1]; break;
        case 1: out[oo++] = 128; out[oo++] = in[2 * i]; out[oo++] = 128; out[oo++] = in[2 * i + 1]; break;
        default: out[oo++] = 128; out[oo++] = in[2 * i]; out[oo++] = in[2 * i + 1]; out[oo++] = 128; break;
        }
    }
    return oo;
}

/* rotate_soft (rotation.cpp:4-62, iqswap false) then signed_soft_to_unsigned (utils.cpp:11-20) */
static void rotate_to_unsigned(const int8_t* in, uint8_t* out, int size, int phase)
{
    for (int i = 0; i < size; i += 2) {
        int a = in[i] == -128 ? -127 : in[i], b = in[i + 1] == -128 ? -127 : in[i + 1], ra = a, rb = b;
        switch (phase) {
        case 1: ra = b; rb = -a; break;
        case 2: ra = -a; rb = -b; break;
        case 3: ra = -b; rb = a; break;
        default: break;
        }
        uint8_t ua = (uint8_t)(ra + 127), ub = (uint8_t)(rb + 127);
        out[i] = ua == 128 ? 127 : ua;
        out[i + 1] = ub == 128 ? 127 : ub;
    }
}

/* Viterbi_DVBS::get_ber (viterbi_all.cpp:60-73) */
static float get_ber(const uint8_t* raw, const uint8_t* renc, int len, float ratio)
{
    float errors = 0, total = 0;
    for (int i = 0; i < len; ++i)
        if (raw[i] != 128) {
            errors += (raw[i] > 127) != renc[i];
            total++;
        }
    return (errors / total) * ratio;
}

enum { R12, R23, R34, R56, R78 };
static const int kShifts[5] = {2, 6, 2, 12, 4};
static const float kRatio[5] = {2.5f, 3.5f, 5.0f, 8.0f, 10.0f};

struct orc_vit {
    float thr, max_outsync;
    int state, rate, phase, shift, invalid;
    float bers[5][2][12];
    float ber;
    ccdec dec_ber[5], dec_main[5];
    ccenc enc_ber[5];
    int ber_len[5];
    depunc dp[5];                       /* only [R23], [R56] used */
    uint8_t ber_buf[TEST_BITS + 4 * TEST_BITS];   /* ber_soft_buffer followed by ber_depunc_buffer (viterbi_all.h:78-79) */
    uint8_t ber_decoded[4 * TEST_BITS], ber_encoded[4 * TEST_BITS];
    uint8_t soft[4 * VIT_BUF], depunc_buf[4 * VIT_BUF];
};

orc_vit* orc_vit_create(float ber_threshold, int max_outsync)
{
    static int once = 0;
    if (!once) { branch_init(); once = 1; }
    orc_vit* v = (orc_vit*)calloc(1, sizeof(orc_vit));
    v->thr = ber_threshold;
    v->max_outsync = (float)max_outsync;
    v->ber = 10;
    /* constructor (viterbi_all.cpp:17-33): the sizes are what the reference's double arithmetic truncates to */
    const int fs_ber[5] = {TEST_BITS / 2, (int)(TEST_BITS * 1.334 / 2), (int)(TEST_BITS * 1.5 / 2), (int)(TEST_BITS * 1.66 / 2), (int)(TEST_BITS * 1.75 / 2)};
    const int fs_main[5] = {VIT_BUF / 2, 10924 / 2, (int)(VIT_BUF * 1.5 / 2), (int)(VIT_BUF * 1.66 / 2), (int)(VIT_BUF * 1.75 / 2)};
    const int len[5] = {TEST_BITS, (int)(TEST_BITS * 1.25), (int)(TEST_BITS * 1.5), (int)(TEST_BITS * 1.66), (int)(TEST_BITS * 1.75)};
    for (int r = 0; r < 5; ++r) {
        ccdec_init(&v->dec_ber[r], fs_ber[r]);
        ccdec_init(&v->dec_main[r], fs_main[r]);
        v->enc_ber[r].fs = fs_ber[r];
        v->ber_len[r] = len[r];
        v->dp[r].buf = 128;
        for (int p = 0; p < 2; ++p)
            for (int s = 0; s < 12; ++s) v->bers[r][p][s] = 10;
    }
    v->dp[R23].period = 3;
    v->dp[R56].period = 6;
    return v;
}
void orc_vit_destroy(orc_vit* v)
{
    if (!v) return;
    for (int r = 0; r < 5; ++r) { free(v->dec_ber[r].dec); free(v->dec_main[r].dec); }
    free(v);
}

/* Viterbi_DVBS::work (viterbi_all.cpp:75-280); `in` is not modified (the reference rotates it in place) */
static int vit_work(orc_vit* v, const int8_t* in, uint8_t* out)
{
    uint8_t* bsoft = v->ber_buf;
    uint8_t* bdep = v->ber_buf + TEST_BITS;
    if (v->state == 0) {
        v->ber = 10;
        for (int phase = 0; phase < 2; ++phase) {
            rotate_to_unsigned(in, bsoft, TEST_BITS, phase);
            for (int r = 0; r < 5; ++r)
                for (int shift = 0; shift < kShifts[r]; ++shift) {
                    const uint8_t* raw = bdep;
                    switch (r) {
                    case R12: raw = bsoft + shift; break;
                    case R23: depunc_static(3, bsoft, bdep, TEST_BITS, shift); break;
                    case R34: depuncture_34(bsoft, bdep, TEST_BITS, shift); break;
                    case R56: depunc_static(6, bsoft, bdep, TEST_BITS, shift); break;
                    default: depuncture_78(bsoft, bdep, TEST_BITS, shift); break;
                    }
                    ccdec_work(&v->dec_ber[r], raw, v->ber_decoded);
                    ccenc_work(&v->enc_ber[r], v->ber_decoded, v->ber_encoded);
                    const float b = get_ber(raw, v->ber_encoded, v->ber_len[r], kRatio[r]);
                    v->bers[r][phase][shift] = b;
                    if (b < v->thr) {           /* every candidate under the threshold takes the lock: the last one keeps it */
                        v->ber = b;
                        v->state = 1;
                        v->phase = phase;
                        v->shift = shift;
                        v->invalid = 0;
                        v->rate = r;
                        if (r == R23 || r == R56) depunc_set_shift(&v->dp[r], shift);
                        memset(v->soft, 128, sizeof v->soft);
                        memset(v->depunc_buf, 128, sizeof v->depunc_buf);
                    }
                }
        }
    }
    int out_n = 0;
    if (v->state == 1) {
        const int r = v->rate;
        rotate_to_unsigned(in, v->soft, VIT_BUF, v->phase);
        const uint8_t* raw = v->depunc_buf;
        int sz = VIT_BUF;
        switch (r) {
        case R12: raw = v->soft + v->shift; break;
        case R23: sz = depunc_cont(&v->dp[R23], v->soft, v->depunc_buf, VIT_BUF); break;
        case R34: sz = depuncture_34(v->soft, v->depunc_buf, VIT_BUF, v->shift); break;
        case R56: sz = depunc_cont(&v->dp[R56], v->soft, v->depunc_buf, VIT_BUF); break;
        default: sz = depuncture_78(v->soft, v->depunc_buf, VIT_BUF, v->shift); break;
        }
        ccdec_work(&v->dec_main[r], raw, out);          /* writes its frame size, whatever sz is */
        out_n = sz / 2;
        ccenc_work(&v->enc_ber[r], out, v->ber_encoded);
        v->ber = get_ber(raw, v->ber_encoded, v->ber_len[r], kRatio[r]);
        if (v->ber > v->thr) {
            v->invalid++;
            if ((float)v->invalid > v->max_outsync) v->state = 0;
        } else
            v->invalid = 0;
    }
    return out_n;
}

/* DVBSVitBlock::process (dvbs_vit.cpp:6-13); count a multiple of 8192 */
int orc_vit_process(orc_vit* v, int count, const int8_t* in, uint8_t* out)
{
    int oidx = 0;
    for (int i = 0; i + VIT_BUF <= count; i += VIT_BUF) oidx += vit_work(v, in + i, out + oidx);
    return oidx;
}
/* Viterbi_DVBS::ber / getState / rate (viterbi_all.cpp:282-318) and the lock parameters */
void orc_vit_stats(const orc_vit* v, float* ber, int* state, int* rate, int* phase, int* shift, int* invalid)
{
    float b = v->ber;
    if (v->state != 1) {
        b = 10;
        for (int r = 0; r < 5; ++r)
            for (int p = 0; p < 2; ++p)
                for (int s = 0; s < kShifts[r]; ++s)
                    if (b > v->bers[r][p][s]) b = v->bers[r][p][s];
    }
    *ber = b; *state = v->state; *rate = v->rate; *phase = v->phase; *shift = v->shift; *invalid = v->invalid;
}

/* ---- DVBSymToSoftBlock::process (dvbs_syms_to_soft.cpp:8-42): clamp(re * 100), clamp(im * 100), handed on in chunks of
 * 8192 soft bits; what does not fill a chunk waits for the next call ---- */
struct orc_sts { int fill; int8_t buf[VIT_BUF]; };
orc_sts* orc_sts_create(void) { return (orc_sts*)calloc(1, sizeof(orc_sts)); }
void orc_sts_destroy(orc_sts* s) { free(s); }
static int8_t sts_clamp(float x)
{
    if (x < -127.0) return -127;
    if (x > 127.0) return 127;
    return (int8_t)x;
}
int orc_sts_process(orc_sts* s, int count, const float* syms, int8_t* out)
{
    int o = 0;
    for (int i = 0; i < count; ++i) {
        s->buf[s->fill] = sts_clamp(syms[2 * i] * 100);
        s->buf[s->fill + 1] = sts_clamp(syms[2 * i + 1] * 100);
        s->fill += 2;
        if (s->fill >= VIT_BUF) {
            memcpy(out + o, s->buf, VIT_BUF);
            o += VIT_BUF;
            s->fill -= VIT_BUF;
        }
    }
    return o;
}
