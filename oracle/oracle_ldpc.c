/* TEST INFRASTRUCTURE -- see oracle.h.  Scalar restatement of the reference's LDPC stage.
 *
 *   table expansion   ldpc.hh:25-109 (LDPC<TABLE> iterator), layered_decoder.hh:79-120 (init)
 *   decoder           layered_decoder.hh:23-74,121-133 (reset/bad/update/operator())
 *   check node        algorithms.hh:206-277 (OffsetMinSumAlgorithm<SIMD<int8_t,W>,NormalUpdate,2>)
 *   wrapper           bbframe_ldpc.cpp:123-139 (lane 0 only, blocks = 1)
 *   encoder           encoder.hh:37-52
 *
 * One frame at a time, every int8 operation spelled out (no SIMD), rows visited strictly in the
 * reference's order: layer i = 0..q-1, row j = 0..359.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define GROUP 360

typedef struct {
    const char* tag;
    const char* name;
    int N, K, nruns;
    int deg[3], len[3];
    int off;
} code_desc;

#define S2_CODE(tag, name, N, K, nruns, d0, l0, d1, l1, d2, l2, off) \
    {#tag, name, N, K, nruns, {d0, d1, d2}, {l0, l1, l2}, off},
static const code_desc g_codes[] = {
#include "../sdrpp-dvbs-demodulator_b200/csrc/s2_ldpc_addr.inc"
};
#undef S2_CODE
#define S2_CODE(...)
#define S2_ADDR_POOL
static const uint16_t g_pool[] = {
#include "../sdrpp-dvbs-demodulator_b200/csrc/s2_ldpc_addr.inc"
};
#undef S2_ADDR_POOL

/* reference rate enum -> index into g_codes (B1..B11 then C1..C10); -1 = no table (note N4) */
static int code_index(int shortframe, int rate)
{
    static const int normal_idx[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, -1, 9, 10};
    static const int short_idx[12] = {11, 12, 13, 14, 15, 16, 17, 18, 19, -1, 20, -1};
    if (rate < 0 || rate > 11)
        return -1;
    return shortframe ? short_idx[rate] : normal_idx[rate];
}

typedef struct {
    int ready;
    int N, K, R, q, CNL, LT;
    uint16_t* pos; /* [R][CNL], layered row order */
    uint8_t* cnc;  /* [R], original check order (decoder indexes it by layer, layered_decoder.hh:30,50) */
} sched;
static sched g_sched[21];

/* bbframe_bch.cpp:41-160: kbch and t per code */
static void bch_params(int shortframe, int rate, int K, int* kbch, int* t)
{
    if (shortframe) {
        *t = 12;
        *kbch = K - 168;
    } else if (rate == 5 || rate == 8) {
        *t = 10;
        *kbch = K - 160;
    } else if (rate == 10 || rate == 11) {
        *t = 8;
        *kbch = K - 128;
    } else {
        *t = 12;
        *kbch = K - 192;
    }
}

/* Walk the address table the way LDPC<TABLE>::first_bit/next_bit do (ldpc.hh:36-107): bit m of
 * group g with table row x[] hits checks (x[d] + q*m) mod R.  cb(bit, check) is called for every
 * edge, bits ascending. */
static int links_max_cn(const code_desc* c);
static sched* get_sched(int idx)
{
    sched* s = &g_sched[idx];
    if (s->ready)
        return s;
    const code_desc* c = &g_codes[idx];
    s->N = c->N;
    s->K = c->K;
    s->R = c->N - c->K;
    s->q = s->R / GROUP;
    s->CNL = links_max_cn(c) - 2;
    s->pos = (uint16_t*)calloc((size_t)s->R * s->CNL, sizeof(uint16_t));
    s->cnc = (uint8_t*)calloc((size_t)s->R, 1);
    uint16_t* tmp = (uint16_t*)calloc((size_t)s->R * s->CNL, sizeof(uint16_t));
    const uint16_t* row = g_pool + c->off;
    int bit = 0, links = 0;
    for (int r = 0; r < c->nruns; ++r)
        for (int g = 0; g < c->len[r]; ++g) {
            for (int m = 0; m < GROUP; ++m, ++bit)
                for (int d = 0; d < c->deg[r]; ++d) {
                    int chk = (row[d] + s->q * m) % s->R;
                    tmp[s->CNL * chk + s->cnc[chk]++] = (uint16_t)bit; /* layered_decoder.hh:103-110 */
                    ++links;
                }
            row += c->deg[r];
        }
    /* permute checks to layered order (layered_decoder.hh:115-119): row (i,j) <- check q*j+i */
    for (int i = 0; i < s->q; ++i)
        for (int j = 0; j < GROUP; ++j)
            for (int k = 0; k < s->CNL; ++k)
                s->pos[s->CNL * (GROUP * i + j) + k] = tmp[s->CNL * (s->q * j + i) + k];
    free(tmp);
    s->LT = links + 2 * s->R - 1; /* TABLE::LINKS_TOTAL */
    s->ready = 1;
    return s;
}

static int links_max_cn(const code_desc* c)
{
    int R = c->N - c->K, q = R / GROUP, mx = 0;
    uint8_t* cnt = (uint8_t*)calloc((size_t)R, 1);
    const uint16_t* row = g_pool + c->off;
    for (int r = 0; r < c->nruns; ++r)
        for (int g = 0; g < c->len[r]; ++g) {
            for (int m = 0; m < GROUP; ++m)
                for (int d = 0; d < c->deg[r]; ++d)
                    cnt[(row[d] + q * m) % R]++;
            row += c->deg[r];
        }
    for (int i = 0; i < R; ++i)
        if (cnt[i] > mx)
            mx = cnt[i];
    free(cnt);
    return mx + 2; /* TABLE::LINKS_MAX_CN counts the two parity links */
}

int orc_code_params(int shortframe, int rate, int* N, int* K, int* kbch, int* bch_t, int* q, int* links_total)
{
    int idx = code_index(shortframe, rate);
    if (idx < 0)
        return -1;
    sched* s = get_sched(idx);
    int kb, t;
    bch_params(shortframe, rate, s->K, &kb, &t);
    if (N) *N = s->N;
    if (K) *K = s->K;
    if (kbch) *kbch = kb;
    if (bch_t) *bch_t = t;
    if (q) *q = s->q;
    if (links_total) *links_total = s->LT;
    return idx;
}

int orc_ldpc_schedule(int shortframe, int rate, uint16_t* pos, uint8_t* cnc)
{
    int idx = code_index(shortframe, rate);
    if (idx < 0)
        return -1;
    sched* s = get_sched(idx);
    if (pos) memcpy(pos, s->pos, (size_t)s->R * s->CNL * sizeof(uint16_t));
    if (cnc) memcpy(cnc, s->cnc, (size_t)s->q);
    return s->CNL;
}

/* ---- int8 lane arithmetic of SIMD<int8_t,W> (sse4_1.hh / simd.hh), one lane ---- */
static inline int8_t sat8(int x) { return (int8_t)(x < -128 ? -128 : x > 127 ? 127 : x); }
static inline int8_t qadd(int8_t a, int8_t b) { return sat8((int)a + (int)b); } /* vqadd */
static inline int8_t qsub(int8_t a, int8_t b) { return sat8((int)a - (int)b); } /* vqsub */
static inline int8_t qabs(int8_t a) { return (int8_t)(a == -128 ? 127 : a < 0 ? -a : a); } /* vqabs */
/* vsign(a,b): b<0 -> -a, b==0 -> 0, b>0 -> a (sse4_1.hh _mm_sign_epi8) */
static inline int8_t vsign(int8_t a, int8_t b) { return (int8_t)(b < 0 ? -a : b > 0 ? a : 0); }

/* OffsetMinSumAlgorithm::finalp, algorithms.hh:235-256, beta = nearbyint(0.5*2) = 1 */
static void finalp(int8_t* links, int cnt)
{
    int8_t mags[32];
    for (int i = 0; i < cnt; ++i) {
        int m = (int)(uint8_t)qabs(links[i]) - 1; /* unsigned saturating subtract of beta */
        mags[i] = (int8_t)(m < 0 ? 0 : m);
    }
    int8_t min0 = mags[0] < mags[1] ? mags[0] : mags[1];
    int8_t min1 = mags[0] < mags[1] ? mags[1] : mags[0];
    for (int i = 2; i < cnt; ++i) {
        int8_t hi = min0 > mags[i] ? min0 : mags[i];
        if (hi < min1) min1 = hi;
        if (mags[i] < min0) min0 = mags[i];
    }
    int8_t signs = links[0];
    for (int i = 1; i < cnt; ++i)
        signs = (int8_t)(signs ^ links[i]);
    for (int i = 0; i < cnt; ++i) {
        int8_t other = mags[i] == min0 ? min1 : min0;      /* other(): vbsl(vceq(a,b), c, b) */
        int8_t s = (int8_t)((signs ^ links[i]) | 127);     /* never zero: sign only */
        links[i] = vsign(other, s);
    }
}
/* OffsetMinSumAlgorithm::update -> NormalUpdate (algorithms.hh:273-276, generic.hh:15-18) */
static inline int8_t msg_clamp(int8_t b) { return (int8_t)(b < -32 ? -32 : b > 31 ? 31 : b); }

/* LDPCDecoder::bad (layered_decoder.hh:28-45), blocks = 1 */
static int is_bad(const sched* s, const int8_t* data, const int8_t* pty)
{
    const int M = GROUP, q = s->q;
    for (int i = 0; i < q; ++i) {
        int cnt = s->cnc[i];
        for (int j = 0; j < M; ++j) {
            int8_t cnv = vsign(1, pty[M * i + j]);
            if (i)
                cnv = vsign(cnv, pty[M * (i - 1) + j]);
            else if (j)
                cnv = vsign(cnv, pty[j + (q - 1) * M - 1]);
            for (int c = 0; c < cnt; ++c)
                cnv = vsign(cnv, data[s->pos[s->CNL * (M * i + j) + c]]);
            if (cnv <= 0)
                return 1;
        }
    }
    return 0;
}

/* LDPCDecoder::update (layered_decoder.hh:46-74) */
static void update(const sched* s, int8_t* data, int8_t* pty, int8_t* bnl)
{
    const int M = GROUP, q = s->q;
    int8_t* bl = bnl;
    for (int i = 0; i < q; ++i) {
        int cnt = s->cnc[i];
        for (int j = 0; j < M; ++j) {
            int deg = cnt + 2 - !(i | j);
            int8_t inp[32], out[32];
            int8_t* lnk[32];
            for (int c = 0; c < cnt; ++c)
                lnk[c] = &data[s->pos[s->CNL * (M * i + j) + c]];
            lnk[cnt] = &pty[M * i + j];
            if (i)
                lnk[cnt + 1] = &pty[M * (i - 1) + j];
            else if (j)
                lnk[cnt + 1] = &pty[j + (q - 1) * M - 1];
            for (int d = 0; d < deg; ++d)
                inp[d] = out[d] = qsub(*lnk[d], bl[d]);
            finalp(out, deg);
            for (int d = 0; d < deg; ++d)
                bl[d] = msg_clamp(out[d]);
            for (int d = 0; d < deg; ++d)
                *lnk[d] = qadd(inp[d], bl[d]);
            bl += deg;
        }
    }
}

int orc_ldpc_decode(int shortframe, int rate, int8_t* llr, int max_trials)
{
    int idx = code_index(shortframe, rate);
    if (idx < 0)
        return -2;
    const sched* s = get_sched(idx);
    const int M = GROUP, q = s->q;
    int8_t* bnl = (int8_t*)calloc((size_t)s->LT, 1); /* reset(): all messages zero */
    int8_t* pty = (int8_t*)malloc((size_t)s->R);
    int8_t* data = llr;
    int8_t* parity = llr + s->K;
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < M; ++j)
            pty[M * i + j] = parity[q * j + i];
    int trials = max_trials;
    while (is_bad(s, data, pty) && --trials >= 0)
        update(s, data, pty, bnl);
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < M; ++j)
            parity[q * j + i] = pty[M * i + j];
    free(bnl);
    free(pty);
    return trials < 0 ? trials : max_trials - trials; /* bbframe_ldpc.cpp:135-138 */
}

/* encoder.hh:37-52 in the bit domain: parity accumulators then the running XOR */
int orc_ldpc_encode_bits(int shortframe, int rate, const uint8_t* data_bits, uint8_t* code_bits)
{
    int idx = code_index(shortframe, rate);
    if (idx < 0)
        return -1;
    const code_desc* c = &g_codes[idx];
    int R = c->N - c->K, q = R / GROUP;
    uint8_t* par = code_bits + c->K;
    memcpy(code_bits, data_bits, (size_t)c->K);
    memset(par, 0, (size_t)R);
    const uint16_t* row = g_pool + c->off;
    int bit = 0;
    for (int r = 0; r < c->nruns; ++r)
        for (int g = 0; g < c->len[r]; ++g) {
            for (int m = 0; m < GROUP; ++m, ++bit)
                for (int d = 0; d < c->deg[r]; ++d)
                    par[(row[d] + q * m) % R] ^= data_bits[bit] & 1;
            row += c->deg[r];
        }
    for (int i = 1; i < R; ++i)
        par[i] ^= par[i - 1];
    return 0;
}
