/* TEST INFRASTRUCTURE -- see oracle.h.  Restatement of the reference's outer-code stage.
 *
 *   repack       module_dvbs2_demod.cpp:357-360
 *   BCH wrapper  bbframe_bch.cpp:39-193 (code parameters), 380-405 (decode dispatch)
 *   BCH decoder  bch/bose_chaudhuri_hocquenghem_decoder.hh:40-143
 *   algebra      bch/reed_solomon_error_correction.hh:34-62 (Chien), 64-96 (Artin-Schreier table),
 *                98-130 (location finder), 132-218 (Forney), 220-277 (Berlekamp-Massey), 280-316
 *   field        bch/galois_field.hh:122-171 (tables), 173-362 (operators)
 *   descrambler  bbframe_descramble.cpp:122-143
 *
 * Field elements are plain unsigned ints; log/exp tables are built exactly as Tables::Tables()
 * does (log[0] = N, exp[N] = 0).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

int orc_code_params(int shortframe, int rate, int* N, int* K, int* kbch, int* bch_t, int* q, int* links_total);

typedef struct {
    int m, Q, N, poly;
    uint16_t *log, *exp, *imap; /* imap: ArtinSchreier table */
    int imap_ready;
} gf_t;

static gf_t g_gf16, g_gf14;

static void gf_init(gf_t* f, int m, int poly)
{
    if (f->log)
        return;
    f->m = m;
    f->Q = 1 << m;
    f->N = f->Q - 1;
    f->poly = poly;
    f->log = (uint16_t*)malloc(sizeof(uint16_t) * f->Q);
    f->exp = (uint16_t*)malloc(sizeof(uint16_t) * f->Q);
    f->imap = (uint16_t*)calloc(f->Q, sizeof(uint16_t));
    f->exp[f->N] = 0;
    f->log[0] = (uint16_t)f->N;
    unsigned a = 1;
    for (int i = 0; i < f->N; ++i) {
        f->exp[i] = (uint16_t)a;
        f->log[a] = (uint16_t)i;
        a = (a & (f->Q >> 1)) ? (((a << 1) ^ poly) & 0xFFFF) : (a << 1); /* Tables::next, TYPE = uint16_t */
    }
}

/* Index arithmetic modulo N (galois_field.hh:220-229, 249-258) */
static inline int idx_mul(const gf_t* f, int a, int b) { int t = a + b; return t >= f->N ? t - f->N : t; }
static inline int idx_div(const gf_t* f, int a, int b) { int t = a - b; return t < 0 ? t + f->N : t; }
static inline unsigned gmul(const gf_t* f, unsigned a, unsigned b)
{
    return (!a || !b) ? 0 : f->exp[idx_mul(f, f->log[a], f->log[b])];
}
static inline unsigned gdiv(const gf_t* f, unsigned a, unsigned b) /* b != 0 */
{
    return !a ? 0 : f->exp[idx_div(f, f->log[a], f->log[b])];
}

/* ArtinSchreier ctor, reed_solomon_error_correction.hh:70-88: imap[x*x+x] = x for even x in
 * [2, N), skipping the one x whose image is the all-ones element */
static void gf_build_imap(gf_t* f)
{
    if (f->imap_ready)
        return;
    for (int x = 2; x < f->N; x += 2) {
        unsigned y = gmul(f, x, x) ^ (unsigned)x;
        if (y == (unsigned)f->N)
            continue;
        f->imap[y] = (uint16_t)x;
    }
    f->imap_ready = 1;
}

static inline int get_bit(const uint8_t* p, int i) { return (p[i >> 3] >> (7 - (i & 7))) & 1; }
static inline void flip_bit(uint8_t* p, int i) { p[i >> 3] ^= (uint8_t)(1 << (7 - (i & 7))); }

void orc_repack(const int8_t* llr, int nbits, uint8_t* bytes)
{
    memset(bytes, 0, (size_t)(nbits + 7) / 8);
    for (int i = 0; i < nbits; ++i)
        bytes[i >> 3] = (uint8_t)((bytes[i >> 3] << 1) | (llr[i] < 0));
}

/* compute_syndromes (bose_..._decoder.hh:40-71): S_i = c(alpha^(1+i)) by Horner, data then parity */
static int syndromes(const gf_t* f, const uint8_t* data, const uint8_t* parity, int data_len, int NP, int NR,
                     unsigned* S)
{
    for (int i = 0; i < NR; ++i)
        S[i] = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const uint8_t* p = pass ? parity : data;
        int n = pass ? NP : data_len;
        for (int j = 0; j < n; ++j) {
            unsigned coeff = (unsigned)get_bit(p, j);
            for (int i = 0; i < NR; ++i) {
                /* fma(root, S, coeff) with root = alpha^(1+i): S*alpha^(1+i) + coeff */
                unsigned s = S[i];
                S[i] = (s ? f->exp[idx_mul(f, 1 + i, f->log[s])] : 0) ^ coeff;
            }
        }
    }
    int nz = 0;
    for (int i = 0; i < NR; ++i)
        nz += S[i] != 0;
    return nz;
}

/* BerlekampMassey::algorithm with count = 0 (reed_solomon_error_correction.hh:225-276) */
static int berlekamp_massey(const gf_t* f, const unsigned* s, unsigned* C, int NR)
{
    unsigned B[32], T[32];
    for (int i = 0; i <= NR; ++i)
        B[i] = C[i];
    int L = 0;
    for (int n = 0, m = 1; n < NR; ++n) {
        unsigned d = s[n];
        for (int i = 1; i <= L; ++i)
            d ^= gmul(f, C[i], s[n - i]);
        if (!d) {
            ++m;
        } else {
            for (int i = 0; i < m; ++i)
                T[i] = C[i];
            for (int i = m; i <= NR; ++i)
                T[i] = gmul(f, d, B[i - m]) ^ C[i];
            if (2 * L <= n) {
                L = n + 1 - L;
                for (int i = 0; i <= NR; ++i)
                    B[i] = gdiv(f, C[i], d);
                m = 1;
            } else {
                ++m;
            }
            for (int i = 0; i <= NR; ++i)
                C[i] = T[i];
        }
    }
    return L;
}

/* LocationFinder::operator() (reed_solomon_error_correction.hh:103-129); locations are Index values */
static int locations(const gf_t* f, const unsigned* loc, int deg, int* out)
{
    if (deg == 1) {
        out[0] = idx_div(f, idx_div(f, f->log[loc[0]], f->log[loc[1]]), 1);
        return 1;
    }
    if (deg == 2) {
        if (!loc[1] || !loc[0])
            return 0;
        unsigned a = loc[2], b = loc[1], c = loc[0];
        unsigned ba = gdiv(f, b, a);
        unsigned arg = gdiv(f, gmul(f, a, c), gmul(f, b, b));
        unsigned R = f->imap[arg];
        if (!R)
            return 0;
        out[0] = idx_div(f, f->log[gmul(f, ba, R)], 1);
        out[1] = idx_div(f, f->log[gmul(f, ba, R) ^ ba], 1);
        return 2;
    }
    /* Chien::search (:39-61): position i tests locator(alpha^(i+1)) */
    unsigned tmp[32];
    for (int i = 0; i <= deg; ++i)
        tmp[i] = loc[i];
    int count = 0;
    for (int i = 0; i < f->N; ++i) {
        unsigned sum = tmp[0];
        for (int j = 1; j <= deg; ++j) {
            tmp[j] = tmp[j] ? f->exp[idx_mul(f, f->log[tmp[j]], j)] : 0;
            sum ^= tmp[j];
        }
        if (!sum)
            out[count++] = i;
    }
    return count;
}

/* Forney::compute_evaluator + compute_magnitudes, FCR = 1 (:137-204) */
static void forney(const gf_t* f, const unsigned* S, const unsigned* loc, const int* locs, int count, int NR,
                   unsigned* mags)
{
    unsigned ev[32];
    int top = count < NR - 1 ? count : NR - 1, evdeg = -1;
    for (int i = 0; i <= top; ++i) {
        ev[i] = gmul(f, S[i], loc[0]);
        for (int j = 1; j <= i; ++j)
            ev[i] ^= gmul(f, S[i - j], loc[j]);
        if (ev[i])
            evdeg = i;
    }
    for (int i = 0; i < count; ++i) {
        int root = idx_mul(f, locs[i], 1), tmp = root;
        unsigned eval = ev[0];
        for (int j = 1; j <= evdeg; ++j) {
            eval ^= ev[j] ? f->exp[idx_mul(f, f->log[ev[j]], tmp)] : 0;
            tmp = idx_mul(f, tmp, root);
        }
        if (!eval) {
            mags[i] = 0;
            continue;
        }
        unsigned deriv = loc[1];
        int root2 = idx_mul(f, root, root), tmp2 = root2;
        for (int j = 3; j <= count; j += 2) {
            deriv ^= loc[j] ? f->exp[idx_mul(f, f->log[loc[j]], tmp2)] : 0;
            tmp2 = idx_mul(f, tmp2, root2);
        }
        /* index(eval) / index(deriv): deriv == 0 would read log[0] = N in the reference */
        mags[i] = f->exp[idx_div(f, f->log[eval], f->log[deriv])];
    }
}

typedef struct { const gf_t* f; int K, NP, NR; } bch_code;

static int bch_setup(int shortframe, int rate, bch_code* bc, int* kbch)
{
    int K, t;
    if (orc_code_params(shortframe, rate, 0, &K, kbch, &t, 0, 0) < 0)
        return -1;
    gf_init(&g_gf16, 16, 0x1002D);
    gf_init(&g_gf14, 14, 0x402B);
    gf_build_imap(&g_gf16);
    gf_build_imap(&g_gf14);
    bc->f = shortframe ? &g_gf14 : &g_gf16;
    bc->NR = 2 * t;
    bc->NP = K - *kbch;
    bc->K = bc->f->N - bc->NP; /* template MSG: 65343 / 65375 / 65407 / 16215 */
    return 0;
}

int orc_bch_decode(int shortframe, int rate, uint8_t* frame)
{
    bch_code bc;
    int kbch;
    if (bch_setup(shortframe, rate, &bc, &kbch) < 0)
        return -2;
    const gf_t* f = bc.f;
    uint8_t* data = frame;
    uint8_t* parity = frame + kbch / 8;
    unsigned S[24];
    if (!syndromes(f, data, parity, kbch, bc.NP, bc.NR, S))
        return 0;
    /* ReedSolomonErrorCorrection::operator() (:289-316), no erasures */
    unsigned loc[32];
    memset(loc, 0, sizeof(loc));
    loc[0] = 1;
    int deg = berlekamp_massey(f, S, loc, bc.NR);
    while (!loc[deg])
        if (--deg < 0)
            return -1;
    int locs[64]; /* a degree-deg polynomial has at most deg <= 24 roots */
    int count = locations(f, loc, deg, locs);
    if (count < deg)
        return -1;
    unsigned mags[32];
    forney(f, S, loc, locs, count, bc.NR, mags);
    if (count <= 0)
        return count;
    for (int i = 0; i < count; ++i)
        if (locs[i] < bc.K - kbch)
            return -1;
    for (int i = 0; i < count; ++i)
        if (mags[i] > 1)
            return -1;
    for (int i = 0; i < count; ++i) {
        int idx = locs[i] + kbch - bc.K;
        if (mags[i]) {
            if (idx < kbch)
                flip_bit(data, idx);
            else
                flip_bit(parity, idx - kbch);
        }
    }
    int corrected = 0;
    for (int i = 0; i < count; ++i)
        corrected += mags[i] != 0;
    return corrected;
}

/* Generator g(x) = lcm of the minimal polynomials of alpha^1..alpha^2t over GF(2); systematic
 * encode = remainder of data(x)*x^NP.  Produces the same codewords as bbframe_bch.cpp:407-456. */
int orc_bch_encode(int shortframe, int rate, uint8_t* frame)
{
    bch_code bc;
    int kbch;
    if (bch_setup(shortframe, rate, &bc, &kbch) < 0)
        return -2;
    const gf_t* f = bc.f;
    /* build g(x) as bit array, degree NP */
    uint8_t g[200];
    memset(g, 0, sizeof(g));
    g[0] = 1;
    int gdeg = 0;
    uint8_t* seen = (uint8_t*)calloc((size_t)f->Q, 1);
    for (int r = 1; r <= bc.NR; ++r) {
        if (seen[r])
            continue;
        /* minimal polynomial of alpha^r: prod over conjugates (x - alpha^(r*2^k)) */
        unsigned mp[20];
        int mdeg = 0;
        mp[0] = 1;
        int e = r;
        do {
            seen[e] = 1;
            unsigned root = f->exp[e];
            mp[mdeg + 1] = 0;
            for (int k = mdeg + 1; k > 0; --k)
                mp[k] = mp[k - 1] ^ gmul(f, mp[k], root);
            mp[0] = gmul(f, mp[0], root);
            ++mdeg;
            e = (e * 2) % f->N;
        } while (e != r);
        uint8_t ng[200];
        memset(ng, 0, sizeof(ng));
        for (int i = 0; i <= gdeg; ++i)
            if (g[i])
                for (int k = 0; k <= mdeg; ++k)
                    ng[i + k] ^= (uint8_t)(mp[k] & 1); /* coefficients are 0/1 */
        gdeg += mdeg;
        memcpy(g, ng, sizeof(g));
    }
    free(seen);
    if (gdeg != bc.NP)
        return -3;
    uint8_t reg[200];
    memset(reg, 0, sizeof(reg)); /* reg[k] = coefficient of x^k of the running remainder */
    for (int i = 0; i < kbch; ++i) {
        int fb = get_bit(frame, i) ^ reg[gdeg - 1];
        for (int k = gdeg - 1; k > 0; --k)
            reg[k] = (uint8_t)(reg[k - 1] ^ (fb & g[k]));
        reg[0] = (uint8_t)(fb & g[0]);
    }
    uint8_t* parity = frame + kbch / 8;
    memset(parity, 0, (size_t)bc.NP / 8);
    for (int i = 0; i < bc.NP; ++i)
        if (reg[gdeg - 1 - i])
            flip_bit(parity, i);
    return 0;
}

/* BBFrameDescrambler::init/work (bbframe_descramble.cpp:122-143): PRBS 1+x^14+x^15 */
int orc_descramble(int shortframe, int rate, uint8_t* frame)
{
    int kbch;
    if (orc_code_params(shortframe, rate, 0, 0, &kbch, 0, 0, 0) < 0)
        return -2;
    int sr = 0x4A80;
    for (int i = 0; i < kbch; ++i) {
        int b = (sr ^ (sr >> 1)) & 1;
        if (b)
            flip_bit(frame, i);
        sr >>= 1;
        if (b)
            sr |= 0x4000;
    }
    return 0;
}

/* module_dvbs2_demod.cpp:349-366 for one frame, with the per-frame LDPC call of SURVEY note N1 */
int orc_decode_frame(int shortframe, int rate, int8_t* llr, int max_trials, uint8_t* bb, int* ldpc_iters, int* bch_corr)
{
    int K, kbch;
    if (orc_code_params(shortframe, rate, 0, &K, &kbch, 0, 0, 0) < 0)
        return -2;
    int it = orc_ldpc_decode(shortframe, rate, llr, max_trials);
    uint8_t* buf = (uint8_t*)malloc((size_t)K / 8);
    orc_repack(llr, K, buf);
    int corr = orc_bch_decode(shortframe, rate, buf);
    orc_descramble(shortframe, rate, buf);
    memcpy(bb, buf, (size_t)kbch / 8);
    free(buf);
    if (ldpc_iters) *ldpc_iters = it;
    if (bch_corr) *bch_corr = corr;
    return 0;
}

/* BBFrameTSParser::check_crc8 (dvbs2/bbframe_ts_parser.cpp:66-80) over the 80 BBHEADER bits */
unsigned orc_bbheader_crc8(const uint8_t* bbframe)
{
    unsigned crc = 0;
    for (int n = 0; n < 80; ++n) {
        unsigned b = (unsigned)get_bit(bbframe, n) ^ (crc & 1u);
        crc >>= 1;
        if (b)
            crc ^= 0xAB;
    }
    return crc;
}
