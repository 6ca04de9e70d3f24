/* TEST INFRASTRUCTURE -- see oracle.h.  Restatement of the reference's symbols-to-soft stage.
 *
 *   constellation points   common/dsp/demod/constellation.cpp:19-150 (ctor), 14-17 (polar)
 *   LLR computation        constellation.cpp:205-261 (demod_soft_calc), 263-270 (clamp)
 *   LUT                    constellation.cpp:272-291 (make_lut), 293-322 (demod_soft_lut)
 *   modulator              constellation.cpp:156-158
 *   deinterleaver          dvbs2/codings/s2_deinterleaver.cpp:6-66 (ctor), 72-136
 *   frame loop             dvbs2/dvbs2_bb_to_soft.cpp:7-33 (pilots off)
 *
 * Float arithmetic follows the reference expression by expression (float vs double promotion
 * included) so that, with the same libm, the int8 outputs are identical.  complex_t semantics are
 * those of oracle/shim/dsp/types.h ("parity unpinned" against the real SDR++ header).
 */
#include "oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_D 3.14159265358979323846
#define SQRT2_D 1.41421356237309504880

struct orc_constellation {
    int type, bits, states;
    float amp, sca, prescale;
    float re[32], im[32];
    int8_t* lut; /* [256][256][bits] */
    float* perr; /* [256][256]: SoftResult::phase_error of the cell (constellation.cpp:284-288) */
};

/* constellation_t::polar: float a = i*2*M_PI/n (double product rounded to float) */
static void polar(float r, int n, float i, float amp, float* re, float* im)
{
    float a = (float)(i * 2 * PI_D / n);
    *re = (r * cosf(a)) * amp;
    *im = (r * sinf(a)) * amp;
}

static void set_ring(orc_constellation* c, const int* idx, const float* k, int cnt, float r, int n)
{
    for (int t = 0; t < cnt; ++t)
        polar(r, n, k[t], c->amp, &c->re[idx[t]], &c->im[idx[t]]);
}

static int8_t clamp8(float x)
{
    while (x < -127 || x > 127) {
        x *= 0.5;
        if (!isfinite(x))
            return (int8_t)(int)x; /* reference: implementation-defined conversion of inf/nan */
    }
    return (int8_t)x;
}

void orc_demod_soft_calc(const orc_constellation* c, float re, float im, int8_t* bits)
{
    float tmp[10];
    for (int i = 0; i < 2 * c->bits; ++i)
        tmp[i] = 0;
    if (c->amp != 1) {
        re = re * c->amp;
        im = im * c->amp;
    }
    if (c->prescale != 1) {
        re = re * c->prescale;
        im = im * c->prescale;
    }
    for (int i = 0; i < c->states; ++i) {
        float dr = re - c->re[i], di = im - c->im[i];
        float dist = sqrtf((dr * dr) + (di * di));
        float d = expf(-dist / 1.0f); /* npwr = 1.0 */
        for (int j = 0; j < c->bits; ++j) {
            if (((i >> j) & 1) == 0)
                tmp[2 * j + 0] += d;
            else
                tmp[2 * j + 1] += d;
        }
    }
    for (int i = 0; i < c->bits; ++i)
        bits[c->bits - 1 - i] = clamp8((logf(tmp[2 * i + 1]) - logf(tmp[2 * i + 0])) * c->sca);
}

/* the phase_error output of demod_soft_calc (constellation.cpp:209-231,258-260): the sample, scaled like the LLR
 * part, against the nearest point (the first one among equals: "dist < min_dist") */
float orc_demod_phase_error_calc(const orc_constellation* c, float re, float im)
{
    if (c->amp != 1) {
        re = re * c->amp;
        im = im * c->amp;
    }
    if (c->prescale != 1) {
        re = re * c->prescale;
        im = im * c->prescale;
    }
    float min_dist = FLT_MAX, cr = 0, ci = 0;
    for (int i = 0; i < c->states; ++i) {
        float dr = re - c->re[i], di = im - c->im[i];
        float dist = sqrtf((dr * dr) + (di * di));
        if (dist < min_dist) {
            min_dist = dist;
            cr = c->re[i];
            ci = c->im[i];
        }
    }
    /* (sample * closest.conj()).phase(), complex_t::operator* of oracle/shim/dsp/types.h */
    float bi = -ci;
    float pr = (re * cr) - (im * bi), pi = (im * cr) + (re * bi);
    return atan2f(pi, pr);
}
/* demod_soft_lut(sample, nullptr, &phase_error) (constellation.cpp:293-322) */
float orc_demod_phase_error(const orc_constellation* c, float re, float im)
{
    if (c->bits == 5) return orc_demod_phase_error_calc(c, re, im);
    int x = (int)((re / 1.5) * 256 + 256 / 2);
    if (x < 0) x = 0;
    if (x >= 256) x = 255;
    int y = (int)((im / 1.5) * 256 + 256 / 2);
    if (y < 0) y = 0;
    if (y >= 256) y = 255;
    return c->perr[(size_t)x * 256 + y];
}
const float* orc_const_phase_lut(const orc_constellation* c) { return c->perr; }

orc_constellation* orc_const_create(int type, float g1, float g2)
{
    orc_constellation* c = (orc_constellation*)calloc(1, sizeof(*c));
    c->type = type;
    c->amp = 1.0f;
    c->sca = 50.0f;
    c->prescale = 1.0f;
    if (type == 1) { /* QPSK */
        c->states = 4;
        c->bits = 2;
        c->amp = 3;
        static const float sg[4][2] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
        for (int i = 0; i < 4; ++i) {
            c->re[i] = (float)(sg[i][0] * SQRT2_D);
            c->im[i] = (float)(sg[i][1] * SQRT2_D);
        }
    } else if (type == 3) { /* 8PSK */
        c->states = 8;
        c->bits = 3;
        const float h = 0.70710678118654752440;
        const float pr[8] = {0.0f, -h, h, 0.0f, -h, -1.0f, 1.0f, h};
        const float pi[8] = {-1.0f, h, -h, 1.0f, -h, 0.0f, 0.0f, h};
        memcpy(c->re, pr, sizeof(pr));
        memcpy(c->im, pi, sizeof(pi));
    } else if (type == 4) { /* 16APSK */
        c->states = 16;
        c->bits = 4;
        c->amp = 100;
        c->sca = 1;
        c->prescale = 0.53;
        float gamma1 = g1;
        if (!gamma1)
            gamma1 = 2.57;
        float r1 = sqrtf(4 / (1 + 3 * gamma1 * gamma1));
        float r2 = gamma1 * r1;
        r1 *= 0.5;
        r2 *= 0.5;
        static const int o_idx[12] = {15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4};
        static const float o_k[12] = {1.5f, 10.5f, 4.5f, 7.5f, 0.5f, 11.5f, 5.5f, 6.5f, 2.5f, 9.5f, 3.5f, 8.5f};
        static const int i_idx[4] = {3, 2, 1, 0};
        static const float i_k[4] = {0.5f, 3.5f, 1.5f, 2.5f};
        set_ring(c, o_idx, o_k, 12, r2, 12);
        set_ring(c, i_idx, i_k, 4, r1, 4);
    } else if (type == 5) { /* 32APSK */
        c->states = 32;
        c->bits = 5;
        c->amp = 100;
        c->sca = 1;
        c->prescale = 0.54;
        float gamma1 = g1, gamma2 = g2;
        if (!gamma1)
            gamma1 = 2.53;
        if (!gamma2)
            gamma2 = 4.30;
        float r1 = sqrtf(8 / (1 + 3 * gamma1 * gamma1 + 4 * gamma2 * gamma2));
        float r2 = gamma1 * r1;
        float r3 = gamma2 * r1;
        r1 *= 0.5;
        r2 *= 0.5;
        r3 *= 0.5;
        static const int m_idx[12] = {31, 30, 29, 28, 27, 26, 25, 24, 15, 13, 11, 9};
        static const float m_k[12] = {1.5f, 2.5f, 10.5f, 9.5f, 4.5f, 3.5f, 7.5f, 8.5f, 0.5f, 11.5f, 5.5f, 6.5f};
        static const int o_idx[16] = {23, 22, 21, 20, 19, 18, 17, 16, 7, 6, 5, 4, 3, 2, 1, 0};
        static const float o_k[16] = {1, 3, 14, 12, 6, 4, 9, 11, 0, 2, 15, 13, 7, 5, 8, 10};
        static const int i_idx[4] = {14, 12, 10, 8};
        static const float i_k[4] = {0.5f, 3.5f, 1.5f, 2.5f};
        set_ring(c, m_idx, m_k, 12, r2, 12);
        set_ring(c, o_idx, o_k, 16, r3, 16);
        set_ring(c, i_idx, i_k, 4, r1, 4);
    } else {
        free(c);
        return NULL;
    }
    /* make_lut(256): sample grid (x - 128)/256 * 1.5 on both axes, x outer */
    c->lut = (int8_t*)malloc((size_t)256 * 256 * c->bits);
    c->perr = (float*)malloc((size_t)256 * 256 * sizeof(float));
    for (int x = 0; x < 256; ++x)
        for (int y = 0; y < 256; ++y) {
            float xv = ((float)(x - 256 / 2) / (float)256) * 1.5f;
            float yv = ((float)(y - 256 / 2) / (float)256) * 1.5f;
            orc_demod_soft_calc(c, xv, yv, &c->lut[((size_t)x * 256 + y) * c->bits]);
            c->perr[(size_t)x * 256 + y] = orc_demod_phase_error_calc(c, xv, yv);
        }
    return c;
}

void orc_const_destroy(orc_constellation* c)
{
    if (c) {
        free(c->lut);
        free(c->perr);
        free(c);
    }
}
int orc_const_bits(const orc_constellation* c) { return c->bits; }
const int8_t* orc_const_lut(const orc_constellation* c) { return c->bits == 5 ? NULL : c->lut; }
void orc_const_points(const orc_constellation* c, float* re_im)
{
    for (int i = 0; i < c->states; ++i) {
        re_im[2 * i] = c->re[i];
        re_im[2 * i + 1] = c->im[i];
    }
}

void orc_demod_soft_lut(const orc_constellation* c, float re, float im, int8_t* bits)
{
    if (c->bits == 5) {
        orc_demod_soft_calc(c, re, im, bits);
        return;
    }
    int x = (int)((re / 1.5) * 256 + 256 / 2); /* double arithmetic, truncation */
    if (x < 0) x = 0;
    if (x >= 256) x = 255;
    int y = (int)((im / 1.5) * 256 + 256 / 2);
    if (y < 0) y = 0;
    if (y >= 256) y = 255;
    memcpy(bits, &c->lut[((size_t)x * 256 + y) * c->bits], (size_t)c->bits);
}

void orc_mod(const orc_constellation* c, int symbol, float* re_im)
{
    re_im[0] = (c->re[symbol] / c->amp) / c->prescale;
    re_im[1] = (c->im[symbol] / c->amp) / c->prescale;
}

void orc_deinterleave(int constellation, int shortframe, int rate, const int8_t* in, int8_t* out)
{
    int n = shortframe ? 16200 : 64800;
    if (constellation == 0) {
        for (int i = 0; i < n / 2; ++i) {
            out[2 * i + 1] = in[2 * i];
            out[2 * i] = in[2 * i + 1];
        }
        return;
    }
    int bits = constellation + 2, rows = n / bits;
    for (int j = 0; j < rows; ++j)
        for (int k = 0; k < bits; ++k) {
            int col = k;
            if (constellation == 1 && rate == 4) /* 8PSK 3/5: columns reversed */
                col = 2 - k;
            out[col * rows + j] = in[j * bits + k];
        }
}

int orc_bb_to_soft(const orc_constellation* c, int constellation, int shortframe, int rate,
                   const float* plframe, int8_t* out)
{
    int n = shortframe ? 16200 : 64800;
    int nsym = n / c->bits;
    int8_t* soft = (int8_t*)malloc((size_t)n);
    for (int i = 0; i < nsym; ++i)
        orc_demod_soft_lut(c, plframe[2 * (90 + i)], plframe[2 * (90 + i) + 1], &soft[i * c->bits]);
    orc_deinterleave(constellation, shortframe, rate, soft, out);
    free(soft);
    return n;
}

/* ---- PL descrambling (row 8(f)-2, the part without a loop state) ---------------------------------
 * S2Scrambling (dvbs2/codings/s2_scrambling.cpp:9-28, s2_scrambling.h:16-25): Gold sequence n from two 18-bit
 * LFSRs, x (taps 0,7; state 1 advanced n times) and y (taps 0,5,7,10; all ones); Rn[i] = z(i) + 2 z(i+131072),
 * z = lsb(x) ^ lsb(y).  descramble (s2_scrambling.cpp:37-58) turns a symbol by -Rn * 90 degrees; S2PLLBlock
 * restarts the sequence at every frame and advances it on every symbol after the 90 header symbols, pilots
 * included (dvbs2_pll.cpp:37-44). */
static uint32_t pl_step_x(uint32_t x) { return ((((x >> 7) ^ x) & 1u) << 18 | x) >> 1; }
static uint32_t pl_step_y(uint32_t y) { return ((((y >> 10) ^ (y >> 7) ^ (y >> 5) ^ y) & 1u) << 18 | y) >> 1; }
void orc_pl_rn(int codenum, uint8_t* rn /* 131072 */)
{
    uint32_t x = 1, y = 0x3FFFF;
    for (int i = 0; i < codenum; ++i) x = pl_step_x(x);
    for (int half = 0; half < 2; ++half)
        for (int i = 0; i < 131072; ++i) {
            unsigned z = (x ^ y) & 1u;
            rn[i] = half ? (uint8_t)(rn[i] | (z << 1)) : (uint8_t)z;
            x = pl_step_x(x);
            y = pl_step_y(y);
        }
}
void orc_pl_descramble(const uint8_t* rn, const float* in, int nsym, float* out)
{
    for (int i = 0; i < nsym; ++i) {
        float re = in[2 * i], im = in[2 * i + 1];
        switch (rn[i]) {
        case 3: out[2 * i] = -im; out[2 * i + 1] = re; break;
        case 2: out[2 * i] = -re; out[2 * i + 1] = -im; break;
        case 1: out[2 * i] = im; out[2 * i + 1] = -re; break;
        default: out[2 * i] = re; out[2 * i + 1] = im; break;
        }
    }
}
void orc_pl_scramble(const uint8_t* rn, const float* in, int nsym, float* out) /* s2_scrambling.cpp:60-81 */
{
    for (int i = 0; i < nsym; ++i) {
        float re = in[2 * i], im = in[2 * i + 1];
        switch (rn[i]) {
        case 3: out[2 * i] = im; out[2 * i + 1] = -re; break;
        case 2: out[2 * i] = -re; out[2 * i + 1] = -im; break;
        case 1: out[2 * i] = -im; out[2 * i + 1] = re; break;
        default: out[2 * i] = re; out[2 * i + 1] = im; break;
        }
    }
}
