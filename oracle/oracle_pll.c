/* TEST INFRASTRUCTURE -- see oracle.h.  Restatement of the reference's payload phase loop:
 *
 *   S2PLLBlock::init / update / process     dvbs2/dvbs2_pll.cpp:5-14,34-86; dvbs2/dvbs2_pll.h:47-59
 *   PhaseControlLoop<float>                  SDR++ core (stand-in: oracle/shim/dsp/loop/phase_control_loop.h, unpinned)
 *   math::phasor, complex_t                  SDR++ core (stand-ins: oracle/shim/dsp/math/phasor.h, dsp/types.h, unpinned)
 *   S2Scrambling::descramble                 dvbs2/codings/s2_scrambling.cpp:37-58 (sequence: orc_pl_rn)
 *   constellation_t::demod_soft_lut error    oracle_demap.c (orc_demod_phase_error)
 *   s2_sof / s2_plscodes symbols             oracle_plsync.c (orc_plheader_symbols)
 *
 * Float arithmetic follows the reference operation by operation (built with -ffp-contract=off); with the same libm the
 * output symbols and the loop state are bit-identical to the compiled reference (tests/test_pll_oracle.py). */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

#define FL_PI 3.1415926535f
typedef struct { float re, im; } cf;
static cf cmul(cf a, cf b) { cf r = {(a.re * b.re) - (a.im * b.im), (a.im * b.re) + (a.re * b.im)}; return r; }
static cf cconj(cf a) { cf r = {a.re, -a.im}; return r; }
static float cphase(cf a) { return atan2f(a.im, a.re); }

struct orc_pll {
    float alpha, beta, phase, freq, error;
    int slots, pilot_cnt, pls_code;
    orc_constellation* c;
    uint8_t* rn;
    cf hdr[90];
};

orc_pll* orc_pll_create(float loop_bw, int const_type, float g1, float g2, int frame_slot_count, int pilots, int pls_code, int codenum)
{
    orc_pll* p = (orc_pll*)calloc(1, sizeof *p);
    /* PhaseControlLoop<float>::criticallyDamped(loop_bw, alpha, beta) (dvbs2_pll.cpp:10) */
    const float bw = loop_bw;
    const float damp = (float)(sqrt(2.0) / 2.0);
    const float den = (float)(1.0 + 2.0 * damp * bw + bw * bw);
    p->alpha = (4 * damp * bw) / den;
    p->beta = (4 * bw * bw) / den;
    p->slots = frame_slot_count;
    p->pls_code = pls_code & 127;
    /* update() (dvbs2_pll.h:47-59): counts pilot blocks from the SLOT number, not the symbol number (SURVEY note N2) */
    p->pilot_cnt = 0;
    if (pilots) {
        int raw_size = (frame_slot_count - 90) / 90;
        p->pilot_cnt = 1;
        raw_size -= 16;
        while (raw_size > 16) {
            raw_size -= 16;
            p->pilot_cnt++;
        }
    }
    p->c = orc_const_create(const_type, g1, g2);
    p->rn = (uint8_t*)malloc(131072);
    orc_pl_rn(codenum, p->rn);
    orc_plheader_symbols(p->pls_code, (float*)p->hdr);
    return p;
}
void orc_pll_destroy(orc_pll* p)
{
    if (!p) return;
    orc_const_destroy(p->c);
    free(p->rn);
    free(p);
}
int orc_pll_pilot_cnt(const orc_pll* p) { return p->pilot_cnt; }

static void advance(orc_pll* p, float error) /* pcl.init(alpha, beta, 0, -pi, pi, 0, -0.01 pi, 0.01 pi) (:11) */
{
    p->freq += p->beta * error;
    if (p->freq > 0.01f * FL_PI) p->freq = 0.01f * FL_PI;
    else if (p->freq < -0.01f * FL_PI) p->freq = -0.01f * FL_PI;
    p->phase += p->freq + (p->alpha * error);
    const float delta = FL_PI - (-FL_PI);
    while (p->phase > FL_PI) p->phase -= delta;
    while (p->phase < -FL_PI) p->phase += delta;
}

int orc_pll_process(orc_pll* p, const float* in_f, float* out_f, float* state) /* process (:34-86) */
{
    const cf* in = (const cf*)in_f;
    cf* out = (cf*)out_f;
    const int total = (p->slots + 1) * 90 + p->pilot_cnt * 36;
    float errorsum = 0;
    int pilotctr = 0, pos = 0;
    for (int i = 0; i < total; ++i) {
        const cf ph = {cosf(-p->phase), sinf(-p->phase)};
        const cf tmp = cmul(in[i], ph);
        float error = 0;
        if (i >= 90) {
            cf descr;
            switch (p->rn[pos++]) {
            case 3: descr.re = -tmp.im; descr.im = tmp.re; break;
            case 2: descr.re = -tmp.re; descr.im = -tmp.im; break;
            case 1: descr.re = tmp.im; descr.im = -tmp.re; break;
            default: descr = tmp; break;
            }
            if (p->pilot_cnt == 0) {
                error = orc_demod_phase_error(p->c, tmp.re, tmp.im);
                out[i] = descr;
            } else {
                if (pilotctr >= 0) {
                    error = orc_demod_phase_error(p->c, tmp.re, tmp.im);
                    out[i] = descr;
                    pilotctr++;
                    if (pilotctr >= 16 * 90) pilotctr = -1;
                }
                if (pilotctr < 0) {   /* (not "else": the 1440th data symbol is also the first "pilot") */
                    const cf ideal = {descr.re > 0 ? 0.707f : -0.707f, descr.im > 0 ? 0.707f : -0.707f};
                    error = cphase(cmul(descr, cconj(ideal))) / 10.0f;
                    out[i] = descr;
                    pilotctr--;
                    if (pilotctr <= -36) pilotctr = 0;
                }
            }
        } else {
            error = cphase(cmul(tmp, cconj(p->hdr[i])));
            if (i & 1) { out[i].re = -tmp.re; out[i].im = tmp.im; }
            else { out[i].re = tmp.im; out[i].im = tmp.re; }
        }
        errorsum += error;
        advance(p, error);
    }
    p->error = errorsum / ((float)(p->slots + 1) * 90 + p->pilot_cnt * 36);
    state[0] = p->phase;
    state[1] = p->freq;
    state[2] = p->error;
    return total;
}
