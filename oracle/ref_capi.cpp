// TEST INFRASTRUCTURE -- C API around the UNMODIFIED reference sources under /root/reference.
// Built by oracle/Makefile into oracle/_ref/libdvbs2_ref.so (git-ignored, travels with gpurun).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it.  Nothing here is product code; no reference source is copied, only #included from
// where it lies.
//
// Call discipline: one BBFrameLDPC::decode per frame (lane-0 semantics, note N1),
// decoder objects are created once per (framesize, rate) and never destroyed (note N5: GF tables
// hang off static pointers nulled by the destructor).
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "dvbs2/codings/bbframe_ldpc.h"
#include "dvbs2/codings/bbframe_bch.h"
#include "dvbs2/codings/bbframe_descramble.h"
#include "dvbs2/codings/s2_deinterleaver.h"
#include "dvbs2/codings/xdsopl-ldpc-pabr/dvb_s2_tables.hh"
#include "common/dsp/demod/constellation.h"
#include "dvbs2/bbframe_ts_parser.h"
#include "dvbs2/codings/s2_scrambling.h"
// row 8(f)-3: PL sync, PLHEADER demodulation, coarse frequency error -- compiled against shim/dsp/processor.h etc.
#define private public   // the phase loop of S2PLHDRDemod is a private member; the tests compare its state
#include "dvbs2/dvbs2_pl_sync.h"
#include "dvbs2/dvbs2_plhdr_demod.h"
#include "dvbs2/dvbs2_pll.h"          // row 8(f)-2: the payload phase loop (its pcl is private too)
#undef private
#include "dvbs2/dvbs2_fed.h"
// row 8(f)-4, byte-domain half: the DVB-S outer decoder (header-only classes + the vendored libcorrect)
#include "dvbs/dvbs_interleaving.h"
#include "dvbs/dvbs_reedsolomon.h"
#include "dvbs/dvbs_scrambling.h"
// row 8(f)-4, inner half: soft symbols, the self-locking punctured Viterbi decoder
#include <sstream>
#include <string>
#include <functional>
#include "dvbs/dvbs_syms_to_soft.h"
#define private public   // Viterbi_DVBS keeps its BER test buffers as uninitialised members that it reads before writing
#include "dvbs/viterbi_all.h"
#undef private
#pragma push_macro("TS_SIZE")
#undef TS_SIZE             // (bbframe_ts_parser.h defines a macro of the name of a member of the deframer)
#define private public   // the deframer's bit window is allocated uninitialised; the harness zeroes it
#include "dvbs/dvbs_ts_deframer.h"
#undef private
#pragma pop_macro("TS_SIZE")

using namespace dsp::dvbs2;

namespace {
struct RefSet {
    BBFrameLDPC* ldpc;
    BBFrameBCH* bch;
    BBFrameDescrambler* descr;
};
std::mutex g_mu;
std::map<int, RefSet> g_sets;

RefSet& get_set(int shortframe, int rate) {
    std::lock_guard<std::mutex> lk(g_mu);
    int key = shortframe * 100 + rate;
    auto it = g_sets.find(key);
    if (it != g_sets.end()) return it->second;
    dvbs2_framesize_t fs = shortframe ? FECFRAME_SHORT : FECFRAME_NORMAL;
    RefSet s;
    s.ldpc = new BBFrameLDPC(fs, (dvbs2_code_rate_t)rate);
    s.bch = new BBFrameBCH(fs, (dvbs2_code_rate_t)rate);
    s.descr = new BBFrameDescrambler(fs, (dvbs2_code_rate_t)rate);
    return g_sets.emplace(key, s).first->second;
}

// "Fair SIMD" use of the vendored library: one frame per SIMD lane, blocks = lanes, the way
// upstream xdsopl/LDPC drives it (SURVEY.md 8d CPU baseline (b)).
struct Simd16 {
    LDPCInterface* ldpc;
    LDPCDecoder<simd_type, algorithm_type> dec;
    simd_type* buf;
    int N, K;
};
std::map<int, Simd16*> g_simd;

LDPCInterface* make_table(int shortframe, int rate) {
    if (!shortframe) {
        switch (rate) {
        case C1_4: return new LDPC<DVB_S2_TABLE_B1>();
        case C1_3: return new LDPC<DVB_S2_TABLE_B2>();
        case C2_5: return new LDPC<DVB_S2_TABLE_B3>();
        case C1_2: return new LDPC<DVB_S2_TABLE_B4>();
        case C3_5: return new LDPC<DVB_S2_TABLE_B5>();
        case C2_3: return new LDPC<DVB_S2_TABLE_B6>();
        case C3_4: return new LDPC<DVB_S2_TABLE_B7>();
        case C4_5: return new LDPC<DVB_S2_TABLE_B8>();
        case C5_6: return new LDPC<DVB_S2_TABLE_B9>();
        case C8_9: return new LDPC<DVB_S2_TABLE_B10>();
        case C9_10: return new LDPC<DVB_S2_TABLE_B11>();
        default: return nullptr;
        }
    }
    switch (rate) {
    case C1_4: return new LDPC<DVB_S2_TABLE_C1>();
    case C1_3: return new LDPC<DVB_S2_TABLE_C2>();
    case C2_5: return new LDPC<DVB_S2_TABLE_C3>();
    case C1_2: return new LDPC<DVB_S2_TABLE_C4>();
    case C3_5: return new LDPC<DVB_S2_TABLE_C5>();
    case C2_3: return new LDPC<DVB_S2_TABLE_C6>();
    case C3_4: return new LDPC<DVB_S2_TABLE_C7>();
    case C4_5: return new LDPC<DVB_S2_TABLE_C8>();
    case C5_6: return new LDPC<DVB_S2_TABLE_C9>();
    case C8_9: return new LDPC<DVB_S2_TABLE_C10>();
    default: return nullptr;
    }
}
} // namespace

extern "C" {

int ref_simd_lanes() { return simd_type::SIZE; }

// BBFrameLDPC::decode (bbframe_ldpc.cpp:123-139): N int8 LLRs in place; returns iterations or -1.
int ref_ldpc_decode(int shortframe, int rate, int8_t* frame, int max_trials) {
    return get_set(shortframe, rate).ldpc->decode(frame, max_trials);
}
int ref_ldpc_data_size(int shortframe, int rate) { return get_set(shortframe, rate).ldpc->dataSize(); }

// LDPCDecoder::operator()(..., blocks = lanes) on `lanes` frames at once (frames: lanes x N int8,
// frame-major).  iters_out[l] is what a per-lane caller would see only when all lanes stop together;
// the SIMD decoder stops when every lane is clean, so a single return value applies to the group.
int ref_ldpc_decode_simd(int shortframe, int rate, int8_t* frames, int max_trials) {
    Simd16* s;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int key = shortframe * 100 + rate;
        auto it = g_simd.find(key);
        if (it == g_simd.end()) {
            s = new Simd16();
            s->ldpc = make_table(shortframe, rate);
            if (!s->ldpc) return -2;
            s->dec.init(s->ldpc);
            s->N = s->ldpc->code_len();
            s->K = s->ldpc->data_len();
            s->buf = new simd_type[s->N];
            g_simd[key] = s;
        } else {
            s = it->second;
        }
    }
    const int L = simd_type::SIZE;
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < s->N; ++i)
            reinterpret_cast<code_type*>(s->buf + i)[l] = frames[(size_t)l * s->N + i];
    int trials = s->dec(s->buf, s->buf + s->K, max_trials, L);
    for (int l = 0; l < L; ++l)
        for (int i = 0; i < s->N; ++i)
            frames[(size_t)l * s->N + i] = reinterpret_cast<code_type*>(s->buf + i)[l];
    return trials < 0 ? trials : max_trials - trials;
}

// LDPCEncoder<int8_t> (encoder.hh:37-52) with the decoder's own convention: bit 1 <-> negative.
// bits: K 0/1 bytes in, N 0/1 bytes out (systematic).
int ref_ldpc_encode_bits(int shortframe, int rate, const uint8_t* data_bits, uint8_t* code_bits) {
    LDPCInterface* t = make_table(shortframe, rate);
    if (!t) return -1;
    LDPCEncoder<int8_t> enc;
    enc.init(t);
    int N = t->code_len(), K = t->data_len();
    std::vector<int8_t> soft(N);
    for (int i = 0; i < K; ++i) soft[i] = data_bits[i] ? -1 : 1;
    enc(soft.data(), soft.data() + K);
    for (int i = 0; i < N; ++i) code_bits[i] = soft[i] < 0;
    delete t;
    return 0;
}

// Dump of LDPCDecoder::init's expansion, re-derived through the public LDPC<TABLE> iterator
// (ldpc.hh:25-109): for data bit j, its check indices.  out must hold LINKS_TOTAL ints; returns
// number of (bit, check) pairs written as bit*65536... no: pairs are written as two arrays.
int ref_ldpc_links(int shortframe, int rate, int* bit_of_link, int* check_of_link, int cap) {
    LDPCInterface* t = make_table(shortframe, rate);
    if (!t) return -1;
    int K = t->data_len(), n = 0;
    t->first_bit();
    for (int j = 0; j < K; ++j) {
        int* acc = t->acc_pos();
        int deg = t->bit_deg();
        for (int d = 0; d < deg; ++d) {
            if (n < cap) { bit_of_link[n] = j; check_of_link[n] = acc[d]; }
            ++n;
        }
        t->next_bit();
    }
    delete t;
    return n;
}

// BBFrameBCH (bbframe_bch.cpp:380-456): nbch/8 bytes in place.
int ref_bch_decode(int shortframe, int rate, uint8_t* frame) { return get_set(shortframe, rate).bch->decode(frame); }
int ref_bch_encode(int shortframe, int rate, uint8_t* frame) { return get_set(shortframe, rate).bch->encode(frame); }
int ref_bch_data_size(int shortframe, int rate) { return get_set(shortframe, rate).bch->dataSize(); }

// BBFrameDescrambler::work (bbframe_descramble.cpp:138-143): kbch/8 bytes in place.
int ref_descramble(int shortframe, int rate, uint8_t* frame) { return get_set(shortframe, rate).descr->work(frame); }

// S2Deinterleaver (s2_deinterleaver.cpp:72-202).
void ref_deinterleave(int constellation, int shortframe, int rate, int8_t* in, int8_t* out) {
    S2Deinterleaver d((dvbs2_constellation_t)constellation, shortframe ? FECFRAME_SHORT : FECFRAME_NORMAL,
                      (dvbs2_code_rate_t)rate);
    d.deinterleave(in, out);
}
void ref_interleave(int constellation, int shortframe, int rate, uint8_t* in, uint8_t* out) {
    S2Deinterleaver d((dvbs2_constellation_t)constellation, shortframe ? FECFRAME_SHORT : FECFRAME_NORMAL,
                      (dvbs2_code_rate_t)rate);
    d.interleave(in, out);
}

// constellation_t (constellation.cpp): type is dsp::constellation_type_t (QPSK=1, PSK8=3, APSK16=4,
// APSK32=5).  Objects are cached per (type, g1, g2) with make_lut(256) done once, as
// DVBS2Demod::init does (module_dvbs2_demod.cpp:70-71).
struct RefConst {
    dsp::constellation_t* c;
};
static std::map<std::vector<float>, dsp::constellation_t*> g_const;
static dsp::constellation_t* get_const(int type, float g1, float g2) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<float> key{(float)type, g1, g2};
    auto it = g_const.find(key);
    if (it != g_const.end()) return it->second;
    auto* c = new dsp::constellation_t((dsp::constellation_type_t)type, g1, g2);
    c->make_lut(256);
    g_const[key] = c;
    return c;
}
// S2BBToSoft::process' demap loop (dvbs2_bb_to_soft.cpp:11-16), pilots off: nsym symbols starting
// at `sym` (i.e. the caller passes &plframe[90]); out gets nsym*bits int8 in symbol order.
int ref_demap(int type, float g1, float g2, const float* sym, int nsym, int8_t* out) {
    dsp::constellation_t* c = get_const(type, g1, g2);
    int bits = c->getBitsCnt();
    for (int i = 0; i < nsym; ++i)
        c->demod_soft_lut(dsp::complex_t{sym[2 * i], sym[2 * i + 1]}, &out[i * bits]);
    return bits;
}
// constellation_t::demod_soft_calc direct (used by the 32APSK path and by make_lut).
void ref_demap_calc(int type, float g1, float g2, const float* sym, int nsym, int8_t* out) {
    dsp::constellation_t* c = get_const(type, g1, g2);
    int bits = c->getBitsCnt();
    for (int i = 0; i < nsym; ++i)
        c->demod_soft_calc(dsp::complex_t{sym[2 * i], sym[2 * i + 1]}, &out[i * bits]);
}
// constellation_t::mod (constellation.cpp:156-158).
void ref_mod(int type, float g1, float g2, const uint8_t* symbols, int nsym, float* out) {
    dsp::constellation_t* c = get_const(type, g1, g2);
    for (int i = 0; i < nsym; ++i) {
        dsp::complex_t v = c->mod(symbols[i]);
        out[2 * i] = v.re;
        out[2 * i + 1] = v.im;
    }
}

// ---- BBFrameTSParser (dvbs2/bbframe_ts_parser.h:67-108), one long-lived object per handle ----
void* ref_ts_create(int kbch_bits) {
    auto* p = new dsp::dvbs2::BBFrameTSParser();
    p->setFrameSize(kbch_bits);
    return p;
}
void ref_ts_destroy(void* h) { delete static_cast<dsp::dvbs2::BBFrameTSParser*>(h); }
int ref_ts_work(void* h, uint8_t* bbframes, int cnt, uint8_t* out, int out_cap) {
    return static_cast<dsp::dvbs2::BBFrameTSParser*>(h)->work(bbframes, cnt, out, out_cap);
}
// public members after work(): fields[0..10] = ts_gs, sis_mis, ccm_acm, issyi, npd, ro, isi, upl, dfl, sync, syncd
void ref_ts_stats(void* h, int* fields, int* last_bb_cnt, int* last_bb_proc, int* last_gse_crc_err) {
    auto* p = static_cast<dsp::dvbs2::BBFrameTSParser*>(h);
    const dsp::dvbs2::BBHeader& b = p->last_header;
    int v[11] = {b.ts_gs, b.sis_mis, b.ccm_acm, b.issyi, b.npd, b.ro, b.isi, b.upl, b.dfl, b.sync, b.syncd};
    for (int i = 0; i < 11; ++i) fields[i] = v[i];
    *last_bb_cnt = p->last_bb_cnt;
    *last_bb_proc = p->last_bb_proc;
    *last_gse_crc_err = p->last_gse_crc_err;
}

// ---- S2Scrambling (dvbs2/codings/s2_scrambling.h:12-45): reset(), then one (de)scramble per symbol ----
void ref_pl_descramble(int codenum, const float* in, int nsym, float* out, int scramble) {
    static std::map<int, std::unique_ptr<dsp::dvbs2::S2Scrambling>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto& sc = cache[codenum];
    if (!sc) sc.reset(new dsp::dvbs2::S2Scrambling(codenum));
    sc->reset();
    for (int i = 0; i < nsym; ++i) {
        dsp::complex_t v{in[2 * i], in[2 * i + 1]};
        dsp::complex_t r = scramble ? sc->scramble(v) : sc->descramble(v);
        out[2 * i] = r.re;
        out[2 * i + 1] = r.im;
    }
}

// ---- S2PLSyncBlock (dvbs2/dvbs2_pl_sync.h:11-66): init(nullptr, slot_num, pilots, sof, pls), then process() ----
namespace {
s2_sof g_sof;
s2_plscodes g_pls;
}
void* ref_plsync_create(int slot_num, int pilots) {
    auto* b = new S2PLSyncBlock();
    b->init(nullptr, slot_num, pilots != 0, &g_sof, &g_pls);
    return b;
}
int ref_plsync_process(void* h, int count, const float* in, float* out) {
    return static_cast<S2PLSyncBlock*>(h)->process(count, (dsp::complex_t*)in, (dsp::complex_t*)out);
}
void ref_plsync_stats(void* h, int* raw_frame_size, int* current_position, double* best_match) {
    auto* b = static_cast<S2PLSyncBlock*>(h);
    *raw_frame_size = b->raw_frame_size;
    *current_position = b->current_position;
    *best_match = b->best_match;
}
// ---- S2PLHDRDemod (dvbs2/dvbs2_plhdr_demod.h:13-49) ----
void* ref_plhdr_create(float loop_bw) {
    auto* b = new S2PLHDRDemod();
    b->init(nullptr, loop_bw, &g_sof, &g_pls);
    return b;
}
// count symbols of one frame in, the 90 header symbols out; res = {modcod, shortframes, pilots}; loop = {phase, freq}
int ref_plhdr_process(void* h, int count, const float* in, float* out90, int* res, float* loop) {
    auto* b = static_cast<S2PLHDRDemod*>(h);
    std::vector<dsp::complex_t> out(count > 90 ? count : 90);
    int r = b->process(count, (dsp::complex_t*)in, out.data());
    memcpy(out90, out.data(), 90 * sizeof(dsp::complex_t));
    res[0] = b->detect_modcod;
    res[1] = b->detect_shortframes;
    res[2] = b->detect_pilots;
    loop[0] = b->pcl.phase;
    loop[1] = b->pcl.freq;
    return r;
}
// ---- dvbs2_pilot_coarse_fed (dvbs2/dvbs2_fed.h:7-48) ----
float ref_coarse_fed(const float* frame, int raw_frame_size, int pilots, int pls_code, int codenum) {
    static std::map<int, std::unique_ptr<dsp::dvbs2::S2Scrambling>> cache;
    auto& sc = cache[codenum];
    if (!sc) sc.reset(new dsp::dvbs2::S2Scrambling(codenum));
    return dvbs2_pilot_coarse_fed((dsp::complex_t*)frame, raw_frame_size, pilots != 0, pls_code, g_sof, g_pls, sc.get());
}
// symbols of the PLHEADER of PLS index `pls_code` (s2_defs.h:16-31,33-86): 26 SOF + 64 code symbols
void ref_plheader_symbols(int pls_code, float* out90) {
    for (int i = 0; i < 26; ++i) { out90[2 * i] = g_sof.symbols[i].re; out90[2 * i + 1] = g_sof.symbols[i].im; }
    for (int i = 0; i < 64; ++i) { out90[52 + 2 * i] = g_pls.symbols[pls_code][i].re; out90[52 + 2 * i + 1] = g_pls.symbols[pls_code][i].im; }
}
unsigned long long ref_pls_codeword(int pls_code) { return g_pls.codewords[pls_code]; }

// ---- S2PLLBlock (dvbs2/dvbs2_pll.h:17-78, dvbs2_pll.cpp:5-86) as DVBS2Demod::init sets it up
//      (module_dvbs2_demod.cpp:59-65): init, pilots, constellation + make_lut(256), frame_slot_count, pls_code, update ----
struct RefPll {
    S2PLLBlock blk;
    std::unique_ptr<dsp::dvbs2::S2Scrambling> sc;
};
void* ref_pll_create(float loop_bw, int const_type, float g1, float g2, int frame_slot_count, int pilots, int pls_code, int codenum) {
    auto* r = new RefPll();
    r->sc.reset(new dsp::dvbs2::S2Scrambling(codenum));
    r->blk.init(nullptr, loop_bw, &g_sof, &g_pls, r->sc.get());
    r->blk.pilots = pilots != 0;
    r->blk.constellation = std::make_shared<dsp::constellation_t>((dsp::constellation_type_t)const_type, g1, g2);
    r->blk.constellation->make_lut(256);
    r->blk.frame_slot_count = frame_slot_count;
    r->blk.pls_code = pls_code;
    r->blk.update();
    return r;
}
// one frame in (at least (frame_slot_count + 1) * 90 + pilot_cnt * 36 symbols), as many symbols out; returns that number.
// state = {pcl.phase, pcl.freq, error (the block's public average)}
int ref_pll_process(void* h, int count, const float* in, float* out, float* state) {
    auto* r = static_cast<RefPll*>(h);
    int n = r->blk.process(count, (dsp::complex_t*)in, (dsp::complex_t*)out);
    state[0] = r->blk.pcl.phase;
    state[1] = r->blk.pcl.freq;
    state[2] = r->blk.error;
    return n;
}
int ref_pll_pilot_cnt(void* h) { return static_cast<RefPll*>(h)->blk.pilot_cnt; }

// ---- DVB-S outer decoder: the objects and the frame loop body of DVBSDemod::process (dvbs/module_dvbs_demod.cpp:91-106) ----
struct RefDvbsOuter {
    dsp::dvbs::DVBSInterleaving dvb_interleaving;
    dsp::dvbs::DVBSReedSolomon reed_solomon;
    dsp::dvbs::DVBSScrambling scrambler;
    uint8_t tmp_deinterleaved_frame[204 * 8];
};
void* ref_dvbs_outer_create() { return new RefDvbsOuter(); }
void ref_dvbs_outer_process(void* h, const uint8_t* frames, int nframes, int stride, uint8_t* out, int* errors) {
    auto* r = static_cast<RefDvbsOuter*>(h);
    int outidx = 0;
    for (int k = 0; k < nframes; k++) {
        uint8_t* current_frame = const_cast<uint8_t*>(&frames[(size_t)k * stride]);
        r->dvb_interleaving.deinterleave(current_frame, r->tmp_deinterleaved_frame);
        for (int i = 0; i < 8; i++) errors[8 * k + i] = r->reed_solomon.decode(&r->tmp_deinterleaved_frame[204 * i]);
        r->scrambler.descramble(r->tmp_deinterleaved_frame);
        for (int i = 0; i < 8; i++) {
            memcpy(&out[outidx], &r->tmp_deinterleaved_frame[204 * i], 188);
            outidx += 188;
        }
    }
}
// ---- DVBS_TS_Deframer (dvbs/dvbs_ts_deframer.h:17-70) ----
void* ref_dvbs_deframer_create() {
    auto* d = new deframing::DVBS_TS_Deframer();
    memset(d->full_frame_shifter, 0, 1632 * 8);   // (new uint8_t[TS_SIZE] in the reference: indeterminate)
    return d;
}
int ref_dvbs_deframer_work(void* h, const uint8_t* input, int size, uint8_t* output) {
    return static_cast<deframing::DVBS_TS_Deframer*>(h)->work(const_cast<uint8_t*>(input), size, output);
}
void ref_dvbs_deframer_stats(void* h, int* errors_nor, int* errors_inv) {
    auto* d = static_cast<deframing::DVBS_TS_Deframer*>(h);
    *errors_nor = d->errors_nor;
    *errors_inv = d->errors_inv;
}
// ---- viterbi::Viterbi_DVBS (dvbs/viterbi_all.h:33-163) as the module constructs it (module_dvbs_demod.cpp:23: BER threshold
// 0.15, 20 calls out of sync, VIT_BUF_SIZE = 8192 soft bits per call, phases 0 and 90 degrees) ----
void* ref_vit_create(float ber_threshold, int max_outsync) {
    auto* v = new viterbi::Viterbi_DVBS(ber_threshold, max_outsync, 8192, {PHASE_0, PHASE_90});
    // members the search reads before anything has written them (viterbi_all.cpp:89,107,127,147,167: the decoders read
    // past what the depuncturers wrote; get_ber reads one re-encoded bit rate 5/6 never writes): zeros here and in the oracle
    memset(v->ber_test_buffer, 0, sizeof v->ber_test_buffer);
    memset(v->ber_soft_buffer, 0, sizeof v->ber_soft_buffer);
    memset(v->ber_depunc_buffer, 0, sizeof v->ber_depunc_buffer);
    memset(v->ber_decoded_buffer, 0, sizeof v->ber_decoded_buffer);
    memset(v->ber_encoded_buffer, 0, sizeof v->ber_encoded_buffer);
    v->d_rate = viterbi::RATE_1_2; v->d_phase = PHASE_0; v->d_shift = 0; v->d_ber = 10;
    return v;
}
void ref_vit_destroy(void* h) { delete static_cast<viterbi::Viterbi_DVBS*>(h); }
// layout the restatement relies on: the 1/2 search decoder reads up to 13 bytes past ber_soft_buffer, into the member behind it
int ref_vit_layout_ok(void* h) {
    auto* v = static_cast<viterbi::Viterbi_DVBS*>(h);
    return v->ber_soft_buffer + sizeof v->ber_soft_buffer == v->ber_depunc_buffer;
}
// DVBSVitBlock::process (dvbs/dvbs_vit.cpp:6-13): one work() per 8192 soft bits; input is rotated in place as the reference does
int ref_vit_process(void* h, int count, int8_t* in, uint8_t* out) {
    auto* v = static_cast<viterbi::Viterbi_DVBS*>(h);
    int oidx = 0;
    for (int i = 0; i < count; i += 8192) oidx += v->work(&in[i], 8192, &out[oidx]);
    return oidx;
}
void ref_vit_stats(void* h, float* ber, int* state, int* rate, int* phase, int* shift, int* invalid) {
    auto* v = static_cast<viterbi::Viterbi_DVBS*>(h);
    *ber = v->ber(); *state = v->getState(); *rate = (int)v->d_rate; *phase = (int)v->d_phase; *shift = v->d_shift; *invalid = v->d_invalid;
}
// ---- DVBSymToSoftBlock::process (dvbs/dvbs_syms_to_soft.cpp:26-42) ----
void* ref_sts_create() {
    auto* s = new dsp::dvbs::DVBSymToSoftBlock();
    s->init(nullptr, STREAM_BUFFER_SIZE);
    s->syms_callback = [](dsp::complex_t*, int) {};
    return s;
}
int ref_sts_process(void* h, int count, const float* syms, int8_t* out) {
    return static_cast<dsp::dvbs::DVBSymToSoftBlock*>(h)->process(count, (dsp::complex_t*)syms, out);
}
// correct_reed_solomon_encode through the same 255-byte layout the decoder wrapper uses: parity of one 188-byte packet
void ref_rs204_parity(const uint8_t* msg188, uint8_t* parity16) {
    static correct_reed_solomon* rs = correct_reed_solomon_create(correct_rs_primitive_polynomial_8_4_3_2_0, 0, 1, 16);
    uint8_t in[239] = {0}, enc[255];
    memcpy(in + 51, msg188, 188);
    correct_reed_solomon_encode(rs, in, 239, enc);
    memcpy(parity16, enc + 239, 16);
}

} // extern "C"
