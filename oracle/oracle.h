/* TEST INFRASTRUCTURE -- CPU oracle for the DVB-S2 decode stage (plain C restatement).
 *
 * This directory restates, function by function, what the reference's hot path computes
 * (SURVEY.md section 8a); every function cites the reference file:line it follows.  It is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg ONLY.  The product
 * (sdrpp-dvbs-demodulator_b200/) never includes, links or calls anything from here.
 *
 * Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4).  The restatement is therefore pinned by differential runs against the reference's
 * own sources compiled unmodified into oracle/_ref/libdvbs2_ref.so (tests/test_oracle_vs_ref.py,
 * run wherever that library exists) and by the committed outputs of those runs under
 * tests/golden/ (made by tools/make_golden.py).  The demapper additionally depends on a
 * 20-line stand-in for SDR++ core's complex_t (oracle/shim): that part is "parity unpinned"
 * against the real SDR++ header.
 */
#ifndef DVBS2_ORACLE_H
#define DVBS2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* rate uses the reference's dvbs2_code_rate_t numbering (dvbs2/dvbs2.h:11-25):
 * 0=1/4 1=1/3 2=2/5 3=1/2 4=3/5 5=2/3 6=3/4 7=4/5 8=5/6 9=7/8(no table) 10=8/9 11=9/10 */
int orc_code_params(int shortframe, int rate, int* N, int* K, int* kbch, int* bch_t, int* q, int* links_total);

/* LDPCDecoder::operator() through BBFrameLDPC::decode (bbframe_ldpc.cpp:123-139,
 * layered_decoder.hh:121-133): llr[N] in place; returns iterations executed (0..max), or -1. */
int orc_ldpc_decode(int shortframe, int rate, int8_t* llr, int max_trials);
/* LDPCEncoder (encoder.hh:37-52), bit domain: K data bits in -> N code bits out. */
int orc_ldpc_encode_bits(int shortframe, int rate, const uint8_t* data_bits, uint8_t* code_bits);
/* layered schedule dump: pos[R*CNL] (layered_decoder.hh:95-119), cnc[q]; returns CNL */
int orc_ldpc_schedule(int shortframe, int rate, uint16_t* pos, uint8_t* cnc);

/* module_dvbs2_demod.cpp:357-360: bit i = (llr[i] < 0), MSB first, nbits bits */
void orc_repack(const int8_t* llr, int nbits, uint8_t* bytes);
/* BBFrameBCH::decode (bbframe_bch.cpp:380-405) -> BoseChaudhuriHocquenghemDecoder::operator() */
int orc_bch_decode(int shortframe, int rate, uint8_t* frame);
/* systematic BCH encode of kbch bits -> appends parity (same codeword as bbframe_bch.cpp:407-456) */
int orc_bch_encode(int shortframe, int rate, uint8_t* frame);
/* BBFrameDescrambler::work (bbframe_descramble.cpp:122-143) */
int orc_descramble(int shortframe, int rate, uint8_t* frame);
/* BBFrameTSParser::check_crc8(bbf, 80) (dvbs2/bbframe_ts_parser.cpp:66-80): 0 = BBHEADER valid */
unsigned orc_bbheader_crc8(const uint8_t* bbframe);
/* ---- upstream row 8(f)-2: S2Scrambling (dvbs2/codings/s2_scrambling.cpp), see oracle_demap.c ---- */
void orc_pl_rn(int codenum, uint8_t* rn /* 131072 */);
void orc_pl_descramble(const uint8_t* rn, const float* in, int nsym, float* out);
void orc_pl_scramble(const uint8_t* rn, const float* in, int nsym, float* out);

/* ---- downstream row 8(f)-1: BBFrameTSParser (dvbs2/bbframe_ts_parser.cpp:31-43,100-392), see oracle_ts.c ---- */
typedef struct orc_ts_parser orc_ts_parser;
orc_ts_parser* orc_ts_create(int kbch_bits);                     /* setFrameSize */
void orc_ts_destroy(orc_ts_parser* p);
/* work(): cnt BBFRAMEs of kbch/8 bytes -> TS packets (or GRE-wrapped GSE PDUs); returns bytes written */
int orc_ts_work(orc_ts_parser* p, const uint8_t* bbframes, int cnt, uint8_t* out, int out_cap);
/* last_header as the 10 raw BBHEADER bytes (have_header = 0 until one frame was accepted), last_bb_cnt,
 * last_bb_proc, last_gse_crc_err, plus the private sync flag and the size of the pending partial unit */
void orc_ts_stats(const orc_ts_parser* p, uint8_t last_header[10], int* have_header, int* last_bb_cnt, int* last_bb_proc,
                  int* last_gse_crc_err, int* synched, int* pending);
/* transmit-side helpers for tests (EN 302 307 5.1.4, 5.1.6) */
void orc_bbheader_seal(uint8_t* hdr10);
uint8_t orc_up_crc8(const uint8_t* payload187);
/* whole A3..A10 chain for one frame: llr[N] in (modified), bb[kbch/8] out */
int orc_decode_frame(int shortframe, int rate, int8_t* llr, int max_trials, uint8_t* bb, int* ldpc_iters, int* bch_corr);

/* constellation types follow dsp::constellation_type_t (constellation.h:9-16): 1=QPSK 3=8PSK 4=16APSK 5=32APSK */
typedef struct orc_constellation orc_constellation;
orc_constellation* orc_const_create(int type, float g1, float g2); /* ctor + make_lut(256) */
void orc_const_destroy(orc_constellation* c);
int orc_const_bits(const orc_constellation* c);
const int8_t* orc_const_lut(const orc_constellation* c); /* [256][256][bits] (x major), NULL for 32APSK */
void orc_const_points(const orc_constellation* c, float* re_im); /* constellation[] incl. const_amp */
void orc_demod_soft_calc(const orc_constellation* c, float re, float im, int8_t* bits);
void orc_demod_soft_lut(const orc_constellation* c, float re, float im, int8_t* bits);
void orc_mod(const orc_constellation* c, int symbol, float* re_im);
/* phase_error of demod_soft_calc / demod_soft_lut (constellation.cpp:258-260,317-318) and the 256x256 table */
float orc_demod_phase_error_calc(const orc_constellation* c, float re, float im);
float orc_demod_phase_error(const orc_constellation* c, float re, float im);
const float* orc_const_phase_lut(const orc_constellation* c);
/* S2Deinterleaver::deinterleave (s2_deinterleaver.cpp:72-136); constellation 0=QPSK 1=8PSK 2=16APSK 3=32APSK */
void orc_deinterleave(int constellation, int shortframe, int rate, const int8_t* in, int8_t* out);
/* S2BBToSoft::process (dvbs2_bb_to_soft.cpp:7-33), pilots off: plframe = 90 header symbols + payload */
int orc_bb_to_soft(const orc_constellation* c, int constellation, int shortframe, int rate,
                   const float* plframe, int8_t* out);

/* ---- row 8(f)-3 (oracle_plsync.c): PL sync, PLHEADER demodulation, coarse frequency error ---- */
/* complex symbols are interleaved float pairs (re, im) */
typedef struct orc_plsync orc_plsync;
int orc_raw_frame_size(int slot_num, int pilots);                                     /* dvbs2_pl_sync.cpp:15-30 */
orc_plsync* orc_plsync_create(int slot_num, int pilots);                              /* S2PLSyncBlock::init */
void orc_plsync_destroy(orc_plsync* p);
int orc_plsync_process(orc_plsync* p, int count, const float* in, float* out);        /* ::process (:80-100) */
void orc_plsync_stats(const orc_plsync* p, int* raw_frame_size, int* current_position, double* best_match);
typedef struct orc_plhdr orc_plhdr;
orc_plhdr* orc_plhdr_create(float loop_bw);                                           /* S2PLHDRDemod::init */
void orc_plhdr_destroy(orc_plhdr* p);
/* S2PLHDRDemod::process: one frame in, the 90 rotated header symbols out; res = {modcod, shortframes, pilots};
 * loop = {phase, freq} of the phase loop after the call */
int orc_plhdr_process(orc_plhdr* p, int count, const float* in, float* out90, int* res, float* loop);
/* dvbs2_pilot_coarse_fed; rn = orc_pl_rn(codenum) */
float orc_coarse_fed(const float* frame, int raw_frame_size, int pilots, int pls_code, const uint8_t* rn);
void orc_plheader_symbols(int pls_code, float* out90);                                /* s2_sof + s2_plscodes symbols */
uint64_t orc_pls_codeword(int pls_code);

/* ---- row 8(f)-2 (oracle_pll.c): S2PLLBlock, the payload phase loop (dvbs2/dvbs2_pll.cpp:5-86, dvbs2_pll.h:47-59) ---- */
typedef struct orc_pll orc_pll;
/* init + the members DVBS2Demod::init sets (module_dvbs2_demod.cpp:59-65) + update(); const_type as orc_const_create */
orc_pll* orc_pll_create(float loop_bw, int const_type, float g1, float g2, int frame_slot_count, int pilots, int pls_code,
                        int codenum);
void orc_pll_destroy(orc_pll* p);
int orc_pll_pilot_cnt(const orc_pll* p);
/* process(): one frame in, (frame_slot_count + 1) * 90 + pilot_cnt * 36 symbols out (returned);
 * state = {pcl.phase, pcl.freq, error} after the call */
int orc_pll_process(orc_pll* p, const float* in, float* out, float* state);

/* ---- row 8(f)-4, byte-domain half (oracle_dvbs.c): the frame loop body of DVBSDemod::process
 *      (dvbs/module_dvbs_demod.cpp:91-106): deinterleave, 8 x RS(204,188), descramble, 8 x 188 bytes out ---- */
typedef struct orc_dvbs_outer orc_dvbs_outer;
orc_dvbs_outer* orc_dvbs_outer_create(void);
void orc_dvbs_outer_destroy(orc_dvbs_outer* p);
/* one frame of 1632 bytes in, 1504 bytes and the eight DVBSReedSolomon::decode return values out */
void orc_dvbs_outer_frame(orc_dvbs_outer* p, const uint8_t* frame, uint8_t* out, int* errors);
/* nframes frames, frame k at frames + k * stride */
void orc_dvbs_outer_process(orc_dvbs_outer* p, const uint8_t* frames, int nframes, int stride, uint8_t* out, int* errors);
/* DVBS_TS_Deframer::work (dvbs/dvbs_ts_deframer.cpp:44-101): unpacked bits (0/1 bytes; other values are ORed in as the
 * reference's pack_8 does) -> frames of 1632 bytes; returns the frame count */
typedef struct orc_dvbs_deframer orc_dvbs_deframer;
orc_dvbs_deframer* orc_dvbs_deframer_create(void);
void orc_dvbs_deframer_destroy(orc_dvbs_deframer* p);
int orc_dvbs_deframer_work(orc_dvbs_deframer* p, const uint8_t* input, int size, uint8_t* output);
void orc_dvbs_deframer_stats(const orc_dvbs_deframer* p, int* errors_nor, int* errors_inv);
/* ---- inner half of the DVB-S chain (oracle_vit.c) ----
 * Viterbi_DVBS as the module constructs it (module_dvbs_demod.cpp:23: phases 0 and 90 degrees, 8192 soft bits per call)
 * behind DVBSVitBlock::process (dvbs_vit.cpp:6-13): count (a multiple of 8192) signed soft bits -> decoded bits, one per
 * byte; returns how many.  out needs room for count bytes; bytes the reference does not write are not written. */
typedef struct orc_vit orc_vit;
orc_vit* orc_vit_create(float ber_threshold, int max_outsync);
void orc_vit_destroy(orc_vit* v);
int orc_vit_process(orc_vit* v, int count, const int8_t* in, uint8_t* out);
/* ber(), getState(), rate() and the lock the decoder holds; rate 0..4 = 1/2, 2/3, 3/4, 5/6, 7/8 */
void orc_vit_stats(const orc_vit* v, float* ber, int* state, int* rate, int* phase, int* shift, int* invalid);
/* DVBSymToSoftBlock::process (dvbs_syms_to_soft.cpp:26-42): count symbols (re, im floats) -> soft bits in chunks of 8192 */
typedef struct orc_sts orc_sts;
orc_sts* orc_sts_create(void);
void orc_sts_destroy(orc_sts* s);
int orc_sts_process(orc_sts* s, int count, const float* syms, int8_t* out);
/* transmit side for the tests: RS(204,188) parity of one packet */
void orc_rs204_parity(const uint8_t* msg188, uint8_t* parity16);

#ifdef __cplusplus
}
#endif
#endif
