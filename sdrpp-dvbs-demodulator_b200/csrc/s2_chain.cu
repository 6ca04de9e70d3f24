// The decode stage of the reference's DVB-S2 module as one call: DVBS2Demod::process (dvbs2/module_dvbs2_demod.cpp:300-367)
// behind its sample-domain front end (AGC, RRC, clock recovery, frequency shifter -- SDR++ DSP, out of scope):
//   symbols -> PL sync (K7) -> per frame: coarse frequency error (K7), payload phase loop with PL descrambling (K8),
//   PLHEADER demodulation (K7), demapper (K1), LDPC (K2), BCH + BB descrambler (K3) -> BBFRAMEs
// Symbols go up once; frames, derotated symbols, LLRs and hard decisions never leave the device.  The objects are the
// library's own C-ABI handles; this file strings their device entry points together on one stream, in the reference's
// order.  Two things the reference's loop does are left to the caller because they belong to its sample-domain side: the
// coarse frequency error is handed back per frame (the module feeds it to its frequency shifter, :302-313, which acts on
// LATER samples anyway), and nothing waits for 16 frames before decoding (SURVEY.md note N1).
#include "../../include/dvbs2fec.h"

#include <algorithm>
#include <cstdio>
#include <memory>

#include <cuda_runtime.h>

namespace s2 {
int api_fail(int code, const char* msg);
}
using s2::api_fail;

namespace {
int failf(int code, const char* what, cudaError_t e) {
    char buf[300];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return api_fail(code, buf);
}
#define CU(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return failf(DVBS2FEC_ECUDA, #call, e_); \
    } while (0)

template <typename T>
cudaError_t reserve(T*& p, size_t& cap, size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}
}  // namespace

struct dvbs2fec_s2_demod {
    int device = 0;
    dvbs2fec_handle* fec = nullptr;
    dvbs2fec_plsync* pl = nullptr;
    dvbs2fec_ts_parser* ts = nullptr;      // BBFRAME -> TS / GSE (K6), for dvbs2fec_s2_demod_process_ts
    cudaStream_t stream = nullptr;
    int modcod = 0, shortframes = 0, pilots = 0, codenum = 0, rfs = 0, kb = 0;
    float* d_x = nullptr; size_t x_cap = 0;
    float* d_fr = nullptr; size_t fr_cap = 0;
    float* d_pl = nullptr; size_t pl_cap = 0;
    float* d_hsym = nullptr; size_t hsym_cap = 0;
    uint8_t* d_bb = nullptr; size_t bb_cap = 0;
    dvbs2fec_result* d_res = nullptr; size_t res_cap = 0;
    float* d_fed = nullptr; size_t fed_cap = 0;
    int32_t* d_hdr = nullptr; size_t hdr_cap = 0;
    uint8_t* d_ts = nullptr; size_t ts_cap = 0;
    int* d_n = nullptr;
    int last_frames = 0;
};

extern "C" {

void dvbs2fec_s2_demod_destroy(dvbs2fec_s2_demod* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) {
        cudaStreamSynchronize(p->stream);
        cudaStreamDestroy(p->stream);
    }
    dvbs2fec_plsync_destroy(p->pl);
    dvbs2fec_ts_destroy(p->ts);
    dvbs2fec_destroy(p->fec);
    cudaFree(p->d_x); cudaFree(p->d_fr); cudaFree(p->d_pl); cudaFree(p->d_hsym); cudaFree(p->d_bb); cudaFree(p->d_res);
    cudaFree(p->d_fed); cudaFree(p->d_hdr); cudaFree(p->d_ts); cudaFree(p->d_n);
    delete p;
}

int dvbs2fec_s2_demod_create(const dvbs2fec_config* cfg, dvbs2fec_s2_demod** out) {
    if (!out) return api_fail(DVBS2FEC_EINVAL, "out is NULL");
    *out = nullptr;
    std::unique_ptr<dvbs2fec_s2_demod, void (*)(dvbs2fec_s2_demod*)> p(new dvbs2fec_s2_demod(), dvbs2fec_s2_demod_destroy);
    dvbs2fec_config c{};
    if (cfg) c = *cfg;
    if (c.n_devices > 1) return api_fail(DVBS2FEC_EINVAL, "one stream of symbols is a recurrence (PL sync, phase loops): one device");
    int dev = 0;
    if (c.n_devices == 1) dev = c.devices[0];
    else if (cudaGetDevice(&dev) != cudaSuccess) return api_fail(DVBS2FEC_ENODEV, "no CUDA device");
    c.n_devices = 1;
    c.devices[0] = dev;
    if (c.max_batch <= 0) c.max_batch = 1024;
    p->device = dev;
    int rc = dvbs2fec_create(&c, &p->fec);
    if (!rc) rc = dvbs2fec_plsync_create(dev, &p->pl);
    if (!rc) rc = dvbs2fec_ts_create(dev, &p->ts);
    if (rc) return rc;
    CU(cudaSetDevice(dev));
    CU(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&p->d_n, sizeof(int)));
    *out = p.release();
    return 0;
}

int dvbs2fec_s2_demod_set_params(dvbs2fec_s2_demod* p, int modcod, int shortframes, int pilots, int max_trials, float pll_loop_bw,
                                 float plhdr_loop_bw, int codenum) {
    if (!p) return api_fail(DVBS2FEC_EINVAL, "handle is NULL");
    int nldpc = 0, bits = 0, plsyms = 0;
    int rc = dvbs2fec_modcod_info(modcod, shortframes, pilots, &nldpc, nullptr, nullptr, nullptr, &bits, &plsyms, nullptr);
    if (rc) return rc;
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->stream));
    rc = dvbs2fec_set_modcod(p->fec, modcod, shortframes, pilots, max_trials);      // DVBS2Demod::setDemodParams (:118-168)
    if (!rc) rc = dvbs2fec_set_pl_scrambling(p->fec, -1);                            // the phase loop hands over descrambled symbols
    if (!rc) rc = dvbs2fec_plsync_set_params(p->pl, nldpc / bits / 90, pilots);
    if (!rc) rc = dvbs2fec_plhdr_set_params(p->pl, plhdr_loop_bw);
    if (!rc) rc = dvbs2fec_pll_set_params(p->pl, pll_loop_bw, modcod, shortframes, pilots, codenum);
    if (rc) return rc;
    p->modcod = modcod; p->shortframes = !!shortframes; p->pilots = !!pilots; p->codenum = codenum;
    p->rfs = dvbs2fec_plsync_raw_frame_size(p->pl);
    p->kb = dvbs2fec_kbch(p->fec) / 8;
    rc = dvbs2fec_ts_set_frame_size(p->ts, p->kb * 8);      // BBFrameTSParser::setFrameSize (main.cpp, on a MODCOD change)
    if (rc) return rc;
    if (p->rfs != plsyms || dvbs2fec_pll_frame_symbols(p->pl) > p->rfs) return api_fail(DVBS2FEC_EINVAL, "frame sizes of the stages disagree");
    return 0;
}

int dvbs2fec_s2_demod_reset(dvbs2fec_s2_demod* p) {
    if (!p || !p->rfs) return api_fail(DVBS2FEC_EINVAL, "set_params has not been called");
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->stream));
    int rc = dvbs2fec_plsync_reset(p->pl);
    if (!rc) rc = dvbs2fec_pll_reset(p->pl);
    return rc;
}

int dvbs2fec_s2_demod_bbframe_bytes(const dvbs2fec_s2_demod* p) { return p ? p->kb : 0; }
int dvbs2fec_s2_demod_max_frames(const dvbs2fec_s2_demod* p, int count) { return p && p->rfs ? count / p->rfs + 2 : 0; }

}  // extern "C"

namespace {
// symbols in host memory -> BBFRAMEs, results, frequency errors and PLHEADER fields on the device (enqueued, not waited for);
// returns the number of frames
int stage_to_device(dvbs2fec_s2_demod* p, int count, const float* syms) {
    const int room = count / p->rfs + 2;      // fewer than two frames are ever carried over
    CU(cudaSetDevice(p->device));
    cudaStream_t st = p->stream;
    const size_t fsz = (size_t)room * p->rfs * 2;
    CU(reserve(p->d_x, p->x_cap, (size_t)2 * count));
    CU(reserve(p->d_fr, p->fr_cap, fsz));
    if (fsz > p->pl_cap) {
        CU(reserve(p->d_pl, p->pl_cap, fsz));
        CU(cudaMemsetAsync(p->d_pl, 0, fsz * sizeof(float), st));      // symbols behind the ones the loop handles are never written
    }
    CU(reserve(p->d_hsym, p->hsym_cap, (size_t)room * 180));
    CU(reserve(p->d_bb, p->bb_cap, (size_t)room * p->kb));
    CU(reserve(p->d_res, p->res_cap, (size_t)room));
    CU(reserve(p->d_fed, p->fed_cap, (size_t)room));
    CU(reserve(p->d_hdr, p->hdr_cap, (size_t)room * 4));
    CU(cudaMemcpyAsync(p->d_x, syms, sizeof(float) * 2 * count, cudaMemcpyHostToDevice, st));
    int rc = dvbs2fec_plsync_process_device(p->pl, count, p->d_x, p->d_fr, room, p->d_n, st);      // :300
    if (rc) return rc;
    int n = 0;
    CU(cudaMemcpyAsync(&n, p->d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n <= 0) return 0;
    const int pls_code = p->modcod << 2 | p->shortframes << 1 | p->pilots;
    rc = dvbs2fec_coarse_fed_device(p->pl, n, p->d_fr, p->pilots, pls_code, p->codenum, p->d_fed, st);      // :302
    if (!rc) rc = dvbs2fec_pll_process_device(p->pl, n, p->rfs, p->d_fr, p->d_pl, nullptr, st);             // :314
    if (!rc) rc = dvbs2fec_plhdr_process_device(p->pl, n, p->d_fr, p->d_hsym, p->d_hdr, st);                // :315
    if (!rc) rc = dvbs2fec_decode_plframes_device(p->fec, p->d_pl, n, p->d_bb, p->d_res, st);               // :316, 334-366
    if (rc) return rc;
    return n;
}

int copy_side_outputs(dvbs2fec_s2_demod* p, int n, dvbs2fec_result* results, float* fed_err, int32_t* plhdr) {
    cudaStream_t st = p->stream;
    if (results) CU(cudaMemcpyAsync(results, p->d_res, sizeof(dvbs2fec_result) * n, cudaMemcpyDeviceToHost, st));
    if (fed_err) CU(cudaMemcpyAsync(fed_err, p->d_fed, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (plhdr) CU(cudaMemcpyAsync(plhdr, p->d_hdr, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost, st));
    return 0;
}
}  // namespace

extern "C" {

int dvbs2fec_s2_demod_process(dvbs2fec_s2_demod* p, int count, const float* syms, uint8_t* bb_out, int max_frames, dvbs2fec_result* results,
                              float* fed_err, int32_t* plhdr) {
    if (!p || !p->rfs) return api_fail(DVBS2FEC_EINVAL, "set_params has not been called");
    if (count < 0 || (count && !syms) || !bb_out) return api_fail(DVBS2FEC_EINVAL, "bad arguments");
    if (max_frames < count / p->rfs + 2) return api_fail(DVBS2FEC_EINVAL, "max_frames < count / raw_frame_size + 2 (dvbs2fec_s2_demod_max_frames)");
    if (!count) return 0;
    const int n = stage_to_device(p, count, syms);
    if (n <= 0) return n;
    cudaStream_t st = p->stream;
    CU(cudaMemcpyAsync(bb_out, p->d_bb, (size_t)n * p->kb, cudaMemcpyDeviceToHost, st));
    int rc = copy_side_outputs(p, n, results, fed_err, plhdr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    return n;
}

int dvbs2fec_s2_demod_process_ts(dvbs2fec_s2_demod* p, int count, const float* syms, uint8_t* ts_out, int ts_cap, int max_frames, int* nframes,
                                 dvbs2fec_result* results, float* fed_err, int32_t* plhdr) {
    if (!p || !p->rfs) return api_fail(DVBS2FEC_EINVAL, "set_params has not been called");
    if (count < 0 || (count && !syms) || !ts_out || ts_cap <= 0) return api_fail(DVBS2FEC_EINVAL, "bad arguments");
    if ((results || fed_err || plhdr) && max_frames < count / p->rfs + 2)
        return api_fail(DVBS2FEC_EINVAL, "max_frames < count / raw_frame_size + 2 (dvbs2fec_s2_demod_max_frames)");
    if (nframes) *nframes = 0;
    if (!count) return 0;
    const int n = stage_to_device(p, count, syms);
    if (n <= 0) return n;
    if (nframes) *nframes = n;
    cudaStream_t st = p->stream;
    CU(reserve(p->d_ts, p->ts_cap, (size_t)ts_cap));
    int rc = dvbs2fec_ts_work_device(p->ts, p->d_bb, n, p->d_ts, ts_cap, p->d_n, st);      // main.cpp:538: ts_parser.work(...)
    if (rc) return rc;
    int produced = 0;
    CU(cudaMemcpyAsync(&produced, p->d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
    rc = copy_side_outputs(p, n, results, fed_err, plhdr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    if (produced < 0) return api_fail(produced, "TS / GSE output does not fit ts_cap");
    if (produced) {
        CU(cudaMemcpyAsync(ts_out, p->d_ts, (size_t)produced, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return produced;
}

}  // extern "C"
