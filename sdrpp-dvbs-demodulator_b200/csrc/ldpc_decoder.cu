// K2 -- see ldpc_decoder.cuh for the design.  Reference semantics: SURVEY.md spec S-LDPC
// (layered_decoder.hh:23-74,121-133; algorithms.hh:235-256,261-276; bbframe_ldpc.cpp:123-139).
#include "ldpc_decoder.cuh"
#include "ldpc_variants.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace s2 {
namespace {

const Variant* pick(int max_cnt) {
    const Variant* parts[4] = {kLdpcVariantsA, kLdpcVariantsB, kLdpcVariantsC, kLdpcVariantsD};
    const int counts[4] = {kLdpcVariantsA_n, kLdpcVariantsB_n, kLdpcVariantsC_n, kLdpcVariantsD_n};
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < counts[k]; ++i)
            if (parts[k][i].cnt >= max_cnt) return &parts[k][i];
    return nullptr;
}
KernelFn pick_fn(const LdpcDev& c, bool streamed = false) {
    const int ch = c.chains ? 1 : 0;
    const Variant* v = pick(c.max_cnt);
    if (!v) return nullptr;
    bool uniform = v->cnt == c.max_cnt;
    for (int i = 0; i < c.q && uniform; ++i) uniform = (c.layer_off[i + 1] - c.layer_off[i]) == v->cnt;
    const int oc = c.occ3 ? 1 : 0;
    return (uniform && v->uniform[0][0][0]) ? v->uniform[streamed][ch][oc] : v->ragged[streamed][ch][oc];
}

}  // namespace

// Opt every launch of `fn` in to the device's full dynamic shared memory.  The value is the same for every
// code that shares a kernel instantiation, so handles configured for different MODCODs can launch concurrently
// from different host threads without racing on the function attribute.
static cudaError_t allow_max_smem(KernelFn fn) {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

// Codes for which the three-phase treatment of chained layers beats the level-by-level one.  Measured with
// tools/modcod_sweep.py (2048 normal / 8192 short frames, fixed noise) against the same kernel without it:
// n1/3 -9 %, n2/3 -3..-6 %, n3/4 -12 %, n8/9 -20 %, s2/5 -12 %, s1/2 -3 %, s2/3 -9 % time (s3/4 -10 %, but it
// gains more from a second CTA with the lean row update, which has no chained mode); within +-2 %
// or slower for the others (few layers with many levels, or their other conflicted layers dominate; for n1/2
// the co-resident CTA already hides the level steps at full load).  Only layers with at least eight levels are
// chained: the three phases cost about two extra passes over the rows.  Index = code table order B1..B11, C1..C10.
bool ldpc_chains_pay_off(int code_index) {
    static const bool table[21] = {false, true,  false, false, false, true,  true,  false, false, true,  false,
                                   false, false, true,  true,  false, true,  false, false, false, false};
    static const int force = [] { const char* e = getenv("DVBS2FEC_LDPC_CHAINS"); return e ? atoi(e) : -1; }();
    if (force >= 0) return force != 0;
    return code_index >= 0 && code_index < 21 && table[code_index];
}

// Codes that run faster with one more resident CTA per SM than the default and the register allocation squeezed
// accordingly (three CTAs / 56 registers for CNT <= 9, two CTAs / 80 registers above): their shared memory
// allows it.  Measured with tools/modcod_sweep.py: n1/4 -11 %, n1/3 -13 %, n2/5 -2 %, s1/4 -6 %, s1/3 -9 %,
// s2/5 -10 %, s1/2 -5 %, s3/5 -7 %, s2/3 -10 %, s3/4 -17 %, s4/5 -25 %, s5/6 -19 % time (above CNT = 9 with the
// lean row update, row_update_lean, which keeps them from spilling).  n1/2, n3/5, n2/3 and the normal codes
// above cannot hold the extra CTA (shared memory) and lose 3-11 % (up to 2x with spills) to the tighter
// allocation; s8/9 still spills 216 B and loses.  Index = code table order B1..B11, C1..C10.
bool ldpc_ctas_wanted3(int code_index) {
    static const bool table[21] = {true,  true,  true,  false, false, false, false, false, false, false, false,
                                   true,  true,  true,  true,  true,  true,  true,  true,  true,  false};
    static const int force = [] { const char* e = getenv("DVBS2FEC_LDPC_OCC3"); return e ? atoi(e) : -1; }();
    if (force >= 0) return force != 0;
    return code_index >= 0 && code_index < 21 && table[code_index];
}

int ldpc_slot_groups(int max_cnt) {
    const Variant* v = pick(max_cnt);
    return v ? (v->cnt + 2 + 7) / 8 : 0;
}

int ldpc_max_ctas_per_sm(const LdpcDev& code) {
    KernelFn fn = pick_fn(code);
    if (!fn) return 0;
    size_t smem = ldpc_smem_bytes(code);
    if (allow_max_smem(fn) != cudaSuccess) return 0;
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kLdpcThreads, smem);
    return n;
}

int ldpc_launch(const LdpcArgs& a, int grid, cudaStream_t stream) {
    const LdpcDev& c = a.code;
    const Variant* v = pick(c.max_cnt);
    KernelFn fn = pick_fn(c, a.arrived != nullptr);
    if (!v || !fn || c.q < 1 || c.q > kMaxLayers) return (int)cudaErrorInvalidValue;
    LdpcParams p;
    p.N = c.N; p.K = c.K; p.R = c.R; p.q = c.q; p.ngroups = c.ngroups;
    p.sg = (v->cnt + 2 + 7) / 8;
    if (p.sg != c.sg) return (int)cudaErrorInvalidValue;
    p.nframes = a.nframes; p.max_trials = a.max_trials; p.hard_stride = a.hard_stride; p.pad_ = 0;
    p.llr_in = a.llr_in; p.hard_out = a.hard_out; p.iters_out = a.iters_out; p.llr_out = a.llr_out;
    p.workspace = a.workspace;
    p.ws_stride = ldpc_workspace_bytes(c);
    p.work_counter = a.work_counter;
    p.arrived = a.arrived;
    p.row_level = c.row_level;
    const int nlinks = c.layer_off[c.q];
    if (nlinks > kMaxLinks) return (int)cudaErrorInvalidValue;
    for (int i = 0; i <= c.q; ++i) p.layer_off[i] = (uint16_t)c.layer_off[i];
    for (int i = 0; i < c.q; ++i) p.layer_nlev[i] = c.layer_nlev[i];
    {   // Barrier elision.  The reference visits rows strictly in order, but two rows commute unless they touch a
        // common bit.  Parity bits chain rows of the same thread only (kept in registers), except pty[q-1][j-1]
        // (layer 0 of thread j, layer q-1 of thread j-1).  Data bits are shared between threads, group by group:
        // a barrier is needed before a layer exactly when it touches a 360-bit group that some layer since the
        // last barrier touched; level-scheduled layers are always fenced on both sides.
        std::vector<char> seen(c.ngroups, 0);
        bool any = false;
        for (int i = 0; i < c.q; ++i) {
            const bool multi = c.layer_nlev[i] > 1;
            bool conflict = multi;
            for (int k = c.layer_off[i]; k < c.layer_off[i + 1] && !conflict; ++k) conflict = seen[c.links[k] >> 16] != 0;
            if (i > 0 && conflict) {
                p.layer_sync[i - 1] = 1;
                any = any || (i - 1 <= c.q - 3);
                std::fill(seen.begin(), seen.end(), 0);
            }
            p.layer_sync[i] = 0;
            for (int k = c.layer_off[i]; k < c.layer_off[i + 1]; ++k) seen[c.links[k] >> 16] = 1;
            if (multi) {
                p.layer_sync[i] = 1;
                any = any || (i <= c.q - 3);
                std::fill(seen.begin(), seen.end(), 0);
            }
        }
        if (c.q > 0) p.layer_sync[c.q - 1] = 1;   // end of the pass
        if (!any) p.layer_sync[0] = 1;        // pty[q-1][j-1]: stored in layer 0, prefetched (by thread j-1) in layer q-2
    }
    for (int i = 0; i < nlinks; ++i) p.links[i] = c.links[i];
    {   // Chained layers: exactly one 360-bit group is hit by exactly two links of the layer (they are adjacent,
        // links are sorted by group).  X is the one whose bit the EARLIER row of a pair owns: row j's X bit is row
        // j+d's Y bit with d = (shift_Y - shift_X) mod 360 <= 180.
        const bool enabled = true;
        // a chained layer costs about two extra passes over the rows; below this many levels the level loop is cheaper
        static const int min_levels = [] { const char* e = getenv("DVBS2FEC_CHAIN_MINLEV"); return e ? atoi(e) : 8; }();
        for (int i = 0; i < c.q; ++i) {
            p.layer_chain[i] = 0;
            if (!enabled || !c.chains || (int)c.layer_nlev[i] < min_levels) continue;
            int pairs = 0, first = -1;
            bool simple = true;
            for (int k = c.layer_off[i]; k + 1 < c.layer_off[i + 1]; ++k) {
                if ((c.links[k] >> 16) != (c.links[k + 1] >> 16)) continue;
                if (k + 2 < c.layer_off[i + 1] && (c.links[k + 2] >> 16) == (c.links[k] >> 16)) simple = false;   // three in a group
                ++pairs;
                first = k - c.layer_off[i];
            }
            if (!simple || pairs != 1 || first > 30) continue;
            const int s0 = (int)(c.links[c.layer_off[i] + first] & 0xFFFFu), s1 = (int)(c.links[c.layer_off[i] + first + 1] & 0xFFFFu);
            int d = ((s1 - s0) % 360 + 360) % 360, orient = 0;   // X = first, Y = first + 1
            if (d > 180) {
                d = 360 - d;
                orient = 1;                                      // X = first + 1, Y = first
            }
            if (d == 0) continue;
            p.layer_chain[i] = (uint16_t)(0x8000u | (unsigned)d << 6 | (unsigned)orient << 5 | (unsigned)first);
        }
    }
    size_t smem = ldpc_smem_bytes(c);
    cudaError_t e = allow_max_smem(fn);
    if (e != cudaSuccess) return (int)e;
    fn<<<grid, kLdpcThreads, smem, stream>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace s2
