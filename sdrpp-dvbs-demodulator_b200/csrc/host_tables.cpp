// See host_tables.h.
#include "host_tables.h"

#include <cmath>
#include <limits>
#include <cstring>
#include <mutex>

namespace s2 {

// ------------------------------------------------------------------ GF(2^m)
const GfHost& gf_host(int m) {
    static GfHost f14, f16;
    static std::once_flag once;
    std::call_once(once, [] {
        for (GfHost* f : {&f14, &f16}) {
            const int mm = (f == &f14) ? 14 : 16;
            const uint32_t poly = (mm == 14) ? 0x402Bu : 0x1002Du;
            f->m = mm;
            f->N = (1 << mm) - 1;
            f->log.assign(1 << mm, 0);
            f->exp.assign(1 << mm, 0);
            f->log[0] = (uint16_t)f->N;
            f->exp[f->N] = 0;
            uint32_t a = 1;
            for (int i = 0; i < f->N; ++i) {
                f->exp[i] = (uint16_t)a;
                f->log[a] = (uint16_t)i;
                a <<= 1;
                if (a >> mm) a ^= poly;
            }
        }
    });
    return m == 14 ? f14 : f16;
}

// ------------------------------------------------------------------ BCH
namespace {
// minimal polynomial over GF(2) of alpha^r, as a bit mask (bit k = coeff of x^k)
uint32_t minimal_poly(const GfHost& f, int r) {
    std::vector<uint32_t> p{1};  // polynomial with coefficients in GF(2^m), p[k] = coeff of x^k
    int e = r;
    do {
        uint32_t root = f.exp[e];
        std::vector<uint32_t> nx(p.size() + 1, 0);
        for (size_t k = 0; k < p.size(); ++k) {
            nx[k + 1] ^= p[k];
            nx[k] ^= f.mul(p[k], root);
        }
        p.swap(nx);
        e = (int)(((long long)e * 2) % f.N);
    } while (e != r);
    uint32_t mask = 0;
    for (size_t k = 0; k < p.size(); ++k) mask |= (p[k] & 1u) << k;  // all coefficients are 0/1
    return mask;
}
}  // namespace

const BchHost& bch_host(int m, int t) {
    static std::mutex mu;
    static std::vector<BchHost*> cache;
    std::lock_guard<std::mutex> lk(mu);
    for (BchHost* b : cache)
        if (b->m == m && b->t == t) return *b;
    const GfHost& f = gf_host(m);
    BchHost* b = new BchHost();
    b->m = m;
    b->t = t;
    b->np = m * t;
    b->gen.assign(1, 1);
    b->crc.assign((size_t)t * 256, 0);
    b->basis.assign((size_t)t * 16, 0);
    for (int k = 0; k < t; ++k) {
        const int i = 2 * k + 1;
        const uint32_t mp = minimal_poly(f, i);  // degree m for every odd i <= 23 in both fields
        std::vector<uint8_t> ng(b->gen.size() + m, 0);
        for (size_t x = 0; x < b->gen.size(); ++x)
            if (b->gen[x])
                for (int y = 0; y <= m; ++y) ng[x + y] ^= (mp >> y) & 1u;
        b->gen.swap(ng);
        for (int idx = 0; idx < 256; ++idx) {  // (idx * x^m) mod mp
            uint32_t v = (uint32_t)idx << m;
            for (int bit = m + 7; bit >= m; --bit)
                if (v >> bit & 1u) v ^= mp << (bit - m);
            b->crc[(size_t)k * 256 + idx] = (uint16_t)v;
        }
        for (int bit = 0; bit < m; ++bit) b->basis[(size_t)k * 16 + bit] = f.exp[(int)(((long long)i * bit) % f.N)];
    }
    cache.push_back(b);
    return *b;
}

void bch_encode(const BchHost& code, uint8_t* frame, int kbch) {
    const int np = code.np;
    std::vector<uint8_t> reg(np, 0);  // reg[k] = coeff of x^k of (data(x) x^np) mod g(x) so far
    for (int i = 0; i < kbch; ++i) {
        int fb = ((frame[i >> 3] >> (7 - (i & 7))) & 1) ^ reg[np - 1];
        for (int k = np - 1; k > 0; --k) reg[k] = reg[k - 1] ^ (fb & code.gen[k]);
        reg[0] = fb & code.gen[0];
    }
    uint8_t* par = frame + kbch / 8;
    memset(par, 0, np / 8);
    for (int i = 0; i < np; ++i)
        if (reg[np - 1 - i]) par[i >> 3] |= 0x80 >> (i & 7);
}

std::vector<uint8_t> pl_scrambling_rn(int codenum, int count) {
    // x: 1 + x^7 + x^18 from state 0...01, advanced by the code number; y: 1 + y^5 + y^7 + y^10 + y^18 from all
    // ones; z(i) = x(i) ^ y(i); Rn(i) = z(i) + 2 z(i + 2^17).  Two copies of the generator 2^17 steps apart.
    auto step_x = [](uint32_t v) { return ((((v >> 7) ^ v) & 1u) << 18 | v) >> 1; };
    auto step_y = [](uint32_t v) { return ((((v >> 10) ^ (v >> 7) ^ (v >> 5) ^ v) & 1u) << 18 | v) >> 1; };
    uint32_t x0 = 1, y0 = 0x3FFFF;
    for (int i = 0; i < codenum; ++i) x0 = step_x(x0);
    uint32_t x1 = x0, y1 = y0;
    for (int i = 0; i < 131072; ++i) {
        x1 = step_x(x1);
        y1 = step_y(y1);
    }
    std::vector<uint8_t> rn((size_t)count);
    for (int i = 0; i < count; ++i) {
        rn[i] = (uint8_t)(((x0 ^ y0) & 1u) | (((x1 ^ y1) & 1u) << 1));
        x0 = step_x(x0);
        y0 = step_y(y0);
        x1 = step_x(x1);
        y1 = step_y(y1);
    }
    return rn;
}

const std::vector<uint8_t>& bb_prbs() {
    static std::vector<uint8_t> seq;
    static std::once_flag once;
    std::call_once(once, [] {
        seq.assign(8100, 0);
        // 15-stage register, taps 14 and 15, loaded with 100101010000000; held here with stage 15 in bit 0
        uint32_t sr = 0x4A80;
        for (int i = 0; i < 64800; ++i) {
            uint32_t b = (sr ^ (sr >> 1)) & 1u;
            if (b) seq[i >> 3] |= 0x80 >> (i & 7);
            sr = (sr >> 1) | (b << 14);
        }
    });
    return seq;
}

// ------------------------------------------------------------------ LDPC encoder
void ldpc_encode_bits(int code, const uint8_t* data_bits, uint8_t* code_bits) {
    const RawTable& t = raw_table(code);
    const int R = t.N - t.K, q = R / kGroup;
    uint8_t* par = code_bits + t.K;
    memcpy(code_bits, data_bits, t.K);
    memset(par, 0, R);
    const uint16_t* row = t.addr;
    int bit = 0;
    for (int r = 0; r < t.nruns; ++r)
        for (int g = 0; g < t.len[r]; ++g) {
            for (int m = 0; m < kGroup; ++m, ++bit) {
                if (!(data_bits[bit] & 1)) continue;
                for (int d = 0; d < t.deg[r]; ++d) par[(row[d] + q * m) % R] ^= 1;
            }
            row += t.deg[r];
        }
    for (int i = 1; i < R; ++i) par[i] ^= par[i - 1];
}

// ------------------------------------------------------------------ constellations / demapper LUT
namespace {
struct RingPos { int8_t ring; float k; };  // ring 1..3, phase = k * 2pi / points_on_ring
const RingPos k16[16] = {{1, 2.5f}, {1, 1.5f}, {1, 3.5f}, {1, 0.5f}, {2, 8.5f}, {2, 3.5f}, {2, 9.5f}, {2, 2.5f},
                         {2, 6.5f}, {2, 5.5f}, {2, 11.5f}, {2, 0.5f}, {2, 7.5f}, {2, 4.5f}, {2, 10.5f}, {2, 1.5f}};
const RingPos k32[32] = {{3, 10}, {3, 8}, {3, 5}, {3, 7}, {3, 13}, {3, 15}, {3, 2}, {3, 0},
                         {1, 2.5f}, {2, 6.5f}, {1, 1.5f}, {2, 5.5f}, {1, 3.5f}, {2, 11.5f}, {1, 0.5f}, {2, 0.5f},
                         {3, 11}, {3, 9}, {3, 4}, {3, 6}, {3, 12}, {3, 14}, {3, 3}, {3, 1},
                         {2, 8.5f}, {2, 7.5f}, {2, 3.5f}, {2, 4.5f}, {2, 9.5f}, {2, 10.5f}, {2, 2.5f}, {2, 1.5f}};

void place(ConstellationHost& c, int idx, float r, int n, float k) {
    float a = (float)(k * 2 * 3.14159265358979323846 / n);
    c.re[idx] = (r * cosf(a)) * c.amp;
    c.im[idx] = (r * sinf(a)) * c.amp;
}

int8_t halving_clamp(float x) {  // constellation_t::clamp: halve until inside +-127, then truncate
    while (x < -127 || x > 127) {
        x *= 0.5f;
        if (!std::isfinite(x)) return (int8_t)(int)x;
    }
    return (int8_t)x;
}
}  // namespace

ConstellationHost make_constellation(Constellation type, float g1, float g2) {
    ConstellationHost c{};
    c.type = type;
    c.bits = 2 + (int)type;
    c.states = 1 << c.bits;
    c.amp = 1.0f;
    c.sca = 50.0f;
    c.prescale = 1.0f;
    if (type == QPSK) {
        c.amp = 3;
        const float s = (float)1.41421356237309504880;
        for (int i = 0; i < 4; ++i) {
            c.re[i] = (i & 1) ? s : -s;
            c.im[i] = (i & 2) ? s : -s;
        }
    } else if (type == PSK8) {
        const float h = 0.70710678118654752440f;
        const float pr[8] = {0.0f, -h, h, 0.0f, -h, -1.0f, 1.0f, h};
        const float pi[8] = {-1.0f, h, -h, 1.0f, -h, 0.0f, 0.0f, h};
        memcpy(c.re, pr, sizeof(pr));
        memcpy(c.im, pi, sizeof(pi));
    } else if (type == APSK16) {
        c.amp = 100;
        c.sca = 1;
        c.prescale = 0.53f;
        float gamma = g1 ? g1 : 2.57f;
        float r1 = sqrtf(4 / (1 + 3 * gamma * gamma));
        float r2 = gamma * r1;
        r1 *= 0.5f;
        r2 *= 0.5f;
        for (int i = 0; i < 16; ++i) place(c, i, k16[i].ring == 1 ? r1 : r2, k16[i].ring == 1 ? 4 : 12, k16[i].k);
    } else {
        c.amp = 100;
        c.sca = 1;
        c.prescale = 0.54f;
        float gamma1 = g1 ? g1 : 2.53f, gamma2 = g2 ? g2 : 4.30f;
        float r1 = sqrtf(8 / (1 + 3 * gamma1 * gamma1 + 4 * gamma2 * gamma2));
        float r2 = gamma1 * r1, r3 = gamma2 * r1;
        r1 *= 0.5f;
        r2 *= 0.5f;
        r3 *= 0.5f;
        for (int i = 0; i < 32; ++i) {
            int ring = k32[i].ring;
            place(c, i, ring == 1 ? r1 : ring == 2 ? r2 : r3, ring == 1 ? 4 : ring == 2 ? 12 : 16, k32[i].k);
        }
    }
    return c;
}

void demap_calc(const ConstellationHost& c, float re, float im, int8_t* bits) {
    float acc[10] = {0};
    if (c.amp != 1) { re = re * c.amp; im = im * c.amp; }
    if (c.prescale != 1) { re = re * c.prescale; im = im * c.prescale; }
    for (int i = 0; i < c.states; ++i) {
        float dr = re - c.re[i], di = im - c.im[i];
        float d = expf(-sqrtf((dr * dr) + (di * di)) / 1.0f);
        for (int j = 0; j < c.bits; ++j) acc[2 * j + ((i >> j) & 1)] += d;
    }
    for (int j = 0; j < c.bits; ++j)
        bits[c.bits - 1 - j] = halving_clamp((logf(acc[2 * j + 1]) - logf(acc[2 * j])) * c.sca);
}

std::vector<uint32_t> demap_lut(const ConstellationHost& c) {
    std::vector<uint32_t> lut(256 * 256);
    for (int x = 0; x < 256; ++x)
        for (int y = 0; y < 256; ++y) {
            float xv = ((float)(x - 128) / 256.0f) * 1.5f, yv = ((float)(y - 128) / 256.0f) * 1.5f;
            int8_t b[5] = {0, 0, 0, 0, 0};
            demap_calc(c, xv, yv, b);
            uint32_t w = 0;
            for (int k = 0; k < c.bits && k < 4; ++k) w |= (uint32_t)(uint8_t)b[k] << (8 * k);
            lut[x * 256 + y] = w;
        }
    return lut;
}

float demap_phase_error(const ConstellationHost& c, float re, float im) {
    if (c.amp != 1) { re = re * c.amp; im = im * c.amp; }
    if (c.prescale != 1) { re = re * c.prescale; im = im * c.prescale; }
    float best = std::numeric_limits<float>::max(), cr = 0, ci = 0;
    for (int i = 0; i < c.states; ++i) {
        float dr = re - c.re[i], di = im - c.im[i];
        float dist = sqrtf((dr * dr) + (di * di));
        if (dist < best) {       // the first point among equally near ones
            best = dist;
            cr = c.re[i];
            ci = c.im[i];
        }
    }
    // (sample * closest.conj()).phase()
    const float bi = -ci;
    const float pr = (re * cr) - (im * bi), pi = (im * cr) + (re * bi);
    return atan2f(pi, pr);
}

std::vector<float> demap_phase_lut(const ConstellationHost& c) {
    std::vector<float> lut(256 * 256);
    for (int x = 0; x < 256; ++x)
        for (int y = 0; y < 256; ++y) {
            float xv = ((float)(x - 128) / 256.0f) * 1.5f, yv = ((float)(y - 128) / 256.0f) * 1.5f;
            lut[x * 256 + y] = demap_phase_error(c, xv, yv);
        }
    return lut;
}

void map_symbol(const ConstellationHost& c, const uint8_t* code_bits, float* re_im) {
    int label = 0;
    for (int k = 0; k < c.bits; ++k) label = (label << 1) | (code_bits[k] & 1);
    int idx = ~label & (c.states - 1);
    re_im[0] = (c.re[idx] / c.amp) / c.prescale;
    re_im[1] = (c.im[idx] / c.amp) / c.prescale;
}

int interleaved_position(const ModcodCfg& cfg, int n) {
    const int N = cfg.shortframes ? 16200 : 64800;
    if (cfg.constellation == QPSK) return n ^ 1;  // the reference swaps every LLR pair
    const int bits = cfg.bits, rows = N / bits;
    int col = n / rows, j = n - col * rows;
    if (cfg.constellation == PSK8 && cfg.rate == R3_5) col = 2 - col;  // 8PSK 3/5 reads columns 2-1-0
    return j * bits + col;
}

}  // namespace s2
