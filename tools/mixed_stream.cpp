// Mixed-MODCOD stream measurement: T logical transponders, each its own
// dvbs2fec handle and host thread (as one DVBS2Demod instance per transponder would be), each with its own
// MODCOD, all feeding the same GPU(s) through the frame queue: dvbs2fec_submit_llr -> dvbs2fec_collect.
// Reports aggregate frames/s, information Gbit/s and the submit->collect latency distribution per batch size.
//
// Input frames are in-tree transmitter output (dvbs2fec_encode_fecframe) through a BPSK-equivalent AWGN
// channel per code bit, quantised like bench.py's L4 generator (rint(4*LLR), clamp +-127), at the per-rate
// Es/N0 of tools/modcod_sweep.py.  Only the public C ABI is used.
//
//   g++ -O2 -std=c++17 tools/mixed_stream.cpp -Iinclude -Lsdrpp-dvbs-demodulator_b200 -ldvbs2fec -lpthread
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "dvbs2fec.h"

namespace {

using Clock = std::chrono::steady_clock;

struct Stream {
    int modcod, shortframes;
    int N = 0, kbch = 0;
    dvbs2fec_handle* h = nullptr;
    std::vector<int8_t> pool;   // P frames of N LLRs
    int pool_frames = 0;
    // results
    std::vector<float> lat_us;
    long frames = 0, failed = 0, iters = 0;
};

// QPSK-equivalent Es/N0 (dB) at which each rate decodes in a handful of iterations (tools/modcod_sweep.py)
double esn0_for(int rate_index, bool shortframes) {
    static const double snr[12] = {-2.0, -1.0, 0.0, 1.3, 2.6, 3.4, 4.3, 5.0, 5.5, 6.0, 6.5, 6.7};
    return snr[rate_index] + (shortframes ? 0.4 : 0.0) + 0.9;
}

int rate_of_modcod(int m) {   // EN 302 307 table 12 -> dvbs2_code_rate_t numbering (dvbs2/dvbs2.h:11-25)
    static const int r[29] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 4, 5, 6, 8, 10, 11, 5, 6, 7, 8, 10, 11, 6, 7, 8, 10, 11};
    return r[m];
}

void make_pool(Stream& s, int frames, uint64_t seed) {
    int kldpc, t, bits, plsyms, links;
    if (dvbs2fec_modcod_info(s.modcod, s.shortframes, 0, &s.N, &kldpc, &s.kbch, &t, &bits, &plsyms, &links) != 0) {
        fprintf(stderr, "modcod %d %s not available\n", s.modcod, s.shortframes ? "short" : "normal");
        exit(2);
    }
    std::mt19937_64 rng(seed);
    std::normal_distribution<float> gauss(0.f, 1.f);
    const double esn0 = esn0_for(rate_of_modcod(s.modcod), s.shortframes);
    const double a = 1.0 / std::sqrt(2.0), sigma2 = 1.0 / (2.0 * std::pow(10.0, esn0 / 10.0));
    const float sigma = (float)std::sqrt(sigma2), gain = (float)(4.0 * 2.0 * a / sigma2);
    std::vector<uint8_t> bb(s.kbch / 8), code(s.N);
    s.pool.resize((size_t)frames * s.N);
    s.pool_frames = frames;
    for (int f = 0; f < frames; ++f) {
        for (auto& b : bb) b = (uint8_t)rng();
        dvbs2fec_encode_fecframe(s.modcod, s.shortframes, bb.data(), code.data());
        int8_t* out = &s.pool[(size_t)f * s.N];
        for (int i = 0; i < s.N; ++i) {
            float y = (code[i] ? -(float)a : (float)a) + sigma * gauss(rng);
            float l = std::nearbyint(gain * y);
            out[i] = (int8_t)std::max(-127.f, std::min(127.f, l));
        }
    }
}

// One transponder: submit `total` frames as fast as the queue accepts them, keeping at most `window` in flight.
void run_stream(Stream& s, long total, int window, std::atomic<int>& go) {
    std::vector<Clock::time_point> t_submit(total);
    std::vector<uint8_t> bb((size_t)64 * (s.kbch / 8));
    std::vector<dvbs2fec_result> res(64);
    s.lat_us.clear();
    s.lat_us.reserve(total);
    s.frames = s.failed = s.iters = 0;
    while (!go.load()) std::this_thread::yield();
    long submitted = 0, collected = 0;
    while (collected < total) {
        bool progressed = false;
        while (submitted < total && submitted - collected < window) {
            t_submit[submitted] = Clock::now();
            int rc = dvbs2fec_submit_llr(s.h, &s.pool[(size_t)(submitted % s.pool_frames) * s.N], (uint64_t)submitted);
            if (rc == DVBS2FEC_EAGAIN) break;
            if (rc != 0) {
                fprintf(stderr, "submit: %s\n", dvbs2fec_last_error());
                exit(3);
            }
            ++submitted;
            progressed = true;
        }
        if (submitted == total) dvbs2fec_flush(s.h);
        int n = dvbs2fec_collect(s.h, bb.data(), res.data(), 64, progressed ? 0 : 200);
        if (n < 0) {
            fprintf(stderr, "collect: %s\n", dvbs2fec_last_error());
            exit(3);
        }
        auto now = Clock::now();
        for (int k = 0; k < n; ++k) {
            if ((long)res[k].tag != collected) {   // delivery order is part of the contract
                fprintf(stderr, "out-of-order delivery: tag %llu at position %ld\n", (unsigned long long)res[k].tag, collected);
                exit(4);
            }
            s.lat_us.push_back(std::chrono::duration<float, std::micro>(now - t_submit[collected]).count());
            s.failed += res[k].bch_corr < 0;
            s.iters += res[k].ldpc_iters < 0 ? 25 : res[k].ldpc_iters;
            ++collected;
        }
        s.frames = collected;
    }
}

float pct(std::vector<float>& v, double p) {
    if (v.empty()) return 0.f;
    size_t k = std::min(v.size() - 1, (size_t)(p * (v.size() - 1) + 0.5));
    std::nth_element(v.begin(), v.begin() + k, v.end());
    return v[k];
}

}  // namespace

int main(int argc, char** argv) {
    int gpus = 1, frames_per_stream = 2048, pool_frames = 24;
    std::vector<int> batches = {1, 4, 16, 64, 256, 1024, 4096};
    const char* out_path = nullptr;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--frames" && i + 1 < argc) frames_per_stream = atoi(argv[++i]);
        else if (a == "--pool" && i + 1 < argc) pool_frames = atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out_path = argv[++i];
        else if (a == "--batches" && i + 1 < argc) {
            batches.clear();
            for (char* tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) batches.push_back(atoi(tok));
        } else {
            fprintf(stderr, "usage: mixed_stream [--gpus N] [--frames F] [--pool P] [--batches 1,4,..] [--out file.json]\n");
            return 1;
        }
    }
    // 8 transponders: MODCODs drawn from {4,6,9,11,12,13,18,24,28} normal and {4,6} short (fixed draw)
    const int plan[8][2] = {{4, 0}, {6, 0}, {9, 0}, {11, 0}, {13, 0}, {24, 0}, {4, 1}, {6, 1}};
    std::vector<Stream> streams(8);
    for (int t = 0; t < 8; ++t) {
        streams[t].modcod = plan[t][0];
        streams[t].shortframes = plan[t][1];
        make_pool(streams[t], pool_frames, 1000 + t);
    }
    std::string json = "{\"workload\": \"8 transponders, MODCOD {4,6,9,11,13,24} normal + {4,6} short, one handle + host thread each, "
                       "dvbs2fec_submit_llr -> dvbs2fec_collect, pageable host LLRs\", \"gpus\": " + std::to_string(gpus) +
                       ", \"frames_per_transponder\": " + std::to_string(frames_per_stream) + ", \"rows\": [";
    bool first_row = true;
    for (int batch : batches) {
        for (auto& s : streams) {
            dvbs2fec_config cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.n_devices = gpus;
            for (int g = 0; g < gpus; ++g) cfg.devices[g] = g;
            cfg.max_batch = batch;
            cfg.max_latency_us = 2000;
            cfg.max_trials = 25;
            if (dvbs2fec_create(&cfg, &s.h) != 0 || dvbs2fec_set_modcod(s.h, s.modcod, s.shortframes, 0, 25) != 0) {
                fprintf(stderr, "create/set_modcod: %s\n", dvbs2fec_last_error());
                return 2;
            }
        }
        const long total = std::max<long>({64L, 5L * batch * gpus, (long)std::min(frames_per_stream, batch * 64)});
        const int window = std::max(2 * batch, 8);
        for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up (lazy allocations, first launches)
            std::atomic<int> go{0};
            std::vector<std::thread> th;
            const long n = pass == 0 ? std::max(5 * batch * gpus, 32) : total;   // warm-up touches every staging batch and both slots
            for (auto& s : streams) th.emplace_back(run_stream, std::ref(s), n, window, std::ref(go));
            auto t0 = Clock::now();
            go.store(1);
            for (auto& t : th) t.join();
            double sec = std::chrono::duration<double>(Clock::now() - t0).count();
            if (pass == 0) continue;
            std::vector<float> all;
            long frames = 0, failed = 0, iters = 0;
            double bits = 0;
            for (auto& s : streams) {
                all.insert(all.end(), s.lat_us.begin(), s.lat_us.end());
                frames += s.frames;
                failed += s.failed;
                iters += s.iters;
                bits += (double)s.frames * s.kbch;
            }
            float p50 = pct(all, 0.50), p99 = pct(all, 0.99), pmax = *std::max_element(all.begin(), all.end());
            char row[512];
            snprintf(row, sizeof row,
                     "%s{\"max_batch\": %d, \"frames\": %ld, \"seconds\": %.4f, \"frames_per_s\": %.0f, \"info_gbit_s\": %.3f, "
                     "\"latency_ms\": {\"p50\": %.3f, \"p99\": %.3f, \"max\": %.3f}, \"mean_ldpc_iters\": %.2f, \"failed\": %ld}",
                     first_row ? "" : ", ", batch, frames, sec, frames / sec, bits / sec / 1e9, p50 / 1e3, p99 / 1e3, pmax / 1e3,
                     (double)iters / frames, failed);
            first_row = false;
            json += row;
            printf("batch %5d: %7ld frames  %9.0f frames/s  %7.3f Gbit/s  latency p50 %.3f ms  p99 %.3f ms  max %.3f ms  iters %.2f  failed %ld\n",
                   batch, frames, frames / sec, bits / sec / 1e9, p50 / 1e3, p99 / 1e3, pmax / 1e3, (double)iters / frames, failed);
            fflush(stdout);
        }
        for (auto& s : streams) {
            dvbs2fec_destroy(s.h);
            s.h = nullptr;
        }
    }
    json += "]}\n";
    if (out_path) {
        FILE* f = fopen(out_path, "w");
        if (f) {
            fputs(json.c_str(), f);
            fclose(f);
        }
    }
    return 0;
}
