// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for SDR++ core's <utils/flog.h>;
// constellation.cpp:148 calls flog::error(const char*) on an undefined constellation type.
#pragma once
#include <cstdio>
namespace flog {
    inline void error(const char* msg) { fprintf(stderr, "[flog] %s\n", msg); }
    template <typename... A>
    inline void info(const char*, A...) {}   // dvbs2_pl_sync.cpp:29,68 log the pilot frame size
}
