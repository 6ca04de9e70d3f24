// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for SDR++ core's <utils/flog.h>;
// constellation.cpp:148 calls flog::error(const char*) on an undefined constellation type.
#pragma once
#include <cstdio>
namespace flog {
    inline void error(const char* msg) { fprintf(stderr, "[flog] %s\n", msg); }
}
