// TEST INFRASTRUCTURE (oracle/_ref build only): SDR++ core's dsp::math::step -- +1 for x > 0, else -1.  Unpinned.
#pragma once
namespace dsp::math {
    template <class T>
    inline T step(T x) { return (x > 0.0) ? 1.0 : -1.0; }
}
