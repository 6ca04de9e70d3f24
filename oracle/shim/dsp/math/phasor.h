// TEST INFRASTRUCTURE (oracle/_ref build only): SDR++ core's dsp::math::phasor -- the unit vector at angle x.  Unpinned.
#pragma once
#include <math.h>
#include "../types.h"
namespace dsp::math {
    inline complex_t phasor(float x) { return complex_t{cosf(x), sinf(x)}; }
}
