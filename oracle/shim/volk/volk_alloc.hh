// TEST INFRASTRUCTURE (oracle/_ref build only): volk::vector is std::vector with VOLK's aligned allocator; alignment
// does not change any value the reference computes.  (The real header pulls in the
// standard headers the reference's viterbi sources rely on without including them.)
#pragma once
#include <cstdint>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace volk {
    template <class T>
    using vector = std::vector<T>;
}
