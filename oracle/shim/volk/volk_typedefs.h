// TEST INFRASTRUCTURE (oracle/_ref build only): cc_encoder.cpp includes it and uses nothing from it.
#pragma once
