// TEST INFRASTRUCTURE (oracle/_ref build only): common/utils.cpp:52-53 includes the nng headers; every use of them there is commented out.
#pragma once
