// TEST INFRASTRUCTURE.  Stand-in for SDR++ core's <dsp/stream.h>, which dvbs2/bbframe_ts_parser.h includes
// but does not use; nothing is needed from it to compile the parser.
#pragma once
