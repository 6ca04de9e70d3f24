// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for SDR++ core's <dsp/processor.h>, enough to compile the
// reference's S2PLSyncBlock / S2PLHDRDemod (dvbs2/dvbs2_pl_sync.h, dvbs2_plhdr_demod.h) and to call their process()
// directly.  The block plumbing (streams, worker thread, tempStop/tempStart) is inert; STREAM_BUFFER_SIZE is SDR++
// core's value (1 000 000 samples, core/src/dsp/stream.h, unverified here -- it only sizes S2PLSyncBlock::in_buffer).
// SDR++ core pulls VOLK in through its DSP headers; the two VOLK kernels dvbs2_pl_sync.cpp uses get their generic
// (portable C) definitions in shim/volk/volk.h.  Parity of all of this with the real headers is unpinned.
#pragma once
#include <cassert>
#include <cstring>
#include <memory>
#include <mutex>
#include "types.h"
#include <volk/volk.h>
#ifndef STREAM_BUFFER_SIZE
#define STREAM_BUFFER_SIZE 1000000
#endif
#ifndef FL_M_PI
#define FL_M_PI 3.1415926535f
#endif
namespace dsp {
    template <class T>
    struct stream {
        T* readBuf = nullptr;
        T* writeBuf = nullptr;
        int read() { return -1; }
        void flush() {}
        bool swap(int) { return true; }
    };
    template <class I, class O>
    class Processor {
    public:
        virtual ~Processor() {}
        virtual void init(stream<I>* in) { _in = in; _block_init = true; }
        void tempStop() {}
        void tempStart() {}
        stream<O> out;
    protected:
        stream<I>* _in = nullptr;
        bool _block_init = false;
        std::recursive_mutex ctrlMtx;
    };
}
