// TEST INFRASTRUCTURE ONLY (oracle/_ref build) — not part of the product.
//
// Timing / stubbing wrappers around the four liblzma entry points the reference calls
// (src/compressor.cpp:260-285, src/decompressor.cpp:188-220).  See oracle/shim/lzma.h.
#define WCREF_NO_LZMA_INTERCEPT
#include "shim/lzma.h"

#include <atomic>
#include <chrono>

namespace {
std::atomic<int>      g_mode { 0 };
std::atomic<uint64_t> g_lzma_ns { 0 };

// Marks streams created by the stubbed encoder so lzma_code/lzma_end know not to call liblzma.
const uint64_t STUB_TAG = 0x57435245465F5354ull; // "WCREF_ST"

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Timer() {
        auto dt = std::chrono::steady_clock::now() - t0;
        g_lzma_ns.fetch_add(
            (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(dt).count(),
            std::memory_order_relaxed);
    }
};
} // namespace

extern "C" {

void wcref_set_lzma_mode(int mode) { g_mode.store(mode); }
int  wcref_get_lzma_mode() { return g_mode.load(); }
void wcref_reset_lzma_seconds() { g_lzma_ns.store(0); }
double wcref_lzma_seconds() { return (double)g_lzma_ns.load() * 1e-9; }

lzma_ret wcref_lzma_easy_encoder(lzma_stream* strm, uint32_t preset, lzma_check check) {
    if (g_mode.load(std::memory_order_relaxed) == 1) {
        strm->internal      = nullptr;
        strm->reserved_int1 = STUB_TAG;
        return LZMA_OK;
    }
    Timer t;
    return lzma_easy_encoder(strm, preset, check);
}

lzma_ret wcref_lzma_stream_decoder(lzma_stream* strm, uint64_t memlimit, uint32_t flags) {
    Timer t;
    return lzma_stream_decoder(strm, memlimit, flags);
}

lzma_ret wcref_lzma_code(lzma_stream* strm, lzma_action action) {
    if (strm->internal == nullptr && strm->reserved_int1 == STUB_TAG) {
        // stubbed encoder: consume everything, produce nothing
        strm->total_in += strm->avail_in;
        strm->next_in += strm->avail_in;
        strm->avail_in = 0;
        return action == LZMA_FINISH ? LZMA_STREAM_END : LZMA_OK;
    }
    Timer t;
    return lzma_code(strm, action);
}

void wcref_lzma_end(lzma_stream* strm) {
    if (strm->internal == nullptr && strm->reserved_int1 == STUB_TAG) {
        strm->reserved_int1 = 0;
        return;
    }
    Timer t;
    lzma_end(strm);
}

} // extern "C"
