// TEST INFRASTRUCTURE ONLY (oracle/_ref build) — not part of the product.
//
// <lzma.h> for the reference build: the image ships liblzma.so.5 (5.4.5) without its header, so
// the hand-declared ABI subset from the product's host shim is reused, and lzma_code is routed
// through a timing/stubbing wrapper so the CPU baseline can report the reference's numeric core
// separately from its LZMA stage ("numeric core only" vs "full").
//
//   mode 0 (default): real liblzma; wall time spent inside lzma_* is accumulated (per thread,
//                     summed on read).
//   mode 1          : encoder stubbed — lzma_code(FINISH) on an encoder stream reports
//                     LZMA_STREAM_END with zero bytes produced.  The resulting .xz files are
//                     empty and NOT decodable; used only to time compress() minus LZMA.
#pragma once

#include "../../wavelet-compression_b200/host/wc_lzma_abi.h"

#ifdef __cplusplus
extern "C" {
#endif
lzma_ret wcref_lzma_easy_encoder(lzma_stream* strm, uint32_t preset, lzma_check check);
lzma_ret wcref_lzma_stream_decoder(lzma_stream* strm, uint64_t memlimit, uint32_t flags);
lzma_ret wcref_lzma_code(lzma_stream* strm, lzma_action action);
void     wcref_lzma_end(lzma_stream* strm);
#ifdef __cplusplus
}
#endif

#ifndef WCREF_NO_LZMA_INTERCEPT
#define lzma_easy_encoder wcref_lzma_easy_encoder
#define lzma_stream_decoder wcref_lzma_stream_decoder
#define lzma_code wcref_lzma_code
#define lzma_end wcref_lzma_end
#endif
