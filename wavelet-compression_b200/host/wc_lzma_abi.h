// Hand-declared subset of the liblzma 5.x C ABI (xz utils), for images that ship the runtime
// library (liblzma.so.5) but not its development header.  Only what the .xz container step of
// the reference needs is declared: the easy encoder, the stream decoder, lzma_code, lzma_end
// (reference call sites: src/compressor.cpp:260-285, src/decompressor.cpp:188-220).
//
// The lzma_stream layout below is liblzma's documented stable public struct (unchanged since
// 5.0; the 5.3+ rename of reserved_int1 to seek_pos does not alter the layout).  When the real
// <lzma.h> is available include that instead and define WC_HAVE_SYSTEM_LZMA_H.
#pragma once

#ifdef WC_HAVE_SYSTEM_LZMA_H
#include <lzma.h>
#else

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    LZMA_OK                = 0,
    LZMA_STREAM_END        = 1,
    LZMA_NO_CHECK          = 2,
    LZMA_UNSUPPORTED_CHECK = 3,
    LZMA_GET_CHECK         = 4,
    LZMA_MEM_ERROR         = 5,
    LZMA_MEMLIMIT_ERROR    = 6,
    LZMA_FORMAT_ERROR      = 7,
    LZMA_OPTIONS_ERROR     = 8,
    LZMA_DATA_ERROR        = 9,
    LZMA_BUF_ERROR         = 10,
    LZMA_PROG_ERROR        = 11
} lzma_ret;

typedef enum {
    LZMA_RUN          = 0,
    LZMA_SYNC_FLUSH   = 1,
    LZMA_FULL_FLUSH   = 2,
    LZMA_FINISH       = 3,
    LZMA_FULL_BARRIER = 4
} lzma_action;

typedef enum {
    LZMA_CHECK_NONE   = 0,
    LZMA_CHECK_CRC32  = 1,
    LZMA_CHECK_CRC64  = 4,
    LZMA_CHECK_SHA256 = 10
} lzma_check;

typedef enum { LZMA_RESERVED_ENUM = 0 } lzma_reserved_enum;

typedef struct lzma_allocator_s lzma_allocator;
typedef struct lzma_internal_s  lzma_internal;

typedef struct {
    const uint8_t*        next_in;
    size_t                avail_in;
    uint64_t              total_in;
    uint8_t*              next_out;
    size_t                avail_out;
    uint64_t              total_out;
    const lzma_allocator* allocator;
    lzma_internal*        internal;
    void*                 reserved_ptr1;
    void*                 reserved_ptr2;
    void*                 reserved_ptr3;
    void*                 reserved_ptr4;
    uint64_t              reserved_int1;
    uint64_t              reserved_int2;
    size_t                reserved_int3;
    size_t                reserved_int4;
    lzma_reserved_enum    reserved_enum1;
    lzma_reserved_enum    reserved_enum2;
} lzma_stream;

#define LZMA_STREAM_INIT                                                              \
    { NULL, 0, 0, NULL, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL, 0, 0, 0, 0,         \
      LZMA_RESERVED_ENUM, LZMA_RESERVED_ENUM }

#define LZMA_TELL_NO_CHECK          UINT32_C(0x01)
#define LZMA_TELL_UNSUPPORTED_CHECK UINT32_C(0x02)
#define LZMA_TELL_ANY_CHECK         UINT32_C(0x04)
#define LZMA_CONCATENATED           UINT32_C(0x08)

lzma_ret lzma_easy_encoder(lzma_stream* strm, uint32_t preset, lzma_check check);
lzma_ret lzma_stream_decoder(lzma_stream* strm, uint64_t memlimit, uint32_t flags);
lzma_ret lzma_code(lzma_stream* strm, lzma_action action);
void     lzma_end(lzma_stream* strm);

#ifdef __cplusplus
}
#endif

#endif // WC_HAVE_SYSTEM_LZMA_H
