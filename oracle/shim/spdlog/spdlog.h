// TEST INFRASTRUCTURE ONLY (oracle/_ref build) — not part of the product.
//
// Stand-in for spdlog 1.12.0 (absent from this image).  The reference's hot path only logs on
// fatal LZMA / file errors before exit(EXIT_FAILURE) (src/compressor.cpp:263-266,280-283;
// src/decompressor.cpp:170-231); the messages are printed to stderr verbatim (no fmt
// substitution) so a failure inside the reference build is still visible.
//
// The real spdlog transitively provides <cstring>/<cstdint>/<string>, which
// src/decompressor.cpp:43 (std::memcpy) relies on — reproduced here.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace spdlog {
namespace wc_detail {
inline void emit(const char* level, const char* msg) {
    std::fprintf(stderr, "[ref:%s] %s\n", level, msg);
}
inline const char* cstr(const char* s) { return s; }
inline const char* cstr(const std::string& s) { return s.c_str(); }
} // namespace wc_detail

template <class Fmt, class... Args>
inline void error(const Fmt& f, Args&&...) { wc_detail::emit("error", wc_detail::cstr(f)); }
template <class Fmt, class... Args>
inline void warn(const Fmt& f, Args&&...) { wc_detail::emit("warn", wc_detail::cstr(f)); }
template <class Fmt, class... Args>
inline void info(const Fmt&, Args&&...) { }
template <class Fmt, class... Args>
inline void debug(const Fmt&, Args&&...) { }
} // namespace spdlog
