// TEST INFRASTRUCTURE ONLY (oracle/_ref build) — not part of the product.
//
// Minimal stand-in for doctest 2.4.12 (absent from this image, no network) so that the
// reference's hot-path sources, which carry their doctest cases inline
// (src/compressor.cpp:300-406, src/calc-loss.cpp:68-86), compile unmodified.  Unlike a no-op
// shim the cases are registered and runnable: wcref_run_doctests() in ref_harness.cpp executes
// them, which pins that this build of the reference passes the reference's own tests.
//
// Simplification vs real doctest: SUBCASE bodies all run in one pass of the enclosing case
// (real doctest re-enters the case once per subcase).  The reference's cases only declare
// scoped locals inside subcases, so this is equivalent for them.
#pragma once

#include <cstdio>
#include <vector>

namespace wc_doctest {
struct Case {
    const char* name;
    void (*fn)();
};
inline std::vector<Case>& registry() {
    static std::vector<Case> r;
    return r;
}
inline int& failures() {
    static int f = 0;
    return f;
}
inline int& assertions() {
    static int a = 0;
    return a;
}
struct Registrar {
    Registrar(const char* name, void (*fn)()) { registry().push_back({ name, fn }); }
};
inline void check(bool ok, const char* expr, const char* file, int line) {
    ++assertions();
    if (!ok) {
        ++failures();
        std::fprintf(stderr, "[wc_doctest] FAILED %s:%d: %s\n", file, line, expr);
    }
}
} // namespace wc_doctest

#define WC_DT_CAT2(a, b) a##b
#define WC_DT_CAT(a, b) WC_DT_CAT2(a, b)

#define WC_DT_CASE_IMPL(fn, name)                                                  \
    static void fn();                                                              \
    static ::wc_doctest::Registrar WC_DT_CAT(fn, _reg)(name, &fn);                 \
    static void fn()

#define TEST_CASE(name) WC_DT_CASE_IMPL(WC_DT_CAT(wc_dt_case_, __COUNTER__), name)
#define SUBCASE(name) if (true)
#define REQUIRE(...) ::wc_doctest::check(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK(...) REQUIRE(__VA_ARGS__)
#define REQUIRE_FALSE(...) ::wc_doctest::check(!static_cast<bool>(__VA_ARGS__), "!(" #__VA_ARGS__ ")", __FILE__, __LINE__)
#define CHECK_FALSE(...) REQUIRE_FALSE(__VA_ARGS__)
#define INFO(...) ((void)0)
#define CAPTURE(...) ((void)0)
