// wc_dropin.hpp — C++ drop-in for the reference's three hot-path entry points, on top of the C ABI.
//
// Include it AFTER the reference's own "box-structs.h" (Grid3D / Box3D / multiBox3D /
// CompressedWavelet, src/grid.h + src/box-structs.h) and link libwcgpu.so + liblzma: the functions in
// namespace wcgpu have exactly the signatures src/modes.cpp calls,
//
//   compress(multiBox3D&, std::vector<int>, double, int, int, int, std::string)   src/compressor.h:9-15
//   decompress(std::string, int, int, int, int) -> Box3D                          src/decompressor.h:6-10
//   calc_rmse_per_box(const multiBox3D&, const multiBox3D&, int)                  src/calc-loss.h:6-8
//   inverse_wavelet_decompose(std::vector<float>, int, int, int) -> Box3D          src/decompressor.h:18
//   deserialize_compressed_wavelet(const std::string&)                              src/decompressor.h:14
//
// so a host re-points its loops with `using namespace wcgpu;` or by swapping the two includes
// (INTEGRATION.md).  What stays on the host is what the reference keeps on the host: the 20-byte
// header, the .xz container (xz preset 6, CRC64, one lzma_code(FINISH)) and the file names.  Errors
// keep the reference's behaviour: a message and exit(EXIT_FAILURE) (src/compressor.cpp:263-266,
// src/decompressor.cpp:170-231) — the C ABI underneath never exits by itself.
//
// Beyond the per-box calls there are batched entry points that a ported modes.cpp should use (the per-box calls
// are correct but launch-bound):
//   compress_all    one plan for the whole run; the chunk callback of wc_plan_compress_to_host_chunked feeds a pool
//                   of LZMA writer threads WHILE later chunks are still on the GPU (src/modes.cpp:100-103)
//   decompress_all  host threads xz-decode the files, the pair bytes form one dense stream for a decode plan
//                   (wc_dplan_*), one pipelined GPU pass (src/modes.cpp:151-166)
//   estimate_all    the `-estimate` body (src/modes.cpp:236-291) on ONE plan: boxes cross the bus once, compress ->
//                   decompress -> RMSE -> min/max -> need32 stay in HBM, only the packed pairs come back for the
//                   LZMA size estimate (overlapped the same way)
// A threaded multi-GPU host calls wcgpu::set_device(g) once per thread; contexts are per (thread, device).
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <limits>
#include <fstream>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/wcgpu.h"
#include "wc_lzma_abi.h"

namespace wcgpu {

namespace detail {

[[noreturn]] inline void die(const std::string& what) {
    std::fprintf(stderr, "[wcgpu] %s\n", what.c_str());
    std::exit(EXIT_FAILURE);
}

inline void check(int status, wc_ctx* ctx, const char* what) {
    if (status != WC_OK)
        die(std::string(what) + ": " + wc_strerror(status) + (ctx ? std::string(" — ") + wc_last_error(ctx) : ""));
}

// One context per (host thread, device), created on first use and destroyed with the thread.  A wc_ctx is not
// thread-safe (include/wcgpu.h), so contexts are never shared between threads: a threaded multi-GPU host
// (one thread per GPU) calls wcgpu::set_device(g) in each thread and then uses the same
// free functions as the single-GPU host.  The default device is 0.
struct ThreadContexts {
    std::vector<wc_ctx*> by_device;
    int                  current = 0;
    ~ThreadContexts() {
        for (wc_ctx* c : by_device)
            if (c) wc_destroy(c);
    }
};
inline ThreadContexts& thread_contexts() {
    static thread_local ThreadContexts tc;
    return tc;
}
inline wc_ctx* context(int device = -1) {
    ThreadContexts& tc = thread_contexts();
    if (device < 0) device = tc.current;
    if ((size_t)device >= tc.by_device.size()) tc.by_device.resize((size_t)device + 1, nullptr);
    if (!tc.by_device[device]) check(wc_create(&tc.by_device[device], device), nullptr, "wc_create");
    return tc.by_device[device];
}

// Validated view of a decoded (xz-decompressed) unit: the 20-byte header of src/compressor.cpp:59-71 and the
// K pairs behind it.  Rejects what the reference would read out of bounds on: a short buffer, negative
// dims / counts, K > ncoef, ncoef != nx*ny*nz, fewer than 8K payload bytes (truncated or hostile file).
inline wc_packed parse_unit(const std::string& raw, const std::string& what) {
    if (raw.size() < 20) die("Deserialization failed: short file " + what);
    int32_t h[5];
    std::memcpy(h, raw.data(), 20);
    if (h[0] < 0 || h[1] < 0 || h[2] < 0 || h[3] < 0 || h[4] < 0) die("Deserialization failed: negative header field " + what);
    if ((long long)h[0] * h[1] * h[2] != (long long)h[3]) die("Deserialization failed: shape / coefficient count mismatch " + what);
    if (h[4] > h[3]) die("Deserialization failed: more pairs than coefficients " + what);
    if (raw.size() - 20 < 8 * (size_t)h[4]) die("Deserialization failed: truncated file " + what);
    wc_packed p;
    p.shape[0] = h[0]; p.shape[1] = h[1]; p.shape[2] = h[2];
    p.ncoef = h[3]; p.npairs = h[4]; p.flags = 0;
    // the pairs sit 4-byte aligned inside the decoded string (offset 20): wc_pair has 4-byte alignment
    p.pairs = reinterpret_cast<wc_pair*>(const_cast<char*>(raw.data()) + 20);
    return p;
}

template <class Box>
inline const float* box_data(const Box& b) {
    return b.data_size() ? &b(0, 0, 0) : nullptr;   // Grid3D storage is one contiguous x-fastest vector
}

// serialize_compressed_wavelet, src/compressor.cpp:55-80: header from the C ABI + the pairs as they are
inline std::string serialize(const wc_packed& p) {
    std::string buf(20 + 8 * (size_t)p.npairs, '\0');
    wc_serialize_header(&p, reinterpret_cast<uint8_t*>(&buf[0]));
    if (p.npairs) std::memcpy(&buf[20], p.pairs, 8 * (size_t)p.npairs);
    return buf;
}

// lzma_easy_encoder(6, CRC64) + one lzma_code(FINISH), src/compressor.cpp:260-285
inline std::vector<uint8_t> xz_encode(const std::string& serialized) {
    lzma_stream strm = LZMA_STREAM_INIT;
    if (lzma_easy_encoder(&strm, 6, LZMA_CHECK_CRC64) != LZMA_OK) die("Failed to initialize LZMA encoder");
    std::vector<uint8_t> out((size_t)(serialized.size() * 1.1) + 128);
    strm.next_in   = reinterpret_cast<const uint8_t*>(serialized.data());
    strm.avail_in  = serialized.size();
    strm.next_out  = out.data();
    strm.avail_out = out.size();
    if (lzma_code(&strm, LZMA_FINISH) != LZMA_STREAM_END) die("LZMA compression failed");
    out.resize(out.size() - strm.avail_out);
    lzma_end(&strm);
    return out;
}

// lzma_stream_decoder(LZMA_CONCATENATED) with a doubling buffer, src/decompressor.cpp:188-220
inline std::string xz_decode_file(const std::string& filename) {
    std::error_code ec;
    auto size = std::filesystem::file_size(filename, ec);
    if (ec) die("Error getting file size: " + ec.message() + " " + filename);
    std::ifstream file(filename, std::ios::binary);
    if (!file) die("Failed to open file: " + filename);
    std::vector<uint8_t> in(size);
    if (size && !file.read(reinterpret_cast<char*>(in.data()), (std::streamsize)size)) die("Failed to read file: " + filename);
    lzma_stream strm = LZMA_STREAM_INIT;
    if (lzma_stream_decoder(&strm, UINT64_MAX, LZMA_CONCATENATED) != LZMA_OK) die("Failed to initialize LZMA decoder.");
    strm.next_in  = in.data();
    strm.avail_in = in.size();
    std::vector<uint8_t> out(4096);
    strm.next_out  = out.data();
    strm.avail_out = out.size();
    for (;;) {
        lzma_ret ret = lzma_code(&strm, LZMA_FINISH);
        if (ret == LZMA_STREAM_END) break;
        if (ret != LZMA_OK) die("LZMA decompression failed with code: " + std::to_string((int)ret));
        size_t old = out.size();
        out.resize(old * 2);
        strm.next_out  = out.data() + old;
        strm.avail_out = old;
    }
    size_t n = out.size() - strm.avail_out;
    lzma_end(&strm);
    return std::string(reinterpret_cast<char*>(out.data()), n);
}

inline std::string unit_path(const std::string& dir, int t, int lev, int comp, int box) {
    return (std::filesystem::path(dir) / ("compressed-wavelet-" + std::to_string(t) + "-" + std::to_string(lev) + "-" +
                                           std::to_string(comp) + "-" + std::to_string(box) + ".xz")).string();
}

inline void write_unit_file(const std::string& path, const wc_packed& p) {
    std::ofstream file(path, std::ios::binary);
    if (!file.is_open()) return;   // the reference silently skips a file it cannot open (src/compressor.cpp:256-257)
    std::vector<uint8_t> xz = xz_encode(serialize(p));
    file.write(reinterpret_cast<const char*>(xz.data()), (std::streamsize)xz.size());
}

inline CompressedWavelet to_compressed_wavelet(const wc_packed& p) {
    CompressedWavelet cw;
    cw.shape       = { p.shape[0], p.shape[1], p.shape[2] };
    cw.coeff_shape = { p.ncoef };
    cw.rle_encoded.resize((size_t)p.npairs);
    bool need32 = false;
    for (int i = 0; i < p.npairs; ++i) {
        cw.rle_encoded[i] = { p.pairs[i].run, p.pairs[i].val };
        double a = p.pairs[i].val < 0 ? -(double)p.pairs[i].val : (double)p.pairs[i].val;
        if (a > INT16_MAX) need32 = true;   // src/compressor.cpp:229 (computed, never serialized)
    }
    cw.need32 = need32;
    return cw;
}

// A small pool of worker threads for the LZMA stage: jobs are queued from the GPU thread (chunk callback) and run
// concurrently with the rest of the GPU batch; wait() drains the queue.
class WorkerPool {
    std::vector<std::thread>          workers_;
    std::queue<std::function<void()>> jobs_;
    std::mutex                        mu_;
    std::condition_variable           cv_, idle_;
    size_t                            active_ = 0;
    bool                              stop_   = false;

public:
    explicit WorkerPool(unsigned threads) {
        if (threads == 0) threads = std::thread::hardware_concurrency();
        if (threads == 0) threads = 1;
        for (unsigned t = 0; t < threads; ++t)
            workers_.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
                        if (jobs_.empty()) return;
                        job = std::move(jobs_.front());
                        jobs_.pop();
                        ++active_;
                    }
                    job();
                    {
                        std::lock_guard<std::mutex> lk(mu_);
                        --active_;
                    }
                    idle_.notify_all();
                }
            });
    }
    void submit(std::function<void()> job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push(std::move(job));
        }
        cv_.notify_one();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        idle_.wait(lk, [this] { return jobs_.empty() && active_ == 0; });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
};

template <class F>
inline void parallel_for(size_t n, unsigned threads, F f) {
    if (threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next { 0 };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&] {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i);
        });
    for (auto& th : pool) th.join();
}

} // namespace detail

// The device the calling thread's wcgpu:: calls run on (default 0).  One host thread per GPU.
inline void set_device(int device) { detail::thread_contexts().current = device; }

// ---- compress(): src/compressor.cpp:192-297 ----------------------------------------------------------
inline std::vector<CompressedWavelet> compress(multiBox3D& box, std::vector<int> components, double keep,
                                               int time, int level, int box_index, std::string compressed_dir) {
    wc_ctx* ctx = detail::context();
    const int n = (int)components.size();
    std::vector<wc_box_desc> in(n);
    std::vector<wc_packed>   out(n);
    for (int c = 0; c < n; ++c)
        in[c] = { detail::box_data(box[c]), WC_F32, (int32_t)box[c].width(), (int32_t)box[c].height(), (int32_t)box[c].depth() };
    detail::check(wc_compress_batch(ctx, in.data(), n, WC_HOST, keep, WC_THRESH_PER_UNIT, out.data(), WC_HOST), ctx,
                  "wc_compress_batch");
    std::vector<CompressedWavelet> ret;
    for (int c = 0; c < n; ++c) {
        detail::write_unit_file(detail::unit_path(compressed_dir, time, level, components[c], box_index), out[c]);
        ret.push_back(detail::to_compressed_wavelet(out[c]));
    }
    return ret;
}

// ---- deserialize_compressed_wavelet(): src/decompressor.cpp:35-74 (host only) --------------------------
inline CompressedWavelet deserialize_compressed_wavelet(const std::string& data) {
    CompressedWavelet cw;
    int32_t h[5];
    std::memcpy(h, data.data(), 20);
    cw.shape       = { h[0], h[1], h[2] };
    cw.coeff_shape = { h[3] };
    cw.rle_encoded.resize((size_t)h[4]);
    for (int i = 0; i < h[4]; ++i) {
        int   run;
        float val;
        std::memcpy(&run, data.data() + 20 + 8 * (size_t)i, 4);
        std::memcpy(&val, data.data() + 24 + 8 * (size_t)i, 4);
        cw.rle_encoded[i] = { run, val };
    }
    cw.need32 = false;
    return cw;
}

// ---- decompress(): src/decompressor.cpp:238-255 -----------------------------------------------------------
inline Box3D decompress(std::string file_path, int /*time*/, int /*level*/, int /*component*/, int /*box_idx*/) {
    wc_ctx* ctx = detail::context();
    std::string raw = detail::xz_decode_file(file_path);
    const wc_packed p = detail::parse_unit(raw, file_path);
    const int32_t* h = p.shape;
    Box3D box((size_t)h[0], (size_t)h[1], (size_t)h[2]);
    wc_box_out o { box.data_size() ? &box(0, 0, 0) : nullptr, WC_F32, h[0], h[1], h[2] };
    detail::check(wc_decompress_batch(ctx, &p, 1, WC_HOST, &o, WC_HOST), ctx, "wc_decompress_batch");
    return box;
}

// ---- inverse_wavelet_decompose(): src/decompressor.cpp:79-159 ---------------------------------------------
inline Box3D inverse_wavelet_decompose(std::vector<float> flat, int x, int y, int z) {
    wc_ctx* ctx = detail::context();
    Box3D box((size_t)x, (size_t)y, (size_t)z);
    detail::check(wc_haar_inverse(ctx, flat.data(), x, y, z, WC_HOST, box.data_size() ? &box(0, 0, 0) : nullptr), ctx,
                  "wc_haar_inverse");
    return box;
}

// wavelet_decompose is `static` in the reference (src/compressor.cpp:85); exposed here for tests.
inline std::vector<float> wavelet_decompose(const Box3D& box) {
    wc_ctx* ctx = detail::context();
    std::vector<float> flat(box.data_size());
    wc_box_desc d { detail::box_data(box), WC_F32, (int32_t)box.width(), (int32_t)box.height(), (int32_t)box.depth() };
    detail::check(wc_haar_forward(ctx, &d, WC_HOST, flat.data()), ctx, "wc_haar_forward");
    return flat;
}

// ---- calc_rmse_per_box(): src/calc-loss.cpp:12-43 -----------------------------------------------------------
inline std::vector<double> calc_rmse_per_box(const multiBox3D& actual, const multiBox3D& pred, int num_components) {
    wc_ctx* ctx = detail::context();
    std::vector<wc_box_desc> a(num_components), b(num_components);
    for (int c = 0; c < num_components; ++c) {
        a[c] = { detail::box_data(actual[c]), WC_F32, (int32_t)actual[c].width(), (int32_t)actual[c].height(), (int32_t)actual[c].depth() };
        b[c] = { detail::box_data(pred[c]), WC_F32, (int32_t)pred[c].width(), (int32_t)pred[c].height(), (int32_t)pred[c].depth() };
    }
    std::vector<double> rmse(num_components, 0.0);
    detail::check(wc_rmse_batch(ctx, a.data(), b.data(), num_components, WC_HOST, rmse.data()), ctx, "wc_rmse_batch");
    return rmse;
}

inline double calc_adj_loss(double rmse, double range) { return rmse / range; }   // src/calc-loss.cpp:49-51

// ---- batched: the whole (t, level, box) x component iteration of src/modes.cpp:100-103 in one call -----------
namespace detail {
struct UnitKey { int t, l, b, c; size_t ci; };

// on_unit(unit index, packed unit with host pairs) runs on the pool's threads, fed chunk by chunk while the GPU is
// still working on later chunks; returns after every job has finished.  The plan stays alive for the caller.
template <class F>
inline wc_plan* compress_units_overlapped(wc_ctx* ctx, const std::vector<wc_box_desc>& in, double keep,
                                          std::vector<wc_packed>& out, unsigned lzma_threads, F on_unit) {
    wc_plan* plan = nullptr;
    check(wc_plan_create(ctx, in.data(), (int)in.size(), WC_HOST, &plan), ctx, "wc_plan_create");
    out.resize(in.size());
    WorkerPool pool(lzma_threads);
    struct Ctx { WorkerPool* pool; F* f; } cb { &pool, &on_unit };
    auto trampoline = [](void* user, int first, int n, const wc_packed* units) {
        Ctx* c = static_cast<Ctx*>(user);
        for (int j = 0; j < n; ++j) {
            const wc_packed u = units[j];                 // pinned pairs stay valid until the next compress on the plan
            const size_t    i = (size_t)first + j;
            c->pool->submit([c, i, u] { (*c->f)(i, u); });
        }
    };
    check(wc_plan_compress_to_host_chunked(plan, keep, out.data(), trampoline, &cb), ctx, "wc_plan_compress_to_host_chunked");
    pool.wait();
    return plan;
}
} // namespace detail

// boxes[t][lev][box] is the reference's AllData::boxes.  The .xz files are written by `lzma_threads` host threads
// (each file is independent; 0 = hardware_concurrency) while the GPU is still compressing later chunks.
inline void compress_all(std::vector<std::vector<std::vector<multiBox3D>>>& boxes, const std::vector<int>& comp_idxs,
                         double keep, const std::string& compressed_dir, unsigned lzma_threads = 0) {
    wc_ctx* ctx = detail::context();
    std::vector<detail::UnitKey> keys;
    std::vector<wc_box_desc>     in;
    for (size_t t = 0; t < boxes.size(); ++t)
        for (size_t l = 0; l < boxes[t].size(); ++l)
            for (size_t b = 0; b < boxes[t][l].size(); ++b)
                for (size_t c = 0; c < comp_idxs.size(); ++c) {
                    const Box3D& bx = boxes[t][l][b][c];
                    keys.push_back({ (int)t, (int)l, (int)b, comp_idxs[c], c });
                    in.push_back({ detail::box_data(bx), WC_F32, (int32_t)bx.width(), (int32_t)bx.height(), (int32_t)bx.depth() });
                }
    std::vector<wc_packed> out;
    wc_plan* plan = detail::compress_units_overlapped(ctx, in, keep, out, lzma_threads, [&](size_t i, const wc_packed& u) {
        detail::write_unit_file(detail::unit_path(compressed_dir, keys[i].t, keys[i].l, keys[i].c, keys[i].b), u);
    });
    wc_plan_destroy(plan);
}

// Mirror of the decompress loop of src/modes.cpp:151-166: reads + xz-decodes every file with host threads, one
// pipelined GPU pass for U + I through a decode plan, returns regen_boxes[t][lev][box] = multiBox3D.
inline std::vector<std::vector<std::vector<multiBox3D>>>
decompress_all(const std::string& compressed_dir, const std::vector<std::vector<int>>& box_counts,
               const std::vector<int>& comp_idxs, unsigned lzma_threads = 0) {
    wc_ctx* ctx = detail::context();
    std::vector<detail::UnitKey> keys;
    for (size_t t = 0; t < box_counts.size(); ++t)
        for (size_t l = 0; l < box_counts[t].size(); ++l)
            for (int b = 0; b < box_counts[t][l]; ++b)
                for (size_t c = 0; c < comp_idxs.size(); ++c) keys.push_back({ (int)t, (int)l, b, comp_idxs[c], c });
    std::vector<std::string> raw(keys.size());
    std::vector<wc_packed>   in(keys.size());
    unsigned nt = lzma_threads ? lzma_threads : std::thread::hardware_concurrency();
    detail::parallel_for(keys.size(), nt, [&](size_t i) {
        const std::string path = detail::unit_path(compressed_dir, keys[i].t, keys[i].l, keys[i].c, keys[i].b);
        raw[i] = detail::xz_decode_file(path);
        in[i]  = detail::parse_unit(raw[i], path);        // validated: sizes, counts, payload length
    });
    std::vector<std::vector<std::vector<multiBox3D>>> regen(box_counts.size());
    for (size_t t = 0; t < box_counts.size(); ++t) {
        regen[t].resize(box_counts[t].size());
        for (size_t l = 0; l < box_counts[t].size(); ++l) regen[t][l].resize((size_t)box_counts[t][l]);
    }
    // one dense pair stream (the files' bytes [20, 20+8K) back to back) + the counts: what wc_dplan_decode takes
    size_t total = 0;
    for (auto const& p : in) total += (size_t)p.npairs;
    std::vector<wc_pair>    stream(total ? total : 1);
    std::vector<int32_t>    counts(keys.size());
    std::vector<wc_box_out> out(keys.size());
    size_t off = 0;
    for (size_t i = 0; i < keys.size(); ++i) {
        if (in[i].npairs) std::memcpy(stream.data() + off, raw[i].data() + 20, 8 * (size_t)in[i].npairs);
        off += (size_t)in[i].npairs;
        counts[i] = in[i].npairs;
        const int32_t* h = in[i].shape;
        multiBox3D& mb = regen[keys[i].t][keys[i].l][keys[i].b];
        if (mb.size() < comp_idxs.size()) mb.resize(comp_idxs.size());
        mb[keys[i].ci] = Box3D((size_t)h[0], (size_t)h[1], (size_t)h[2]);
        Box3D& bx = mb[keys[i].ci];
        out[i] = { bx.data_size() ? &bx(0, 0, 0) : nullptr, WC_F32, h[0], h[1], h[2] };
    }
    wc_dplan* dp = nullptr;
    detail::check(wc_dplan_create(ctx, out.data(), (int)out.size(), WC_HOST, &dp), ctx, "wc_dplan_create");
    detail::check(wc_dplan_decode(dp, stream.data(), counts.data(), WC_HOST), ctx, "wc_dplan_decode");
    detail::check(wc_dplan_finish(dp), ctx, "wc_dplan_finish");
    wc_dplan_destroy(dp);
    return regen;
}

// ---- estimate: the body of `-estimate` (src/modes.cpp:236-324) on one plan, device-resident ------------------
struct Estimate {
    std::vector<double> mean_rmse;      // per component: std::accumulate(rmse) / size            src/modes.cpp:284-285
    std::vector<double> adjusted_loss;  // per component: calc_adj_loss(mean_rmse, max - min)     src/modes.cpp:289
    std::vector<float>  min_values, max_values;   // per component, the reference's running extrema incl. its initial
                                                  // values FLT_MAX / FLT_MIN                      src/preprocess.cpp:30-31,82-88
    std::vector<int>    npairs;         // per unit (box-major, component-minor)
    std::vector<char>   need32;         // per unit: CompressedWavelet::need32                     src/compressor.cpp:224-229
    size_t              xz_bytes = 0;   // total size of the .xz streams (what calc_size(scratch_dir) sums, :321)
    uint64_t            h2d_bytes = 0;  // bytes the call moved host -> device (the boxes, once)
};

// boxes = AllData::boxes[0][0] (one file, one level: src/modes.cpp:215-233); num_components = boxes[b].size().
inline Estimate estimate_all(const std::vector<multiBox3D>& boxes, int num_components, double keep, unsigned lzma_threads = 0) {
    wc_ctx* ctx = detail::context();
    std::vector<wc_box_desc> in;
    for (auto const& mb : boxes)
        for (int c = 0; c < num_components; ++c)
            in.push_back({ detail::box_data(mb[c]), WC_F32, (int32_t)mb[c].width(), (int32_t)mb[c].height(), (int32_t)mb[c].depth() });
    const size_t n = in.size();
    Estimate est;
    uint64_t h2d0 = 0, h2d1 = 0;
    wc_get_counter(ctx, WC_CTR_H2D_BYTES, &h2d0);
    detail::check(wc_set_option(ctx, WC_OPT_INGEST_STATS, 1), ctx, "wc_set_option");
    std::vector<size_t>    xz(n, 0);
    std::vector<wc_packed> out;
    wc_plan* plan = detail::compress_units_overlapped(ctx, in, keep, out, lzma_threads, [&](size_t i, const wc_packed& u) {
        xz[i] = detail::xz_encode(detail::serialize(u)).size();      // the size of the file compress() would write
    });
    wc_set_option(ctx, WC_OPT_INGEST_STATS, 0);
    // U, I, R in HBM: reconstruct into device boxes, RMSE against the plan's own (device) inputs
    size_t cells = 0;
    for (auto const& d : in) cells += (size_t)d.nx * d.ny * d.nz;
    void* dbuf = nullptr;
    detail::check(wc_device_alloc(ctx, &dbuf, sizeof(float) * (cells ? cells : 1)), ctx, "wc_device_alloc");
    std::vector<wc_box_out>  rec(n);
    std::vector<wc_box_desc> recd(n);
    size_t off = 0;
    for (size_t i = 0; i < n; ++i) {
        float* p = static_cast<float*>(dbuf) + off;
        rec[i]  = { p, WC_F32, in[i].nx, in[i].ny, in[i].nz };
        recd[i] = { p, WC_F32, in[i].nx, in[i].ny, in[i].nz };
        off += (size_t)in[i].nx * in[i].ny * in[i].nz;
    }
    std::vector<double>  rmse(n ? n : 1);
    std::vector<float>   lo(n ? n : 1), hi(n ? n : 1);
    std::vector<int32_t> n32(n ? n : 1);
    detail::check(wc_plan_decompress(plan, rec.data(), WC_DEVICE), ctx, "wc_plan_decompress");
    detail::check(wc_plan_rmse(plan, recd.data(), rmse.data()), ctx, "wc_plan_rmse");
    detail::check(wc_plan_unit_stats(plan, lo.data(), hi.data(), n32.data()), ctx, "wc_plan_unit_stats");
    wc_device_free(ctx, dbuf);
    wc_plan_destroy(plan);
    wc_get_counter(ctx, WC_CTR_H2D_BYTES, &h2d1);
    est.h2d_bytes = h2d1 - h2d0;
    est.min_values.assign((size_t)num_components, std::numeric_limits<float>::max());
    est.max_values.assign((size_t)num_components, std::numeric_limits<float>::min());   // sic: smallest POSITIVE float
    std::vector<double> sum((size_t)num_components, 0.0);
    for (size_t i = 0; i < n; ++i) {
        const size_t c = i % (size_t)num_components;
        if (lo[i] < est.min_values[c]) est.min_values[c] = lo[i];
        if (hi[i] > est.max_values[c]) est.max_values[c] = hi[i];
        sum[c] += rmse[i];                                   // left to right, as std::accumulate
        est.npairs.push_back(out[i].npairs);
        est.need32.push_back((char)(n32[i] != 0));
        est.xz_bytes += xz[i];
    }
    for (int c = 0; c < num_components; ++c) {
        const double mean = boxes.empty() ? 0.0 : sum[(size_t)c] / (double)boxes.size();
        est.mean_rmse.push_back(mean);
        est.adjusted_loss.push_back(calc_adj_loss(mean, est.max_values[(size_t)c] - est.min_values[(size_t)c]));
    }
    return est;
}

} // namespace wcgpu
