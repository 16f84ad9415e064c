// Drop-in evidence: the reference's own doctest cases for this path (src/compressor.cpp:369-406,
// src/calc-loss.cpp:68-86), re-expressed against namespace wcgpu, compiled with the REFERENCE's
// own headers (grid.h / box-structs.h from $(REF)/src — see oracle/Makefile target `dropin`), so the
// types and signatures are the reference's, and every number comes from the GPU through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <string>
#include <thread>
#include <unistd.h>

#include "box-structs.h"   // the reference's: Grid3D, Box3D, multiBox3D, CompressedWavelet

#include "wc_dropin.hpp"

static int g_fail = 0;
#define REQUIRE(x)                                                                  \
    do {                                                                            \
        if (!(x)) {                                                                 \
            ++g_fail;                                                               \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x);     \
        }                                                                           \
    } while (0)

static std::string scratch_dir() {
    std::string t = (std::filesystem::temp_directory_path() / "wcgpu_dropin.XXXXXX").string();
    if (!mkdtemp(t.data())) { std::perror("mkdtemp"); std::exit(2); }
    return t;
}

int main() {
    using namespace wcgpu;

    {   // "Wavelet decomposition" (src/compressor.cpp:369-384)
        Box3D test(4, 8, 16, 5.0f);
        test.set(1, 2, 3, 8.5f);
        test.set(2, 5, 6, 5.44f);
        test.set(1, 1, 1, 3.3999932f);
        test.set(2, 2, 2, 3.19229f);
        test.set(3, 5, 12, 199.39029f);
        std::vector<float> wavelet = wavelet_decompose(test);
        Box3D result = inverse_wavelet_decompose(wavelet, 4, 8, 16);
        REQUIRE(test.equals(result, 1e-6));
        REQUIRE(wavelet.size() == 512);
        REQUIRE(wavelet[0] == 4.79999924f);   // SURVEY.md §8c known answer
    }
    {   // "File writing/compression" (src/compressor.cpp:387-406)
        Box3D box(4, 8, 16, 5.0f);
        multiBox3D test;
        test.push_back(std::move(box));
        std::vector<int> components = { 0 };
        std::string dir = scratch_dir();
        auto cw = compress(test, components, 0.999, 0, 0, 0, dir);
        REQUIRE(cw.size() == 1);
        REQUIRE(cw[0].rle_encoded.size() == 64);
        Box3D result = decompress(dir + "/compressed-wavelet-0-0-0-0.xz", 0, 0, 0, 0);
        REQUIRE(test[0].equals(result, 0));
        std::filesystem::remove_all(dir);
    }
    {   // "Calc RMSE" (src/calc-loss.cpp:68-86)
        Box3D b1(2, 2, 2, 0.0f), b2(2, 2, 2, 3.5f);
        multiBox3D t1, t2;
        t1.push_back(b1.clone()); t2.push_back(b2.clone());
        t1.push_back(b1.clone()); t2.push_back(b2.clone());
        std::vector<double> rmses = calc_rmse_per_box(t1, t2, 2);
        std::vector<double> truth = { 3.5, 3.5 };
        REQUIRE(rmses == truth);
    }
    {   // batched run: same files as the per-box calls, and a lossless-enough round trip
        const int T = 2, L = 2, NB = 3, NC = 2;
        std::vector<int> comp_idxs = { 0, 3 };
        std::vector<std::vector<std::vector<multiBox3D>>> boxes(T);
        std::vector<std::vector<int>> counts(T, std::vector<int>(L, NB));
        for (int t = 0; t < T; ++t) {
            boxes[t].resize(L);
            for (int l = 0; l < L; ++l)
                for (int b = 0; b < NB; ++b) {
                    multiBox3D mb;
                    for (int c = 0; c < NC; ++c) {
                        int X = l ? 32 : 16, Y = l ? 32 : 8, Z = l ? 32 : 12;
                        Box3D bx(X, Y, Z);
                        for (int k = 0; k < Z; ++k)
                            for (int j = 0; j < Y; ++j)
                                for (int i = 0; i < X; ++i)
                                    bx(i, j, k) = (c ? 0.f : 300.f) + 50.f * std::sin(0.1f * i + t) * std::cos(0.07f * j + b) *
                                                                      std::sin(0.05f * k + l);
                        mb.push_back(std::move(bx));
                    }
                    boxes[t][l].push_back(std::move(mb));
                }
        }
        std::string d1 = scratch_dir(), d2 = scratch_dir();
        double keep = (double)0.9999f;
        compress_all(boxes, comp_idxs, keep, d1, 4);
        for (int t = 0; t < T; ++t)
            for (int l = 0; l < L; ++l)
                for (int b = 0; b < NB; ++b) compress(boxes[t][l][b], comp_idxs, keep, t, l, b, d2);
        size_t nfiles = 0;
        for (auto& e : std::filesystem::directory_iterator(d1)) {
            std::string name = e.path().filename().string();
            std::ifstream fa(e.path(), std::ios::binary), fb(std::filesystem::path(d2) / name, std::ios::binary);
            std::string a((std::istreambuf_iterator<char>(fa)), {}), bb((std::istreambuf_iterator<char>(fb)), {});
            REQUIRE(!a.empty() && a == bb);
            ++nfiles;
        }
        REQUIRE(nfiles == (size_t)T * L * NB * NC);
        auto regen = decompress_all(d1 + "/", counts, comp_idxs, 4);
        for (int t = 0; t < T; ++t)
            for (int l = 0; l < L; ++l)
                for (int b = 0; b < NB; ++b) {
                    std::vector<double> r = calc_rmse_per_box(boxes[t][l][b], regen[t][l][b], NC);
                    for (int c = 0; c < NC; ++c) {
                        Box3D one = decompress(detail::unit_path(d1, t, l, comp_idxs[c], b), t, l, comp_idxs[c], b);
                        REQUIRE(one.equals(regen[t][l][b][c], 0));
                        REQUIRE(r[c] < 0.05);
                    }
                }
        std::filesystem::remove_all(d1);
        std::filesystem::remove_all(d2);
    }
    {   // estimate_all: the `-estimate` body on one plan (src/modes.cpp:236-324); BASELINE config 1 known answer
        // (bundled plt00074, level 0, temp, keep = 0.999f: 4096 + 8 pairs, RMSE 0, 168 + 88 bytes of .xz)
        std::vector<multiBox3D> boxes(2);
        boxes[0].push_back(Box3D(16, 32, 64, 3902.4f));
        boxes[1].push_back(Box3D(8, 4, 2, 16.0f));
        Estimate e = estimate_all(boxes, 1, (double)0.999f, 2);
        REQUIRE(e.npairs.size() == 2 && e.npairs[0] == 4096 && e.npairs[1] == 8);
        REQUIRE(e.mean_rmse.size() == 1 && e.mean_rmse[0] == 0.0 && e.adjusted_loss[0] == 0.0);
        REQUIRE(e.min_values[0] == 16.0f && e.max_values[0] == 3902.4f);
        REQUIRE(e.xz_bytes == 168 + 88);
        REQUIRE(!e.need32[0] && !e.need32[1]);
        REQUIRE(e.h2d_bytes == sizeof(float) * (16 * 32 * 64 + 8 * 4 * 2));     // the boxes cross the bus exactly once
        // a non-trivial level: per-box RMSE of estimate_all == the per-box calls, need32 on large values
        std::vector<multiBox3D> lvl;
        for (int b = 0; b < 5; ++b) {
            multiBox3D mb;
            for (int c = 0; c < 2; ++c) {
                Box3D bx(32, 32, 32);
                for (int k = 0; k < 32; ++k)
                    for (int j = 0; j < 32; ++j)
                        for (int i = 0; i < 32; ++i)
                            bx(i, j, k) = (c ? 90000.f : -3.f) + 50.f * std::sin(0.1f * i + b) * std::cos(0.07f * j) * std::sin(0.05f * k + c);
                mb.push_back(std::move(bx));
            }
            lvl.push_back(std::move(mb));
        }
        double keep = (double)0.999f;
        Estimate f = estimate_all(lvl, 2, keep, 3);
        std::string d = scratch_dir();
        std::vector<int> comps = { 0, 1 };
        std::vector<double> acc(2, 0.0);
        size_t xz = 0;
        for (int b = 0; b < 5; ++b) {
            auto cw = compress(lvl[b], comps, keep, 0, 0, b, d);
            multiBox3D regen;
            for (int c = 0; c < 2; ++c) {
                std::string path = detail::unit_path(d, 0, 0, c, b);
                xz += (size_t)std::filesystem::file_size(path);
                regen.push_back(decompress(path, 0, 0, c, b));
                REQUIRE((int)cw[c].rle_encoded.size() == f.npairs[(size_t)b * 2 + c]);
                REQUIRE(cw[c].need32 == (bool)f.need32[(size_t)b * 2 + c]);
            }
            std::vector<double> r = calc_rmse_per_box(lvl[b], regen, 2);
            acc[0] += r[0]; acc[1] += r[1];
        }
        REQUIRE(f.need32[1] && !f.need32[0]);
        REQUIRE(f.xz_bytes == xz);
        for (int c = 0; c < 2; ++c) REQUIRE(std::fabs(f.mean_rmse[c] - acc[c] / 5.0) <= 1e-12 * (acc[c] / 5.0));
        std::filesystem::remove_all(d);
    }
    {   // odd dimensions through compress() / decompress() (the x-slab kernels): the trailing element of an odd axis passes
        // through the forward stage (src/compressor.cpp:98-175) and the inverse leaves it at 0 (src/decompressor.cpp:99-108),
        // everything else comes back as for even boxes; keep = 1 keeps every non-zero coefficient
        const int shapes[3][3] = { { 5, 7, 3 }, { 41, 39, 37 }, { 40, 40, 42 } };
        for (auto& sh : shapes) {
            const int X = sh[0], Y = sh[1], Z = sh[2];
            Box3D bx(X, Y, Z), want(X, Y, Z);
            for (int k = 0; k < Z; ++k)
                for (int j = 0; j < Y; ++j)
                    for (int i = 0; i < X; ++i) {
                        bx(i, j, k) = 3.f + 0.5f * std::sin(0.1f * i) * std::cos(0.07f * j) * std::sin(0.05f * k + 1.f);
                        const bool trailing = ((X & 1) && i == X - 1) || ((Y & 1) && j == Y - 1) || ((Z & 1) && k == Z - 1);
                        want(i, j, k) = trailing ? 0.f : bx(i, j, k);
                    }
            multiBox3D mb;
            mb.push_back(bx.clone());
            std::vector<int> comps = { 0 };
            std::string d = scratch_dir();
            auto cw = compress(mb, comps, 1.0, 0, 0, 0, d);
            REQUIRE(cw.size() == 1 && cw[0].rle_encoded.size() <= (size_t)X * Y * Z);
            Box3D back = decompress(d + "/compressed-wavelet-0-0-0-0.xz", 0, 0, 0, 0);
            REQUIRE(want.equals(back, 1e-5f));
            std::filesystem::remove_all(d);
        }
    }
    {   // one host thread per GPU: contexts are per (thread, device) — two threads on device 0 must not share one
        std::atomic<int> ok { 0 };
        auto work = [&] {
            set_device(0);
            Box3D b1(2, 2, 2, 0.0f), b2(2, 2, 2, 3.5f);
            multiBox3D t1, t2;
            t1.push_back(b1.clone()); t2.push_back(b2.clone());
            for (int i = 0; i < 20; ++i)
                if (calc_rmse_per_box(t1, t2, 1)[0] == 3.5) ++ok;
        };
        std::thread a(work), b(work);
        a.join(); b.join();
        REQUIRE(ok == 40);
    }
    std::printf(g_fail ? "dropin FAILED (%d)\n" : "dropin ok\n", g_fail);
    return g_fail ? 1 : 0;
}
