// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// oracle/_ref/libwcref.so: the reference's OWN hot-path sources, compiled unmodified from where
// they lie under $(REF)/src (default /root/reference/src; nothing is copied into this repo), with
// a flat C API on top so tests and the bench's cpu_baseline leg can call them through ctypes.
//
// compressor.cpp is #included (not linked) because wavelet_decompose, rle_encode and
// serialize_compressed_wavelet are `static` there (src/compressor.cpp:24,55,85); the same is
// done for decompressor.cpp (static rle_decode, src/decompressor.cpp:14) and calc-loss.cpp.
// Third-party headers the sources include are satisfied by oracle/shim/ (doctest, spdlog, lzma).
//
// What each entry point runs:
//   wcref_haar_forward   -> wavelet_decompose                (src/compressor.cpp:85-185)
//   wcref_haar_inverse   -> inverse_wavelet_decompose        (src/decompressor.cpp:79-159)
//   wcref_rle_encode     -> rle_encode                       (src/compressor.cpp:24-42)
//   wcref_rle_decode     -> rle_decode                       (src/decompressor.cpp:14-30)
//   wcref_serialize      -> serialize_compressed_wavelet     (src/compressor.cpp:55-80)
//   wcref_deserialize    -> deserialize_compressed_wavelet   (src/decompressor.cpp:35-74)
//   wcref_compress       -> compress(multiBox3D&, ...)       (src/compressor.cpp:192-297) incl. LZMA + file
//   wcref_decompress     -> decompress(std::string, ...)     (src/decompressor.cpp:238-255) incl. LZMA + file
//   wcref_rmse           -> calc_rmse_per_box                (src/calc-loss.cpp:12-43)
//   wcref_adj_loss       -> calc_adj_loss                    (src/calc-loss.cpp:49-51)
//   wcref_run_doctests   -> every TEST_CASE in the three files (src/compressor.cpp:300-406,
//                           src/calc-loss.cpp:68-86)
#include "compressor.cpp"
#include "decompressor.cpp"
#include "calc-loss.cpp"

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

Box3D box_from(const float* src, int x, int y, int z) {
    Box3D b(x, y, z, 0.0f);
    // Grid3D storage is x-fastest (src/grid.h:18) -> identical to the caller's layout
    if ((size_t)x * y * z > 0) std::memcpy(&b(0, 0, 0), src, sizeof(float) * (size_t)x * y * z);
    return b;
}

void box_to(const Box3D& b, float* dst) {
    size_t n = b.data_size();
    if (n > 0) std::memcpy(dst, &b(0, 0, 0), sizeof(float) * n);
}

} // namespace

extern "C" {

int wcref_run_doctests(int* n_cases, int* n_assertions) {
    wc_doctest::failures()   = 0;
    wc_doctest::assertions() = 0;
    for (auto const& c : wc_doctest::registry()) c.fn();
    if (n_cases) *n_cases = (int)wc_doctest::registry().size();
    if (n_assertions) *n_assertions = wc_doctest::assertions();
    return wc_doctest::failures();
}

const char* wcref_doctest_name(int i) {
    auto& r = wc_doctest::registry();
    return (i >= 0 && i < (int)r.size()) ? r[i].name : nullptr;
}

void wcref_haar_forward(const float* box, int x, int y, int z, float* flat_out) {
    Box3D              b = box_from(box, x, y, z);
    std::vector<float> w = wavelet_decompose(b);
    if (!w.empty()) std::memcpy(flat_out, w.data(), sizeof(float) * w.size());
}

void wcref_haar_inverse(const float* flat, int x, int y, int z, float* box_out) {
    std::vector<float> f(flat, flat + (size_t)x * y * z);
    Box3D              b = inverse_wavelet_decompose(std::move(f), x, y, z);
    box_to(b, box_out);
}

int wcref_rle_encode(const uint8_t* mask, int n, const float* values, int n_values,
                     int32_t* runs_out, float* vals_out) {
    std::vector<bool>  m(n);
    for (int i = 0; i < n; ++i) m[i] = mask[i] != 0;
    std::vector<float> v(values, values + n_values);
    auto               rle = rle_encode(m, v);
    for (size_t i = 0; i < rle.size(); ++i) {
        runs_out[i] = rle[i].first;
        vals_out[i] = rle[i].second;
    }
    return (int)rle.size();
}

void wcref_rle_decode(const int32_t* runs, const float* vals, int k, int total, float* out) {
    std::vector<std::pair<int, float>> rle(k);
    for (int i = 0; i < k; ++i) rle[i] = { runs[i], vals[i] };
    std::vector<float> r = rle_decode(std::move(rle), total);
    if (total > 0) std::memcpy(out, r.data(), sizeof(float) * (size_t)total);
}

long wcref_serialize(const int32_t* shape, int ncoef, const int32_t* runs, const float* vals,
                     int k, uint8_t* out, long cap) {
    CompressedWavelet cw;
    cw.shape       = { shape[0], shape[1], shape[2] };
    cw.coeff_shape = { ncoef };
    cw.rle_encoded.resize(k);
    for (int i = 0; i < k; ++i) cw.rle_encoded[i] = { runs[i], vals[i] };
    cw.need32     = false;
    std::string s = serialize_compressed_wavelet(cw);
    if ((long)s.size() > cap) return -(long)s.size();
    std::memcpy(out, s.data(), s.size());
    return (long)s.size();
}

int wcref_deserialize(const uint8_t* buf, long len, int32_t* shape, int32_t* ncoef,
                      int32_t* runs_out, float* vals_out, int cap) {
    std::string       s(reinterpret_cast<const char*>(buf), (size_t)len);
    CompressedWavelet cw = deserialize_compressed_wavelet(s);
    for (int i = 0; i < 3; ++i) shape[i] = cw.shape[i];
    *ncoef = cw.coeff_shape[0];
    int k  = (int)cw.rle_encoded.size();
    if (k > cap) return -k;
    for (int i = 0; i < k; ++i) {
        runs_out[i] = cw.rle_encoded[i].first;
        vals_out[i] = cw.rle_encoded[i].second;
    }
    return k;
}

// boxes: ncomp boxes of x*y*z floats, back to back (multiBox3D is SoA by component,
// src/box-structs.h:10).  comp_ids: the Header indices that end up in the file names
// (src/compressor.cpp:250-254).  Outputs per component c: npairs_out[c] and the pairs at
// runs_out/vals_out + c*N.  The .xz files are written into `dir` exactly as the reference does.
int wcref_compress(const float* boxes, int ncomp, int x, int y, int z, double keep, int t, int lev,
                   int box_idx, const int32_t* comp_ids, const char* dir, int32_t* npairs_out,
                   int32_t* runs_out, float* vals_out) {
    size_t     n = (size_t)x * y * z;
    multiBox3D mb;
    for (int c = 0; c < ncomp; ++c) mb.push_back(box_from(boxes + c * n, x, y, z));
    std::vector<int> comps(comp_ids, comp_ids + ncomp);
    auto             out = compress(mb, comps, keep, t, lev, box_idx, std::string(dir));
    if ((int)out.size() != ncomp) return -1;
    for (int c = 0; c < ncomp; ++c) {
        auto const& cw = out[c];
        npairs_out[c]  = (int32_t)cw.rle_encoded.size();
        if (runs_out && vals_out) {
            for (size_t i = 0; i < cw.rle_encoded.size(); ++i) {
                runs_out[c * n + i] = cw.rle_encoded[i].first;
                vals_out[c * n + i] = cw.rle_encoded[i].second;
            }
        }
    }
    return 0;
}

// Returns the number of cells written (x*y*z) or -needed if cap is too small.
long wcref_decompress(const char* path, float* box_out, long cap, int32_t* dims) {
    Box3D b = decompress(std::string(path), 0, 0, 0, 0);
    dims[0] = (int32_t)b.width();
    dims[1] = (int32_t)b.height();
    dims[2] = (int32_t)b.depth();
    long n  = (long)b.data_size();
    if (n > cap) return -n;
    box_to(b, box_out);
    return n;
}

void wcref_rmse(const float* actual, const float* pred, int ncomp, int x, int y, int z,
                double* rmse_out) {
    size_t     n = (size_t)x * y * z;
    multiBox3D a, p;
    for (int c = 0; c < ncomp; ++c) {
        a.push_back(box_from(actual + c * n, x, y, z));
        p.push_back(box_from(pred + c * n, x, y, z));
    }
    std::vector<double> r = calc_rmse_per_box(a, p, ncomp);
    for (int c = 0; c < ncomp; ++c) rmse_out[c] = r[c];
}

double wcref_adj_loss(double rmse, double range) { return calc_adj_loss(rmse, range); }

} // extern "C"
