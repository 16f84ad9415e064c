/* TEST INFRASTRUCTURE ONLY — the parity oracle.  Never linked into, imported by or executed from
 * the product path (wavelet-compression_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * A plain-C CPU restatement of the numeric core of carsonmw3/wavelet-compression, one function
 * per row of SURVEY.md §8(a).  Each function cites the reference lines it restates.  It is a
 * restatement (written from the algorithm, scalar, single-threaded), not a copy: parity is
 * PINNED by tests/test_oracle_vs_ref.py against oracle/_ref/libwcref.so (the reference's own
 * unmodified sources compiled here) and by the golden vectors under tests/golden/ that
 * oracle/make_golden.py generated from that library, including the reference's doctest vectors.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math: the arithmetic below
 * relies on IEEE-754 binary32/binary64 evaluation exactly as written).
 *
 * Layout conventions: a box is X*Y*Z float32, x-fastest, m = i + X*(j + Y*k)
 * (src/grid.h:18).  Coefficients are flattened z-fastest, f = (i*Y + j)*Z + k
 * (src/compressor.cpp:178-181).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WCO_API __attribute__((visibility("default")))

/* ---- A1: ingest narrowing, src/preprocess.cpp:78 (`float value = mfdata(i,j,k,comp)`) ------ */
WCO_API void wco_narrow_f64(const double* src, long n, float* dst) {
    for (long m = 0; m < n; ++m) dst[m] = (float)src[m]; /* round-to-nearest-even */
}

/* One 1-D pass of the reference's transform along `axis` for every line of the volume.
 * Forward (src/compressor.cpp:98-125 for Z, :128-150 for Y, :153-175 for X): consecutive pairs
 * (a,b) -> low=(a+b)/2.0, high=(a-b)/2.0; a+b / a-b are evaluated in float, the division in
 * double, the result stored back to float; lows go to [0,h), highs to [h,2h); an odd trailing
 * element keeps its value. */
static void forward_axis(float* vol, int X, int Y, int Z, int axis) {
    int  dims[3]   = { X, Y, Z };
    long stride[3] = { 1, X, (long)X * Y };
    int  n         = dims[axis];
    int  h         = n / 2;
    int  u_ax = (axis + 1) % 3, v_ax = (axis + 2) % 3;
    float* line = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int u = 0; u < dims[u_ax]; ++u) {
        for (int v = 0; v < dims[v_ax]; ++v) {
            float* base = vol + u * stride[u_ax] + v * stride[v_ax];
            for (int p = 0; p < n; ++p) line[p] = base[p * stride[axis]];
            for (int p = 0; p < h; ++p) {
                float a = line[2 * p], b = line[2 * p + 1];
                float s = a + b, d = a - b;
                base[p * stride[axis]]       = (float)((double)s / 2.0);
                base[(h + p) * stride[axis]] = (float)((double)d / 2.0);
            }
            /* element n-1 of an odd line is left as it was */
        }
    }
    free(line);
}

/* ---- F: wavelet_decompose, src/compressor.cpp:85-185 -------------------------------------- */
WCO_API void wco_haar_forward(const float* box, int X, int Y, int Z, float* flat) {
    size_t n = (size_t)X * Y * Z;
    if (n == 0) return;
    float* t = (float*)malloc(sizeof(float) * n);
    memcpy(t, box, sizeof(float) * n);
    forward_axis(t, X, Y, Z, 2); /* Z first  (:98-125)  */
    forward_axis(t, X, Y, Z, 1); /* then Y   (:128-150) */
    forward_axis(t, X, Y, Z, 0); /* then X   (:153-175) */
    /* flatten with z fastest (:178-181) */
    size_t f = 0;
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            for (int k = 0; k < Z; ++k) flat[f++] = t[i + (size_t)X * (j + (size_t)Y * k)];
    free(t);
}

/* ---- T: threshold selection, src/compressor.cpp:212-216 ------------------------------------
 * max_val = signed value of the FIRST element (lowest f) whose |.| is largest, found with the
 * sequential std::max_element rule "replace when |best| < |candidate|" evaluated in double (so a
 * NaN candidate never replaces, and a NaN at f=0 is never replaced); thresh = max_val*(1-keep)
 * in double.  keep is the caller's double (the CLI widens a float, src/argparse.h:13). */
WCO_API double wco_select_threshold(const float* flat, long n, double keep, long* argmax_out) {
    long best = 0;
    if (n <= 0) {
        if (argmax_out) *argmax_out = -1;
        return 0.0;
    }
    for (long f = 1; f < n; ++f) {
        if (fabs((double)flat[best]) < fabs((double)flat[f])) best = f;
    }
    if (argmax_out) *argmax_out = best;
    double max_val = (double)flat[best];
    return max_val * (1 - keep);
}

/* ---- M + P: mask (src/compressor.cpp:222-234) and rle_encode (:24-42) ----------------------
 * keep coefficient f iff |(double)c| > thresh; emit (number of dropped coefficients since the
 * previous kept one, value) in f order; trailing drops emit nothing.  Returns K. */
WCO_API long wco_threshold_pack(const float* flat, long n, double thresh, int32_t* runs,
                                float* vals) {
    long    k   = 0;
    int32_t run = 0;
    for (long f = 0; f < n; ++f) {
        double v = (double)flat[f];
        if (fabs(v) > thresh) {
            runs[k] = run;
            vals[k] = (float)v;
            ++k;
            run = 0;
        } else {
            ++run;
        }
    }
    return k;
}

/* ---- S: serialize_compressed_wavelet, src/compressor.cpp:47-80 -----------------------------
 * native-endian int32 X,Y,Z | int32 ncoef | int32 K | K x (int32 run, float32 val) = 20+8K bytes */
WCO_API long wco_serialize(const int32_t shape[3], int32_t ncoef, const int32_t* runs,
                           const float* vals, int32_t k, uint8_t* out) {
    uint8_t* p = out;
    memcpy(p, shape, 12);
    p += 12;
    memcpy(p, &ncoef, 4);
    p += 4;
    memcpy(p, &k, 4);
    p += 4;
    for (int32_t i = 0; i < k; ++i) {
        memcpy(p, &runs[i], 4);
        memcpy(p + 4, &vals[i], 4);
        p += 8;
    }
    return (long)(p - out);
}

/* inverse of S: deserialize_compressed_wavelet, src/decompressor.cpp:35-74.  Returns K. */
WCO_API long wco_deserialize(const uint8_t* buf, int32_t shape[3], int32_t* ncoef, int32_t* runs,
                             float* vals, long cap) {
    int32_t k;
    memcpy(shape, buf, 12);
    memcpy(ncoef, buf + 12, 4);
    memcpy(&k, buf + 16, 4);
    if (k > cap) return -(long)k;
    for (int32_t i = 0; i < k; ++i) {
        memcpy(&runs[i], buf + 20 + 8 * (size_t)i, 4);
        memcpy(&vals[i], buf + 24 + 8 * (size_t)i, 4);
    }
    return k;
}

/* ---- U: rle_decode, src/decompressor.cpp:14-30 ---------------------------------------------
 * zero-filled float[total]; idx += run; if (idx < total) { out[idx] = val; ++idx; }
 * (idx is an int in the reference; a pair that lands at or past `total` is dropped but idx keeps
 * its advanced value, so every later pair is dropped as well.) */
WCO_API void wco_rle_decode(const int32_t* runs, const float* vals, long k, long total,
                            float* out) {
    for (long f = 0; f < total; ++f) out[f] = 0.0f;
    int32_t idx = 0;
    for (long i = 0; i < k; ++i) {
        idx += runs[i];
        if (idx < total) {
            out[idx] = vals[i];
            ++idx;
        }
    }
}

/* inverse 1-D pass (src/decompressor.cpp:90-114 X, :117-135 Y, :138-156 Z): avg=in[p],
 * diff=in[h+p] widened to double, out[2p]=avg+diff, out[2p+1]=avg-diff narrowed to float;
 * `restored` starts zero-filled and only 2h entries are written, so an odd trailing element
 * becomes 0. */
static void inverse_axis(float* vol, int X, int Y, int Z, int axis) {
    int  dims[3]   = { X, Y, Z };
    long stride[3] = { 1, X, (long)X * Y };
    int  n         = dims[axis];
    int  h         = n / 2;
    int  u_ax = (axis + 1) % 3, v_ax = (axis + 2) % 3;
    double* line = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int u = 0; u < dims[u_ax]; ++u) {
        for (int v = 0; v < dims[v_ax]; ++v) {
            float* base = vol + u * stride[u_ax] + v * stride[v_ax];
            for (int p = 0; p < n; ++p) line[p] = (double)base[p * stride[axis]];
            for (int p = 0; p < h; ++p) {
                double avg = line[p], diff = line[h + p];
                base[(2 * p) * stride[axis]]     = (float)(avg + diff);
                base[(2 * p + 1) * stride[axis]] = (float)(avg - diff);
            }
            if (n & 1) base[(n - 1) * stride[axis]] = 0.0f;
        }
    }
    free(line);
}

/* ---- I: inverse_wavelet_decompose, src/decompressor.cpp:79-159 ----------------------------- */
WCO_API void wco_haar_inverse(const float* flat, int X, int Y, int Z, float* box) {
    size_t n = (size_t)X * Y * Z;
    if (n == 0) return;
    size_t f = 0;
    for (int i = 0; i < X; ++i) /* un-flatten (:82-87) */
        for (int j = 0; j < Y; ++j)
            for (int k = 0; k < Z; ++k) box[i + (size_t)X * (j + (size_t)Y * k)] = flat[f++];
    inverse_axis(box, X, Y, Z, 0); /* X first (:90-114) */
    inverse_axis(box, X, Y, Z, 1); /* then Y  (:117-135) */
    inverse_axis(box, X, Y, Z, 2); /* then Z  (:138-156) */
}

/* ---- R: calc_rmse_per_box for one component, src/calc-loss.cpp:12-43 -----------------------
 * diff is the FLOAT difference widened to double (:33); squares are summed sequentially in
 * memory order k,j,i (:30-35); divided by the int product X*Y*Z (:39). */
WCO_API double wco_rmse(const float* actual, const float* pred, int X, int Y, int Z) {
    double sum = 0.0;
    size_t n   = (size_t)X * Y * Z;
    for (size_t m = 0; m < n; ++m) {
        float  df   = actual[m] - pred[m];
        double diff = (double)df;
        sum += diff * diff;
    }
    return sqrt(sum / (X * Y * Z));
}

/* ---- R': calc_adj_loss, src/calc-loss.cpp:49-51 -------------------------------------------- */
WCO_API double wco_adj_loss(double rmse, double range) { return rmse / range; }

/* ---- C (numeric part): F -> T -> M -> P for one unit, src/compressor.cpp:203-247 ----------- */
WCO_API long wco_compress_unit(const float* box, int X, int Y, int Z, double keep, int32_t* runs,
                               float* vals, double* thresh_out) {
    size_t n = (size_t)X * Y * Z;
    float* flat = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
    wco_haar_forward(box, X, Y, Z, flat);
    double thresh = wco_select_threshold(flat, (long)n, keep, NULL);
    if (thresh_out) *thresh_out = thresh;
    long k = wco_threshold_pack(flat, (long)n, thresh, runs, vals);
    free(flat);
    return k;
}

/* same with the float64 FAB slab as input (A1 narrowing fused in front) */
WCO_API long wco_compress_unit_f64(const double* box, int X, int Y, int Z, double keep,
                                   int32_t* runs, float* vals, double* thresh_out) {
    size_t n = (size_t)X * Y * Z;
    float* b32 = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
    wco_narrow_f64(box, (long)n, b32);
    long k = wco_compress_unit(b32, X, Y, Z, keep, runs, vals, thresh_out);
    free(b32);
    return k;
}

/* ---- D (numeric part): U -> I for one unit, src/decompressor.cpp:245-254 ------------------- */
WCO_API void wco_decompress_unit(const int32_t* runs, const float* vals, long k, int X, int Y,
                                 int Z, long ncoef, float* box) {
    size_t n = (size_t)X * Y * Z;
    size_t m = (size_t)(ncoef > 0 ? ncoef : 0);
    float* flat = (float*)calloc((m > n ? m : n) + 1, sizeof(float));
    wco_rle_decode(runs, vals, k, ncoef, flat);
    wco_haar_inverse(flat, X, Y, Z, box);
    free(flat);
}

/* ---- EXTENSION (no reference counterpart; BASELINE config 5): one threshold
 * shared by a set of units, max-semantics = the reference's rule (:212-216) applied to the
 * concatenation of the units' coefficient arrays in the given order.  PARITY UNPINNED by the
 * reference (it never shares a threshold, D4); the oracle defines the semantics. */
WCO_API double wco_select_threshold_global(const float* const* flats, const long* ns, int n_units,
                                           double keep) {
    int    have = 0;
    double best = 0.0;
    for (int u = 0; u < n_units; ++u) {
        for (long f = 0; f < ns[u]; ++f) {
            double c = (double)flats[u][f];
            if (!have) {
                best = c;
                have = 1;
            } else if (fabs(best) < fabs(c)) {
                best = c;
            }
        }
    }
    return best * (1 - keep);
}
