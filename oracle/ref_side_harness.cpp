// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// oracle/_ref/libwcref_side.so: the reference's OWN side-file code (src/readandwrite.cpp, unmodified,
// #included from where it lies; it needs no AMReX) behind a flat C API, so that the Python writer / reader of
// the five .raw files (wavelet-compression_b200/sidefiles.py) can be checked byte for byte against it:
//
//   wcref_side_write     -> write_runinfo / write_loc_dim_to_bin x2 / write_box_counts / write_amrexinfo
//                           (src/readandwrite.cpp:226-395) as src/modes.cpp:71-89 calls them
//   wcref_side_rewrite   -> reads the five files of dir_in with the reference's readers (src/modes.cpp:117-181 order)
//                           and writes what it read into dir_out with the reference's writers
//   wcref_side_doctests  -> the four TEST_CASEs of src/readandwrite.cpp:397-490
// <iostream> first: this image's g++ links libstdc++ statically into the .so, and only the ios_base::Init object that
// <iostream> defines initialises that private copy's locale / stream state when the library is dlopen'ed.
#include <iostream>

#include "readandwrite.cpp"

#include <sstream>

namespace {
std::vector<std::string> split_lines(const char* s) {
    std::vector<std::string> out;
    std::istringstream iss(s ? s : "");
    std::string ln;
    while (std::getline(iss, ln)) out.push_back(ln);
    return out;
}
} // namespace

extern "C" {

int wcref_side_doctests(int* n_cases) {
    wc_doctest::failures() = 0;
    for (auto const& c : wc_doctest::registry()) c.fn();
    if (n_cases) *n_cases = (int)wc_doctest::registry().size();
    return wc_doctest::failures();
}

// dir must end with '/' (the reference concatenates path + file name, src/readandwrite.cpp:200).
// locs / dims: 3 ints per box in (t, level, box) order; times_text: one decimal string per timestep, parsed with
// operator>> into long double exactly as src/preprocess.cpp:183-185 does.
int wcref_side_write(const char* dir, int num_times, int num_levels, const int* counts, const int* locs, const int* dims,
                     const char* files_nl, int min_level, int max_level, const char* comps_nl, const int* comp_idxs,
                     int n_comp, const double* geomcell6, const int* ref_ratios3, const char* times_text_nl,
                     const int* level_steps, int xdim, int ydim, int zdim) {
    std::vector<std::vector<int>> bc(num_times, std::vector<int>(num_levels));
    for (int t = 0; t < num_times; ++t)
        for (int l = 0; l < num_levels; ++l) bc[t][l] = counts[t * num_levels + l];
    LocDimData L(num_times, std::vector<std::vector<std::vector<int>>>(num_levels)), D = L;
    size_t k = 0;
    for (int t = 0; t < num_times; ++t)
        for (int l = 0; l < num_levels; ++l)
            for (int b = 0; b < bc[t][l]; ++b, ++k) {
                L[t][l].push_back({ locs[3 * k], locs[3 * k + 1], locs[3 * k + 2] });
                D[t][l].push_back({ dims[3 * k], dims[3 * k + 1], dims[3 * k + 2] });
            }
    RunInfo ri;
    ri.files = split_lines(files_nl);
    ri.min_level = min_level;
    ri.max_level = max_level;
    ri.components = split_lines(comps_nl);
    ri.comp_idxs.assign(comp_idxs, comp_idxs + n_comp);
    AMReXInfo ai;
    for (int t = 0; t < num_times; ++t) ai.geomcellinfo.push_back(std::vector<double>(geomcell6 + 6 * t, geomcell6 + 6 * t + 6));
    ai.ref_ratios.assign(ref_ratios3, ref_ratios3 + 3);
    for (auto const& s : split_lines(times_text_nl)) {
        std::istringstream iss(s);
        long double v = 0;
        iss >> v;
        ai.true_times.push_back(v);
    }
    for (int t = 0; t < num_times; ++t) ai.level_steps.push_back(std::vector<int>(level_steps + t * num_levels, level_steps + (t + 1) * num_levels));
    ai.xDim = xdim; ai.yDim = ydim; ai.zDim = zdim;
    AMRIterator it(num_times, num_levels, bc, n_comp);
    std::string d(dir);
    write_runinfo(ri, d, "runinfo.raw");
    write_loc_dim_to_bin(L, d, "locations.raw", it);
    write_loc_dim_to_bin(D, d, "dimensions.raw", it);
    write_box_counts(bc, d, "boxcounts.raw", num_times, num_levels);
    write_amrexinfo(ai, d, "amrexinfo.raw");
    return 0;
}

int wcref_side_rewrite(const char* dir_in, const char* dir_out) {
    std::string in(dir_in), out(dir_out);
    RunInfo ri = read_runinfo(in, "runinfo.raw");
    int num_times = (int)ri.files.size(), num_levels = ri.max_level - ri.min_level + 1;
    auto bc = read_box_counts(in, "boxcounts.raw", num_times, num_levels);
    AMRIterator it(num_times, num_levels, bc, ri.components.size());
    AMReXInfo ai = read_amrex_info(in, "amrexinfo.raw");
    LocDimData L = read_loc_dim_from_bin(in, "locations.raw", bc, it, num_times, num_levels);
    LocDimData D = read_loc_dim_from_bin(in, "dimensions.raw", bc, it, num_times, num_levels);
    write_runinfo(ri, out, "runinfo.raw");
    write_loc_dim_to_bin(L, out, "locations.raw", it);
    write_loc_dim_to_bin(D, out, "dimensions.raw", it);
    write_box_counts(bc, out, "boxcounts.raw", num_times, num_levels);
    write_amrexinfo(ai, out, "amrexinfo.raw");
    return num_times * 1000 + num_levels;
}

} // extern "C"
