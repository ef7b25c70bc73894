// C++ host-side exercise of the C ABI through include/nle_b200.hpp (no Python, no OpenCV, no Eigen):
//   host_mirror_test <bgr.raw> <rows> <cols> <nRowSamples> <nColSamples> <hx> <hy> <nSinkhornIter> <nEigenVectors> <out.raw> w0 w1 ...
// -- the argument order of the reference's `enhance` CLI (enhance.cpp:20-31) with raw interleaved BGR files in place of
// cv::imread / cv::imwrite.  Also runs the reference's Catch2 known-answer test of eigenDecomposition
// (test/test_filter.cpp:42-68) and its error paths.  Prints "ok ..." and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "nle_b200.hpp"

using namespace nle_b200;

static int fail(const std::string& m) { std::cerr << "FAIL: " << m << std::endl; return 1; }

static int run(int argc, char** argv);

int main(int argc, char** argv) {
    if (argc < 12) { std::cerr << "usage: see the header comment" << std::endl; return 2; }
    try {
        return run(argc, argv);
    } catch (const std::exception& e) {            // e.g. no CUDA device: the library has no CPU fallback
        return fail(e.what());
    }
}

static int run(int argc, char** argv) {
    const std::string in_path = argv[1], out_path = argv[10];
    const int rows = atoi(argv[2]), cols = atoi(argv[3]);
    const int nRS = atoi(argv[4]), nCS = atoi(argv[5]);
    const double hx = atof(argv[6]), hy = atof(argv[7]);
    const int T = atoi(argv[8]), K = atoi(argv[9]);
    std::vector<double> weights;
    for (int i = 11; i < argc; ++i) weights.push_back(atof(argv[i]));

    // test_filter.cpp:42-68: eigenDecomposition of [[2,-1,0],[-1,2,-1],[0,-1,2]]
    {
        const std::vector<double> R = {2, -1, 0, -1, 2, -1, 0, -1, 2};
        auto [U, D] = eigenDecomposition(R, 3);
        const double expect[3] = {3.41421356, 2.0, 0.58578644};
        if (D.size() != 3) return fail("eigenDecomposition: expected 3 eigenvalues");
        for (int i = 0; i < 3; ++i)
            if (std::fabs(D[i] - expect[i]) > 1e-5) return fail("eigenDecomposition: eigenvalue " + std::to_string(i));
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double dot = 0, rec = 0;
                for (int i = 0; i < 3; ++i) { dot += U[i + 3 * a] * U[i + 3 * b]; rec += U[a + 3 * i] * D[i] * U[b + 3 * i]; }
                if (std::fabs(dot - (a == b)) > 1e-10) return fail("eigenDecomposition: U^T U != I");
                if (std::fabs(rec - R[a + 3 * b]) > 1e-10) return fail("eigenDecomposition: U D U^T != R");
            }
    }
    std::vector<uint8_t> bgr((size_t)rows * cols * 3);
    {
        std::ifstream f(in_path, std::ios::binary);
        if (!f.read(reinterpret_cast<char*>(bgr.data()), (std::streamsize)bgr.size())) return fail("cannot read " + in_path);
    }
    const ImageView img{bgr.data(), rows, cols};
    // error behaviour of the reference (filter.cpp:117-119): more samples than pixels along an axis
    try {
        NLEFilter bad;
        bad.trainForEnhancement(img, rows + 1, nCS, hx, hy, T, K);
        return fail("expected std::runtime_error for nRowSamples > rows");
    } catch (const std::runtime_error&) {}
    try {
        NLEFilter untrained;
        untrained.enhance(img, weights);
        return fail("expected std::runtime_error for an untrained filter");
    } catch (const std::runtime_error&) {}

    NLEFilter filter;                                                      // enhance.cpp:39-44
    filter.trainForEnhancement(img, nRS, nCS, hx, hy, T, K);
    NLEFilter copy = filter;                                               // NLEFilter is copyable (enhance.cpp:39)
    const std::vector<uint8_t> out = copy.enhance(img, weights);
    {
        std::ofstream f(out_path, std::ios::binary);
        f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
    }
    // a second weight set on the same trained filter (train once, enhance many)
    const std::vector<uint8_t> out2 = filter.enhance(img, std::vector<double>(weights.size(), 1.0));
    size_t changed = 0;
    for (size_t i = 0; i < out.size(); ++i) changed += out[i] != out2[i];
    const auto& inf = filter.info();
    const auto sel = samplePixels(rows, cols, nRS, nCS);
    if ((int)sel.size() != inf.p) return fail("samplePixels count != filter p");
    std::printf("ok p=%d r=%d r2=%d k=%d S0=%.9f differing_bytes_vs_unit_weights=%zu\n", inf.p, inf.r, inf.r2, inf.k,
                filter.eigvals()[0], changed);
    return 0;
}
