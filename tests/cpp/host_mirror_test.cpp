// C++ host-side exercise of the C ABI through include/nle_b200.hpp (no Python, no OpenCV, no Eigen):
//   host_mirror_test <bgr.raw> <rows> <cols> <nRowSamples> <nColSamples> <hx> <hy> <nSinkhornIter> <nEigenVectors> <out.raw> w0 w1 ...
// -- the argument order of the reference's `enhance` CLI (enhance.cpp:20-31) with raw interleaved BGR files in place of
// cv::imread / cv::imwrite.  Before that it runs the reference's Catch2 cases through the C++ free functions of
// nle_b200.hpp with the reference's tolerance (test/test_filter.cpp:8, tol = 1e-10; isApprox = relative Frobenius):
//   "Eigen Decomposition" :42-68, "Sinkhorn" :70-122 (identity and balanced random matrix), "Orthogonalize" :126-153
// (Mat::Random is unseeded in the reference; here a fixed LCG), then the error paths of enhance / denoise / apply
// with the reference's messages (filter.cpp:118,352,356,415,419,448) and denoise end to end.
// Prints "ok ..." and exits 0 on success.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>

#include "nle_b200.hpp"

using namespace nle_b200;

static int fail(const std::string& m) { std::cerr << "FAIL: " << m << std::endl; return 1; }

static const double tol = 1e-10;                                           // test_filter.cpp:8
static double fro(const std::vector<double>& a) { double s = 0; for (double x : a) s += x * x; return std::sqrt(s); }
// Eigen's isApprox: ||a - b||_F <= prec * min(||a||_F, ||b||_F)
static bool is_approx(const std::vector<double>& a, const std::vector<double>& b, double prec) {
    if (a.size() != b.size()) return false;
    std::vector<double> d(a.size());
    for (size_t i = 0; i < a.size(); ++i) d[i] = a[i] - b[i];
    return fro(d) <= prec * std::min(fro(a), fro(b));
}
static Mat transpose(const Mat& A) { Mat T(A.cols, A.rows); for (int i = 0; i < A.rows; ++i) for (int j = 0; j < A.cols; ++j) T(j, i) = A(i, j); return T; }
static Mat gram(const Mat& V) {                                            // V^T V
    Mat G(V.cols, V.cols);
    for (int a = 0; a < V.cols; ++a) for (int b = 0; b < V.cols; ++b) { double s = 0; for (int i = 0; i < V.rows; ++i) s += V(i, a) * V(i, b); G(a, b) = s; }
    return G;
}
struct Lcg {                                                               // stands in for Mat::Random of the reference
    unsigned long long x;
    double next() { x = x * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(x >> 11) / 9007199254740992.0; }
    Mat mat(int r, int c) { Mat A(r, c); for (int j = 0; j < c; ++j) for (int i = 0; i < r; ++i) A(i, j) = next(); return A; }
};
// rows / cols of [Wa Wab] sum to one (test_filter.cpp:80-93, 110-121)
static bool sums_to_one(const Mat& Wa, const Mat& Wab) {
    std::vector<double> rs(Wa.rows, 0.0), cs(Wa.cols, 0.0), ones_r(Wa.rows, 1.0), ones_c(Wa.cols, 1.0);
    for (int i = 0; i < Wa.rows; ++i) {
        for (int j = 0; j < Wa.cols; ++j) { rs[i] += Wa(i, j); cs[j] += Wa(i, j); }
        for (int j = 0; j < Wab.cols; ++j) rs[i] += Wab(i, j);
    }
    for (int j = 0; j < Wa.cols; ++j) for (int q = 0; q < Wab.cols; ++q) cs[j] += Wab(j, q);      // [Wa; Wab^T] column sums
    return is_approx(rs, ones_r, tol) && is_approx(cs, ones_c, tol);
}

static int catch2_cases() {
    // "Eigen Decomposition", test_filter.cpp:42-68
    {
        Mat R(3, 3);
        const double r[9] = {2, -1, 0, -1, 2, -1, 0, -1, 2};
        R.a.assign(r, r + 9);
        auto [U, D] = eigenDecomposition(R, tol);
        if (D.size() != 3 || !is_approx(D, {3.41421356, 2.0, 0.58578644}, 1e-5)) return fail("eigenDecomposition: eigenvalues");
        Mat rec(3, 3);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int i = 0; i < 3; ++i) rec(a, b) += U(a, i) * D[i] * U(b, i);
        if (!is_approx(rec.a, R.a, tol)) return fail("eigenDecomposition: U D U^T != R");
        Mat I(3, 3); for (int i = 0; i < 3; ++i) I(i, i) = 1.0;
        if (!is_approx(I.a, gram(U).a, tol)) return fail("eigenDecomposition: U^T U != I");
    }
    // "Sinkhorn", test_filter.cpp:70-94: phi = I_2, eigvals = 1, 10 iterations
    {
        Mat I(2, 2); I(0, 0) = I(1, 1) = 1.0;
        auto [Wa, Wab] = sinkhorn(I, Vec{1.0, 1.0}, 10);
        if (Wab.cols != 0) return fail("sinkhorn(I): Wab must be 2 x 0");
        if (!is_approx(Wa.a, transpose(Wa).a, 1e-12)) return fail("sinkhorn(I): Wa not symmetric");
        if (!sums_to_one(Wa, Wab)) return fail("sinkhorn(I): row / column sums");
    }
    Lcg rng{5};
    // "Balanced random matrix", test_filter.cpp:96-122
    {
        Mat R = rng.mat(5, 5);
        auto [U, D] = eigenDecomposition(R, tol);
        if (D.empty()) return fail("balanced random: no positive eigenvalue");
        auto [Wa, Wab] = sinkhorn(U, D, 20);
        if (Wa.rows != (int)D.size() || Wab.cols != 5 - (int)D.size()) return fail("balanced random: shapes");
        if (!is_approx(Wa.a, transpose(Wa).a, tol)) return fail("balanced random: Wa not symmetric");
        if (!sums_to_one(Wa, Wab)) return fail("balanced random: row / column sums");
    }
    // "Orthogonalize", test_filter.cpp:126-153
    {
        const int p = 10, n = 100, k = 5;
        Mat Wa = rng.mat(p, p);
        Mat WaT = transpose(Wa);
        for (size_t i = 0; i < Wa.a.size(); ++i) Wa.a[i] = (Wa.a[i] + WaT.a[i]) / 2;
        Mat Wab = rng.mat(p, n - p);
        auto [V, S] = orthogonalize(Wa, Wab, k);
        if (S.empty() || V.cols == 0 || (int)S.size() != V.cols) return fail("orthogonalize: counts");
        Mat I(V.cols, V.cols); for (int i = 0; i < V.cols; ++i) I(i, i) = 1.0;
        if (V.rows != n || !is_approx(gram(V).a, I.a, tol)) return fail("orthogonalize: V^T V != I");
    }
    return 0;
}

static bool throws_with(const std::function<void()>& fn, const std::string& msg) {
    try { fn(); } catch (const std::runtime_error& e) { return msg == e.what(); }
    return false;
}

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

    if (int rc = catch2_cases()) return rc;
    std::vector<uint8_t> bgr((size_t)rows * cols * 3);
    {
        std::ifstream f(in_path, std::ios::binary);
        if (!f.read(reinterpret_cast<char*>(bgr.data()), (std::streamsize)bgr.size())) return fail("cannot read " + in_path);
    }
    const ImageView img{bgr.data(), rows, cols};
    // error behaviour of the reference (filter.cpp:117-119): more samples than pixels along an axis
    if (!throws_with([&] { NLEFilter bad; bad.trainForEnhancement(img, rows + 1, nCS, hx, hy, T, K); },
                     "Number of samples per row and col must be <= that of image."))                      // filter.cpp:118
        return fail("expected the reference's runtime_error for nRowSamples > rows");
    try {
        NLEFilter untrained;
        untrained.enhance(img, weights);
        return fail("expected std::runtime_error for an untrained filter");
    } catch (const std::runtime_error&) {}

    NLEFilter filter;                                                      // enhance.cpp:39-44
    filter.trainForEnhancement(img, nRS, nCS, hx, hy, T, K);
    NLEFilter copy = filter;                                               // NLEFilter is copyable (enhance.cpp:39)
    const std::vector<uint8_t> out = copy.enhance(img, weights);
    // the reference's checks, raised by the C ABI itself with the reference's messages
    const ImageView gray{bgr.data(), rows, cols, 1}, shorter{bgr.data(), rows - 1, cols, 3};
    const std::string size_msg = "Cannot apply filter on image with different size from the image filter was trained on.";
    if (!throws_with([&] { filter.enhance(gray, weights); }, "Can only enhance RGB image.")) return fail("enhance: :415 message");
    if (!throws_with([&] { filter.enhance(shorter, weights); }, size_msg)) return fail("enhance: :419 message");
    const BilateralFn copy_filter = [](const uint8_t* src, uint8_t* dst, int r, int c, int, int) { std::copy(src, src + (size_t)r * c, dst); };
    if (!throws_with([&] { filter.denoise(gray, 2.0, copy_filter); }, "Can only enchance RGB image.")) return fail("denoise: :352 message");
    if (!throws_with([&] { filter.denoise(shorter, 2.0, copy_filter); }, size_msg)) return fail("denoise: :356 message");
    if (!throws_with([&] { filter.apply(std::vector<double>(7, 0.0), filter.eigvals()); },
                     "Number of values in channel must match that of training image."))
        return fail("apply: :448 message");
    // denoise end to end (bilateralFilter replaced by a copy: OpenCV's C++ SDK is not in this image) and trainForDenoise
    const std::vector<uint8_t> den = filter.denoise(img, 2.0, copy_filter);
    if (den.size() != bgr.size()) return fail("denoise: output size");
    NLEFilter dfilter;
    dfilter.trainForDenoise(img, nRS, nCS, hx, hy, T, K, copy_filter);
    if (dfilter.info().p != filter.info().p || dfilter.eigvals().size() != filter.eigvals().size()) return fail("trainForDenoise: shape of the trained filter");
    const Mat Vm = filter.eigvecs();
    if (Vm.rows != rows * cols || Vm.cols != filter.info().k) return fail("eigvecs: shape");
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
