// nle_b200.hpp -- C++ host mirror of the reference's include/filter.hpp on top of the C ABI (nle_b200.h).
//
// The reference is C++ (class nle::NLEFilter, filter.hpp:35-54; free functions filter.hpp:20-33) over cv::Mat and
// Eigen types.  Neither OpenCV's C++ SDK nor Eigen exists in this image, so this mirror keeps the reference's names,
// argument order, defaults and error behaviour (std::runtime_error with the reference's messages,
// filter.cpp:118,352,356,415,419,448) but speaks plain buffers: an image is a (rows, cols, channels, pointer) view of
// interleaved 8-bit data -- exactly cv::Mat::data of a continuous CV_8UC3 -- and matrices are column-major
// (rows, cols, std::vector<double>) -- Eigen's default layout, filter.hpp:10-11.
// With OpenCV/Eigen present the replacement translation unit src/filter_b200.cpp is a thin adapter
// over the same C calls.  Header-only; link with libnle_b200.so.  There is no CPU fallback: without a CUDA device
// every call throws.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <utility>
#include <vector>

#include "nle_b200.h"

namespace nle_b200 {

using DType = double;                                   // filter.hpp:12
constexpr DType EPS = 1e-10;                            // filter.hpp:14

inline void ok(int rc) {
    if (rc != NLE_B200_OK) throw std::runtime_error(nle_b200_last_error());
}

struct ImageView {                                      // continuous 8-bit cv::Mat: rows x cols x channels (BGR when 3)
    const uint8_t* data;
    int rows, cols;
    int channels = 3;
    size_t total() const { return (size_t)rows * cols; }   // cv::Mat::total()
};

struct Mat {                                            // Eigen::MatrixXd stand-in: column-major
    int rows = 0, cols = 0;
    std::vector<double> a;
    Mat() = default;
    Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * rows]; }
    double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * rows]; }
};
using Vec = std::vector<double>;                        // Eigen::VectorXd stand-in

// nle::computeKernel (filter.cpp:114-167): (P.indices(), Ka p x p, Kab p x (N-p)); mat = luminance as doubles (raster)
inline std::tuple<std::vector<int32_t>, Mat, Mat> computeKernel(const std::vector<double>& mat, int rows, int cols,
                                                                int nRowSamples, int nColSamples, DType hx, DType hy) {
    int p = 0;
    ok(nle_b200_sample_count(rows, cols, nRowSamples, nColSamples, &p));        // throws filter.cpp:118's message
    const long long N = (long long)rows * cols;
    std::vector<int32_t> perm((size_t)N);
    Mat Ka(p, p), Kab(p, (int)(N - p));
    ok(nle_b200_compute_kernel(mat.data(), rows, cols, nRowSamples, nColSamples, hx, hy, perm.data(), Ka.a.data(),
                               Kab.a.data()));
    return {std::move(perm), std::move(Ka), std::move(Kab)};
}

// nle::eigenDecomposition (filter.cpp:204-228): U (n x r), D (r), descending, D >= eps; reads the lower triangle
inline std::pair<Mat, Vec> eigenDecomposition(const Mat& M, DType eps = EPS) {
    const int n = M.rows;
    Mat U(n, n);
    Vec D(n);
    int r = 0;
    ok(nle_b200_eigen_decomposition(M.a.data(), n, eps, U.a.data(), D.data(), &r));
    U.a.resize((size_t)n * r);
    U.cols = r;
    D.resize(r);
    return {std::move(U), std::move(D)};
}

// nle::topkEigenDecomposition (filter.cpp:169-200, USE_SPECTRA build): the min(nLargest, n-1) eigenpairs of largest magnitude,
// descending, cut at the first eigenvalue below eps.  assume_psd (true for the reference's only caller, Q at :311) selects the
// block solver of the training path; the result is the same either way.
inline std::pair<Mat, Vec> topkEigenDecomposition(const Mat& M, int nLargest, DType eps = EPS, bool assume_psd = false) {
    const int n = M.rows;
    const int nev = std::max(1, std::min(nLargest, n - 1));
    Mat U(n, nev);
    Vec D(nev);
    int r = 0;
    ok(nle_b200_topk_eigen_decomposition(M.a.data(), n, nLargest, eps, assume_psd ? 1 : 0, U.a.data(), D.data(), &r, nullptr));
    U.a.resize((size_t)n * r);
    U.cols = r;
    D.resize(r);
    return {std::move(U), std::move(D)};
}

// nle::nystromApproximation (filter.cpp:257-280): (eigvals r, phi (p+nrest) x r)
inline std::pair<Vec, Mat> nystromApproximation(const Mat& Ka, const Mat& Kab) {
    const int p = Ka.rows, nrest = Kab.cols;
    Vec lam(p);
    Mat phi(p + nrest, p);
    int r = 0;
    ok(nle_b200_nystrom_approximation(Ka.a.data(), p, Kab.a.data(), nrest, lam.data(), phi.a.data(), &r));
    lam.resize(r);
    phi.a.resize((size_t)(p + nrest) * r);
    phi.cols = r;
    return {std::move(lam), std::move(phi)};
}

// nle::sinkhorn (filter.cpp:230-254): (Wa r x r, Wab r x (n-r)) with r = phi.cols() (:247)
inline std::pair<Mat, Mat> sinkhorn(const Mat& phi, const Vec& eigvals, int maxIter = 10) {
    const int n = phi.rows, r = phi.cols;
    Mat Wa(r, r), Wab(r, n - r);
    ok(nle_b200_sinkhorn(phi.a.data(), n, r, eigvals.data(), maxIter, Wa.a.data(), n > r ? Wab.a.data() : nullptr));
    return {std::move(Wa), std::move(Wab)};
}

// nle::orthogonalize (filter.cpp:282-331): (V (p+nrest) x k', S k')
inline std::pair<Mat, Vec> orthogonalize(const Mat& Wa, const Mat& Wab, int nEigVectors = 5, DType eps = EPS) {
    const int p = Wa.rows, nrest = Wab.cols;
    Mat V(p + nrest, nEigVectors);
    Vec S(nEigVectors);
    int k = 0;
    ok(nle_b200_orthogonalize(Wa.a.data(), p, nrest ? Wab.a.data() : nullptr, nrest, nEigVectors, eps, V.a.data(), S.data(), &k));
    V.a.resize((size_t)(p + nrest) * k);
    V.cols = k;
    S.resize(k);
    return {std::move(V), std::move(S)};
}

// transformEigenValues (filter.cpp:334-347)
inline Vec transformEigenValues(const Vec& S, const std::vector<DType>& weights) {
    Vec fS(S.size());
    ok(nle_b200_transform_eigenvalues(S.data(), (int)S.size(), weights.data(), (int)weights.size(), fS.data()));
    return fS;
}

// sample indices of samplePixels + to1DIndex (filter.cpp:56-80, utils.hpp:11-14)
inline std::vector<int32_t> samplePixels(int rows, int cols, int nRowSamples, int nColSamples) {
    int p = 0;
    ok(nle_b200_sample_count(rows, cols, nRowSamples, nColSamples, &p));
    std::vector<int32_t> sel(p);
    ok(nle_b200_sample_indices(rows, cols, nRowSamples, nColSamples, sel.data(), nullptr));
    return sel;
}

// cv::bilateralFilter(src, dst, -1, sigmaColor, sigmaSpace, BORDER_DEFAULT) on one 8-bit channel.  The reference calls
// OpenCV for it (filter.cpp:366-371, 535) and so does every host of this library; the plain-buffer mirror takes it as a
// callable because OpenCV's C++ SDK is not in this image.
using BilateralFn = std::function<void(const uint8_t* src, uint8_t* dst, int rows, int cols, int sigmaColor, int sigmaSpace)>;

class NLEFilter {                                       // nle::NLEFilter, filter.hpp:35-54
public:
    NLEFilter() = default;

    // filter.cpp:514-519 (getLuminanceChannel :460-469 + trainFilter :480-502); BGR2Lab runs on the device
    void trainForEnhancement(const ImageView& image, int nRowSamples, int nColSamples, DType hx, DType hy,
                             int nSinkhornIter = 10, int nEigenVectors = 5) {
        nle_b200_filter* f = nullptr;
        ok(nle_b200_train_bgr_u8(image.data, image.rows, image.cols, 0, image.rows, nRowSamples, nColSamples, hx, hy,
                                 nSinkhornIter, nEigenVectors, nullptr, nullptr, &f));
        adopt(f);
    }

    // filter.cpp:521-538: BGR2Lab, bilateralFilter(L), trainFilter on the filtered luminance
    void trainForDenoise(const ImageView& image, int nRowSamples, int nColSamples, DType hx, DType hy, int nSinkhornIter,
                         int nEigenVectors, const BilateralFn& bilateral, int sigmaColor = 10, int sigmaSpace = 10) {
        const size_t n = image.total();
        std::vector<uint8_t> lab(3 * n), L(n), den(n);
        ok(nle_b200_bgr_to_lab_u8(image.data, (long long)n, lab.data()));                       // :528
        for (size_t j = 0; j < n; ++j) L[j] = lab[3 * j];
        bilateral(L.data(), den.data(), image.rows, image.cols, sigmaColor, sigmaSpace);        // :535
        trainFilter(den.data(), image.rows, image.cols, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors);
    }

    // NLEFilter::trainFilter on an 8-bit luminance channel (filter.cpp:480-502)
    void trainFilter(const uint8_t* lum, int rows, int cols, int nRowSamples, int nColSamples, DType hx, DType hy,
                     int nSinkhornIter, int nEigenVectors) {
        nle_b200_filter* f = nullptr;
        ok(nle_b200_train_u8(lum, rows, cols, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors, &f));
        adopt(f);
    }

    // filter.cpp:412-443: returns the enhanced BGR image (rows*cols*3 bytes).  The two checks of :414-420 are raised by
    // the C ABI itself, with the reference's messages.
    std::vector<uint8_t> enhance(const ImageView& image, const std::vector<DType>& weights) const {
        require();
        std::vector<uint8_t> out(image.total() * 3);
        ok(nle_b200_enhance_bgr_u8(m_h.get(), image.data, image.rows, image.cols, image.channels, weights.data(),
                                   (int)weights.size(), out.data()));
        return out;
    }

    // filter.cpp:349-410 without the imshow side effects: L <- bilateral(L) unfiltered by V, a and b <- V f(S) V^T (.)
    std::vector<uint8_t> denoise(const ImageView& image, DType k, const BilateralFn& bilateral, int sigmaColor = 10,
                                 int sigmaSpace = 10) const {
        if (image.channels != 3) throw std::runtime_error("Can only enchance RGB image.");       // :351-353 (sic)
        require();
        if ((long long)image.total() != (long long)m_info.rows * m_info.cols)                    // :355-357
            throw std::runtime_error("Cannot apply filter on image with different size from the image filter was trained on.");
        const size_t n = image.total();
        std::vector<uint8_t> lab(3 * n), ch(n), res(n), out(3 * n);
        ok(nle_b200_bgr_to_lab_u8(image.data, (long long)n, lab.data()));                       // :361
        for (size_t j = 0; j < n; ++j) ch[j] = lab[3 * j];
        bilateral(ch.data(), res.data(), image.rows, image.cols, sigmaColor, sigmaSpace);       // :371
        for (size_t j = 0; j < n; ++j) out[3 * j] = res[j];                                      // :387 is commented out
        for (int c = 1; c <= 2; ++c) {                                                           // :388-399
            for (size_t j = 0; j < n; ++j) ch[j] = lab[3 * j + c];
            ok(nle_b200_denoise_channel_u8(m_h.get(), ch.data(), image.rows, image.cols, k, res.data()));
            for (size_t j = 0; j < n; ++j) out[3 * j + c] = res[j];
        }
        std::vector<uint8_t> bgr(3 * n);
        ok(nle_b200_lab_to_bgr_u8(out.data(), (long long)n, bgr.data()));                       // :408
        return bgr;
    }

    // NLEFilter::apply (filter.cpp:445-458): V diag(fS) V^T channel; the size check of :447-449 is raised by the C ABI
    std::vector<double> apply(const std::vector<double>& channel, const Vec& fS) const {
        require();
        std::vector<double> out(channel.size());
        ok(nle_b200_apply(m_h.get(), channel.data(), (long long)channel.size(), fS.data(), out.data()));
        return out;
    }

    const Vec& eigvals() const { return m_eigvals; }                          // m_eigvals, filter.hpp:53
    Mat eigvecs() const {                                                     // m_eigvecs, filter.hpp:52 (N x k copy)
        require();
        Mat V((m_info.row1 - m_info.row0) * m_info.cols, m_info.k);
        ok(nle_b200_eigenvectors(m_h.get(), V.a.data()));
        return V;
    }
    const nle_b200_info& info() const { return m_info; }
    bool trained() const { return (bool)m_h; }

private:
    struct Deleter { void operator()(nle_b200_filter* f) const { nle_b200_free(f); } };
    void adopt(nle_b200_filter* f) {
        m_h.reset(f, Deleter());                        // copies of the object share the device-resident eigenvectors
        ok(nle_b200_filter_info(f, &m_info));
        m_eigvals.resize(m_info.k);
        ok(nle_b200_eigenvalues(f, m_eigvals.data()));
    }
    void require() const { if (!m_h) throw std::runtime_error("NLEFilter: filter has not been trained"); }
    std::shared_ptr<nle_b200_filter> m_h;               // m_eigvecs (N x k) stays in HBM behind this handle
    Vec m_eigvals;
    nle_b200_info m_info{};
};

}  // namespace nle_b200
