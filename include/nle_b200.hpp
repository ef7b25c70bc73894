// nle_b200.hpp -- C++ host mirror of the reference's include/filter.hpp on top of the C ABI (nle_b200.h).
//
// The reference is C++ (class nle::NLEFilter, filter.hpp:35-54; free functions filter.hpp:20-33) over cv::Mat and
// Eigen types.  Neither OpenCV's C++ SDK nor Eigen exists in this image, so this mirror keeps the reference's names,
// argument order and error behaviour (std::runtime_error with the reference's messages, filter.cpp:118,415,419,448) but
// speaks plain buffers: an image is a (rows, cols, pointer) view of interleaved 8-bit BGR -- exactly cv::Mat::data of
// a continuous CV_8UC3 -- and matrices are column-major std::vector<double> (Eigen's default layout, filter.hpp:10-11).
// With OpenCV/Eigen present the replacement translation unit of INTEGRATION.md is a thin adapter over this class.
// Header-only; link with libnle_b200.so.  There is no CPU fallback: without a CUDA device every call throws.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "nle_b200.h"

namespace nle_b200 {

using DType = double;                                   // filter.hpp:12
constexpr DType EPS = 1e-10;                            // filter.hpp:14

inline void ok(int rc) {
    if (rc != NLE_B200_OK) throw std::runtime_error(nle_b200_last_error());
}

struct ImageView {                                      // continuous CV_8UC3 cv::Mat: rows x cols x 3, BGR
    const uint8_t* data;
    int rows, cols;
    size_t total() const { return (size_t)rows * cols; }
};

// nle::eigenDecomposition (filter.cpp:204-228): U (n x r, column-major), D (r), descending, D >= eps
inline std::pair<std::vector<double>, std::vector<double>> eigenDecomposition(const std::vector<double>& M, int n, DType eps = EPS) {
    std::vector<double> U((size_t)n * n), D(n);
    int r = 0;
    ok(nle_b200_eigen_decomposition(M.data(), n, eps, U.data(), D.data(), &r));
    U.resize((size_t)n * r);
    D.resize(r);
    return {std::move(U), std::move(D)};
}

// transformEigenValues (filter.cpp:334-347)
inline std::vector<double> transformEigenValues(const std::vector<double>& S, const std::vector<DType>& weights) {
    std::vector<double> fS(S.size());
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

    // NLEFilter::trainFilter on an 8-bit luminance channel (filter.cpp:480-502)
    void trainFilter(const uint8_t* lum, int rows, int cols, int nRowSamples, int nColSamples, DType hx, DType hy,
                     int nSinkhornIter, int nEigenVectors) {
        nle_b200_filter* f = nullptr;
        ok(nle_b200_train_u8(lum, rows, cols, nRowSamples, nColSamples, hx, hy, nSinkhornIter, nEigenVectors, &f));
        adopt(f);
    }

    // filter.cpp:412-443: returns the enhanced BGR image (rows*cols*3 bytes)
    std::vector<uint8_t> enhance(const ImageView& image, const std::vector<DType>& weights) const {
        require();
        if ((long long)image.total() != (long long)m_info.rows * m_info.cols)
            throw std::runtime_error("Cannot apply filter on image with different size from the image filter was trained on.");   // :419
        std::vector<uint8_t> out(image.total() * 3);
        ok(nle_b200_enhance_bgr_u8(m_h.get(), image.data, weights.data(), (int)weights.size(), out.data()));
        return out;
    }

    // NLEFilter::apply (filter.cpp:445-458): V diag(fS) V^T channel
    std::vector<double> apply(const std::vector<double>& channel, const std::vector<double>& fS) const {
        require();
        if (channel.size() != (size_t)m_info.rows * m_info.cols)
            throw std::runtime_error("Number of values in channel must match that of training image.");                          // :448
        std::vector<double> out(channel.size());
        ok(nle_b200_apply(m_h.get(), channel.data(), fS.data(), out.data()));
        return out;
    }

    const std::vector<double>& eigvals() const { return m_eigvals; }          // m_eigvals, filter.hpp:53
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
    std::vector<double> m_eigvals;
    nle_b200_info m_info{};
};

}  // namespace nle_b200
