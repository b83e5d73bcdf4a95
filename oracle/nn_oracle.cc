/*
 * nn_oracle.cc -- CPU restatement of the reference's legacy feed-forward network forward pass and
 * of Nn::BatchFeatureScorer's score definition.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows: NeuralNetwork<T>::forward / forwardLayers src/Nn/NeuralNetwork.cc:313-331,409-425;
 * LinearLayer<T>::_forward src/Nn/LinearLayer.cc:298-321 (C = W^T X with beta 0, then
 * addToAllColumns(bias)); activations: FastMatrix::sigmoid src/Math/FastMatrix.hh:802-808,
 * softmax :820-836, ensureMinimalValue(0) (RectifiedLayer src/Nn/ActivationLayer.cc:272-282),
 * tanh; score definition src/Nn/BatchFeatureScorer.cc:148-171 with
 * BiasLayer::removeLogPriorFromBias src/Nn/LinearLayer.cc:499-519.
 *
 * The dense product itself lives in a third-party BLAS the reference does not pin
 * (cblas_sgemm via src/Math/Blas.hh:410-421, any system BLAS; or cublasSgemm via
 * src/Math/CublasWrapper.hh:320-331).  It is restated as the textbook contraction with a
 * sequential f32 accumulator (mode F32), with an f64 accumulator to bound re-ordering (F64ACC),
 * and with operands rounded to bf16 / f32 accumulation (BF16) as the comparison target of the
 * tensor-core path.  Pinned by the reference's own unit-test vectors (tests/test_oracle_nn.py).
 */
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

inline float bf16Round(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u)
        return x; /* inf / nan unchanged */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    u &= 0xffff0000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

template<typename T>
void activate(int act, T* col, int n) {
    switch (act) {
        case ORC_ACT_LINEAR: break;
        case ORC_ACT_SIGMOID:
            /* scale(-gamma); exp(); 1.0 / (1.0 + e)  (gamma = 1) */
            for (int i = 0; i < n; ++i) {
                T e    = std::exp(col[i] * (T)-1);
                col[i] = 1.0 / (1.0 + e);
            }
            break;
        case ORC_ACT_RELU:
            for (int i = 0; i < n; ++i)
                if (col[i] < (T)0)
                    col[i] = (T)0;
            break;
        case ORC_ACT_TANH:
            for (int i = 0; i < n; ++i)
                col[i] = std::tanh(col[i]);
            break;
        case ORC_ACT_SOFTMAX: {
            /* subtract column max, exp, divide by the column sum */
            T mx = col[0];
            for (int i = 1; i < n; ++i)
                mx = std::max(mx, col[i]);
            T sum = 0;
            for (int i = 0; i < n; ++i) {
                col[i] = std::exp(col[i] + (T)-1.0 * mx);
                sum += col[i];
            }
            for (int i = 0; i < n; ++i)
                col[i] = col[i] / sum;
            break;
        }
    }
}

template<typename T, typename Acc>
void forwardImpl(int nLayers, const int* dims, const int* act, const T* const* W, const T* const* bias,
                 const T* lastBiasOverride, const T* x, long nFrames, T* out, bool bf16) {
    int maxDim = 0;
    for (int l = 0; l <= nLayers; ++l)
        maxDim = std::max(maxDim, dims[l]);
    std::vector<T> a(maxDim), b(maxDim);
    std::vector<std::vector<T>> Wr;
    if (bf16) {
        Wr.resize(nLayers);
        for (int l = 0; l < nLayers; ++l) {
            size_t n = (size_t)dims[l] * dims[l + 1];
            Wr[l].resize(n);
            for (size_t i = 0; i < n; ++i)
                Wr[l][i] = (T)bf16Round((float)W[l][i]);
        }
    }
    for (long t = 0; t < nFrames; ++t) {
        std::copy(x + (size_t)t * dims[0], x + (size_t)(t + 1) * dims[0], a.begin());
        for (int l = 0; l < nLayers; ++l) {
            const int in = dims[l], on = dims[l + 1];
            const T*  w  = bf16 ? Wr[l].data() : W[l];
            const T*  bs = (l == nLayers - 1 && lastBiasOverride) ? lastBiasOverride : bias[l];
            if (bf16)
                for (int i = 0; i < in; ++i)
                    a[i] = (T)bf16Round((float)a[i]);
            for (int o = 0; o < on; ++o) {
                Acc      acc = 0;
                const T* wo  = w + (size_t)o * in;
                for (int i = 0; i < in; ++i)
                    acc += (Acc)wo[i] * (Acc)a[i];
                T v = (T)acc;
                if (bs)
                    v += bs[o]; /* axpy with alpha 1 */
                b[o] = v;
            }
            activate<T>(act[l], b.data(), on);
            std::swap(a, b);
        }
        std::copy(a.begin(), a.begin() + dims[nLayers], out + (size_t)t * dims[nLayers]);
    }
}

}  // namespace

extern "C" int orc_nn_forward(int n_layers, const int* dims, const int* act, const float* const* weights,
                              const float* const* bias, const float* x, long T, float* out, int mode) {
    if (n_layers < 1)
        return -1;
    if (mode == ORC_NN_F64ACC)
        forwardImpl<float, double>(n_layers, dims, act, weights, bias, 0, x, T, out, false);
    else
        forwardImpl<float, float>(n_layers, dims, act, weights, bias, 0, x, T, out, mode == ORC_NN_BF16);
    return 0;
}

extern "C" int orc_nn_forward_f64(int n_layers, const int* dims, const int* act, const double* const* weights,
                                  const double* const* bias, const double* x, long T, double* out) {
    if (n_layers < 1)
        return -1;
    forwardImpl<double, double>(n_layers, dims, act, weights, bias, 0, x, T, out, false);
    return 0;
}

extern "C" int orc_nn_scores(int n_layers, const int* dims, const int* act, const float* const* weights,
                             const float* const* bias, const float* log_prior, float prior_scale, const float* x,
                             long T, float* scores, int mode) {
    if (n_layers < 1)
        return -1;
    const int          nOut = dims[n_layers];
    std::vector<float> topBias(bias[n_layers - 1], bias[n_layers - 1] + nOut);
    /* bias[c] -= prioriScale * prior[c]   (CPU branch of removeLogPriorFromBias) */
    if (log_prior && prior_scale != 0.0f)
        for (int c = 0; c < nOut; ++c)
            topBias[c] -= prior_scale * log_prior[c];
    std::vector<int> a(act, act + n_layers);
    if (a[n_layers - 1] == ORC_ACT_SOFTMAX)
        a[n_layers - 1] = ORC_ACT_LINEAR; /* topLayer->setEvaluateSoftmax(false) */
    if (mode == ORC_NN_F64ACC)
        forwardImpl<float, double>(n_layers, dims, a.data(), weights, bias, topBias.data(), x, T, scores, false);
    else
        forwardImpl<float, float>(n_layers, dims, a.data(), weights, bias, topBias.data(), x, T, scores,
                                  mode == ORC_NN_BF16);
    const size_t n = (size_t)T * nOut;
    for (size_t i = 0; i < n; ++i)
        scores[i] = -scores[i];
    return 0;
}
