/*
 * frontend_oracle.cc -- CPU restatement of the reference's MFCC Flow pipeline.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Each block names the reference lines it follows.  PARITY PINNED: bit for
 * bit against Flow networks built by the reference's own NetworkParser from its own .flow files (oracle/_ref,
 * tests/test_ref_parity.py: every stage, 14 boundary lengths, packet sizes, sample rates, window types, DC detection).
 * The network restated is src/Tools/FeatureExtraction/share/mfcc.flow:8-34 followed by
 * derivationWithRegression.flow:7-27 and a generic-vector-f32-concat of static|delta|delta-delta.
 *
 * Build with -ffp-contract=off: contraction is applied explicitly (cfg.use_fma) at exactly the
 * places where the reference's default build (gcc -O2 -march=native, GNU mode => -ffp-contract=fast;
 * cmake_resources/CompileOptions.cmake:39-48, ConfigurationTypes.cmake:13-14) fuses a multiply-add.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <sstream>
#include <vector>

namespace {

typedef double Time;

/* Flow attributes travel as text: Attributes::set(name, f64) prints with default ostream precision
 * and consumers atof() them back (src/Flow/Attributes.hh:104-113, src/Signal/Filterbank.cc:852). */
double attributeRoundTrip(double v) {
    std::ostringstream s;
    s << v;
    return atof(s.str().c_str());
}

inline float mulAdd(float a, float b, float c, bool fuse) {
    return fuse ? std::fmaf(a, b, c) : (a * b + c);
}

/* ---- Core::isAlmostEqual(f64,f64) src/Core/Utility.hh:322-327, constants src/Core/Types.hh:153-159 */
bool almostEqual(double a, double b) {
    const double eps = 2.2204460492503131e-16, delta = 2.2250738585072014e-308;
    double       d   = std::fabs(a - b);
    double       e   = (std::fabs(a) + std::fabs(b) + delta) * eps;
    return d < e;
}

/* ---- FilterBank::isAlmostInteger src/Signal/Filterbank.cc:690-693 */
bool almostInteger(double x) {
    return std::fabs(x - std::round(x)) < 1e-10;
}

struct Geometry {
    double sampleRate;  // what the nodes read back from the "sample-rate" attribute
    int    L, S, N;     // window length / shift in samples, FFT points
    double fftOutRate;  // "sample-rate" attribute after the FFT node (N / sampleRate, as text)
};

Geometry makeGeometry(const orc_frontend_cfg& c) {
    Geometry g;
    g.sampleRate = attributeRoundTrip(c.sample_rate);
    /* Window::init src/Signal/Window.cc:69-82 */
    g.L = (int)(unsigned)rint(c.window_length_s * g.sampleRate);
    g.S = (int)(unsigned)rint(c.window_shift_s * g.sampleRate);
    /* FastFourierTransformNode::length src/Signal/FastFourierTransform.hh:298-308 and
     * FastFourierTransform::setLength src/Signal/FastFourierTransform.cc:30-41 */
    unsigned maximumLength = (unsigned)ceil(c.fft_max_input_s * g.sampleRate);
    if (maximumLength == 0) {
        g.N = 0;
    }
    else {
        double power = std::log((double)maximumLength) / std::log((double)2);
        if (almostEqual(power, rint(power)))
            power = rint(power);
        else
            power = ceil(power);
        g.N = 1 << (unsigned)power;
    }
    /* outputSampleRate() = length_/sampleRate_, written as an attribute
     * (src/Signal/FastFourierTransform.hh:93-95,293) */
    g.fftOutRate = attributeRoundTrip(g.N / g.sampleRate);
    return g;
}

/* ------------------------------------------------------------------ pre-emphasis
 * Signal::Preemphasis::apply src/Signal/Preemphasis.cc:51-74, init :31-35.  Packets are assumed
 * time-contiguous so only the first packet (re)initialises previous_ with its own first sample. */
struct Preemphasis {
    float alpha;
    float previous;
    bool  needInit;
    bool  fuse;
    Preemphasis(float a, bool f) : alpha(a), previous(0), needInit(true), fuse(f) {}
    void apply(std::vector<float>& v) {
        if (needInit) {
            previous = v.empty() ? 0.0f : v[0];
            needInit = false;
        }
        if (v.empty())
            return;
        if (alpha != 1.0) {
            for (size_t i = 0; i < v.size(); ++i) {
                float current = v[i];
                /* v[i] -= alpha_ * previous_  (fnmadd when contracted) */
                v[i]     = fuse ? std::fmaf(-alpha, previous, v[i]) : (v[i] - alpha * previous);
                previous = current;
            }
        }
        else {
            float carried = previous;
            previous      = v[v.size() - 1];
            for (size_t i = v.size() - 1; i > 0; --i)
                v[i] -= v[i - 1];
            v[0] -= carried;
        }
    }
};

/* ------------------------------------------------------------------ framing
 * Signal::WindowBuffer::{put,get,flush,copy} src/Signal/WindowBuffer.cc:50-126 */
struct Frame {
    std::vector<float> data;
    Time               start, end;
};

struct WindowBuffer {
    unsigned          length, shift;
    double            sampleRate;
    std::deque<float> buffer;
    Time              bufferStart;
    bool              flushed;
    WindowBuffer(unsigned L, unsigned S, double sr) : length(L), shift(S), sampleRate(sr), bufferStart(0), flushed(false) {}

    void put(const std::vector<float>& in, Time inStart) {
        if (flushed) { /* needInit_ after the last flush: init() -> reset() (WindowBuffer.cc:45-48,117-118) */
            flushed = false;
            buffer.clear();
        }
        if (buffer.empty())
            bufferStart = inStart;
        buffer.insert(buffer.end(), in.begin(), in.end());
    }
    Time endTime() const {
        return bufferStart + (Time)buffer.size() / (Time)sampleRate;
    }
    void copyOut(Frame& out, unsigned n) {
        out.data.assign(buffer.begin(), buffer.begin() + n);
        out.start = bufferStart;
        out.end   = bufferStart + (Time)out.data.size() / (Time)sampleRate;
    }
    void advance() {
        buffer.erase(buffer.begin(), buffer.begin() + std::min<size_t>(shift, buffer.size()));
        bufferStart += (Time)shift / (Time)sampleRate;
    }
    bool get(Frame& out) {
        if (buffer.size() < 2 * std::max(shift, length))
            return false;
        copyOut(out, length);
        advance();
        return true;
    }
    /* flush-all = false (src/Signal/Window.cc:112-113) */
    bool flush(Frame& out) {
        if (flushed || buffer.empty())
            return false;
        flushed = (std::max(shift, length) >= buffer.size());
        copyOut(out, std::min<unsigned>(length, (unsigned)buffer.size()));
        if (!flushed)
            advance();
        return true;
    }
};

/* ------------------------------------------------------------------ DC detection
 * Signal::DcDetection::{put,get,nextBlock,lastBlock,copyBlock,eraseBlock,flush} src/Signal/DcDetection.cc:89-226,
 * isNonDC / isDcDetected src/Signal/DcDetection.hh:73-82; node parameters :231-241 (samples.flow:34-35 sets
 * min-dc-length .0125, max-dc-increment 0.9, min-non-dc-segment-length .026).  Input packets are time-contiguous. */
struct DcBlock {
    std::vector<float> data;
    Time               start;
    long               firstSample; /* index of data[0] in the input stream (bookkeeping for the tests) */
};

struct DcDetection {
    double            sampleRate;
    float             maxDcIncrement;
    unsigned          minDcLength, minNonDcSegmentLength, maximalOutputSize;
    unsigned          nonDcLength, dcLength, nonDcSegmentLength;
    std::deque<float> buffer;
    Time              bufferStart;
    long              consumed; /* input samples erased from the buffer so far */
    DcDetection(double sr, double minDcS, float maxInc, double minNonDcS, unsigned maxOut)
            : sampleRate(sr), maxDcIncrement(maxInc), minDcLength((unsigned)rint(minDcS * sr)),
              minNonDcSegmentLength((unsigned)rint(minNonDcS * sr)), maximalOutputSize(maxOut), nonDcLength(1),
              dcLength(0), nonDcSegmentLength(0), bufferStart(0), consumed(0) {}
    bool isNonDC(float v) const {
        return fabs(v - buffer[nonDcLength - 1]) >= maxDcIncrement;
    }
    bool isDcDetected() const {
        return dcLength >= minDcLength;
    }
    void put(const std::vector<float>& in, Time inStart) {
        if (buffer.empty())
            bufferStart = inStart;
        buffer.insert(buffer.end(), in.begin(), in.end());
    }
    bool nextBlock() {
        while ((nonDcLength + dcLength) < buffer.size()) {
            if (isNonDC(buffer[nonDcLength + dcLength])) {
                if (isDcDetected())
                    return true;
                nonDcLength += dcLength; /* include the DC hypotheses */
                dcLength = 0;
                if (nonDcLength >= std::max(minNonDcSegmentLength, maximalOutputSize))
                    return true;
                nonDcLength++; /* include the new non-DC sample */
            }
            else
                dcLength++;
        }
        return false;
    }
    bool lastBlock() {
        if (buffer.empty())
            return false;
        if (isDcDetected())
            return true;
        nonDcLength += dcLength;
        dcLength = 0;
        return true;
    }
    bool flushBlock(DcBlock& out) {
        bool result = false;
        out.data.clear();
        if ((nonDcSegmentLength += nonDcLength) >= minNonDcSegmentLength) { /* copyBlock */
            out.data.assign(buffer.begin(), buffer.begin() + nonDcLength);
            out.start       = bufferStart;
            out.firstSample = consumed;
            result          = true;
        }
        if (dcLength > 0)
            nonDcSegmentLength = 0;
        buffer.erase(buffer.begin(), buffer.begin() + nonDcLength + dcLength); /* eraseBlock */
        consumed += nonDcLength + dcLength;
        bufferStart += (Time)(nonDcLength + dcLength) / sampleRate;
        nonDcLength = 1;
        dcLength    = 0;
        return result;
    }
    bool get(DcBlock& out) {
        do {
            if (!nextBlock())
                return false;
        } while (!flushBlock(out));
        return true;
    }
    bool flush(DcBlock& out) {
        if (!lastBlock())
            return false;
        return flushBlock(out);
    }
};

/* ------------------------------------------------------------------ window function
 * {Rectangular,Bartlett,Hamming,Hanning,Blackman}WindowFunction::init src/Signal/WindowFunction.cc:62-132; type
 * numbering: 0 hamming (the default), 1 rectangular, 2 hanning, 3 periodic-hanning, 4 bartlett, 5 blackman,
 * 6 kaiser (KaiserWindowFunction.cc:22-33 with Math::Nr::bessi0, src/Math/Nr/BesselFunctions.cc:22-37; beta is 0:
 * WindowFunction::create constructs it with the default and nothing calls setBeta) */
double nrBessi0(double x) {
    double ax, ans, y;
    if ((ax = fabs(x)) < 3.75) {
        y = x / 3.75;
        y *= y;
        ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    else {
        y   = 3.75 / ax;
        ans = (exp(ax) / sqrt(ax)) *
              (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
    }
    return ans;
}
std::vector<float> windowFunction(int type, unsigned length) {
    std::vector<float> w(length, 0.0f);
    if (type == 1) {
        std::fill(w.begin(), w.end(), 1.0f);
        return w;
    }
    if (length <= 1)
        return w;
    if (type == 2 || type == 3) {
        unsigned M = length - (type == 3 ? 0 : 1);
        for (unsigned n = 0; n <= M / 2; ++n) {
            w[n] = 0.5 - 0.5 * cos(2.0 * M_PI * n / M);
            if (M - n < length)
                w[M - n] = w[n];
        }
        return w;
    }
    unsigned M = length - 1;
    if (type == 6) {
        const double beta = 0.0;
        for (unsigned n = 0; n <= M / 2; ++n)
            w[n] = w[M - n] = nrBessi0(beta * sqrt(1.0 - ((double)n / (M / 2.0) - 1.0) * ((double)n / (M / 2.0) - 1.0))) / nrBessi0(beta);
        return w;
    }
    for (unsigned n = 0; n <= M / 2; ++n) {
        if (type == 0)
            w[n] = w[M - n] = 0.54 - 0.46 * cos(2.0 * M_PI * n / M);
        else if (type == 4)
            w[n] = w[M - n] = 2.0 * (float)n / (float)M;
        else
            w[n] = w[M - n] = 0.42 - 0.5 * cos(2.0 * M_PI * n / M) + 0.08 * cos(4.0 * M_PI * n / M);
    }
    return w;
}
std::vector<float> hammingWindow(unsigned length) {
    return windowFunction(0, length);
}

/* WindowFunction::work src/Signal/WindowFunction.hh:80-94 via Window::transform src/Signal/Window.cc:84-96 */
void applyWindow(std::vector<float>& frame, const std::vector<float>& w) {
    size_t n = std::min(frame.size(), w.size());
    for (size_t i = 0; i < n; ++i)
        frame[i] = w[i] * frame[i];
    for (size_t i = n; i < frame.size(); ++i)
        frame[i] = 0.0f;
}

/* ------------------------------------------------------------------ FFT
 * Math::FastFourierTransform::{createBitReversalReordering,transform,transformReal}
 * src/Math/FastFourierTransform.cc:28-146.  f32 data, f64 trigonometric recurrence, the products
 * with the f64 twiddle are formed in f64 and narrowed.  The two literals for pi are the reference's. */
const double kPi  = 3.141592653589793238;
const double kDPi = 6.28318530717959;

void bitReverseComplexPairs(float* v, unsigned size) {
    /* the reference precomputes a swap table (reording_) and applies std::swap(v[i], v[table[i]])
     * for every i; the table only has non-identity entries at (i-1,i) -> (j-1,j) for j > i, so
     * every pair is exchanged exactly once. */
    unsigned half = size / 2;
    unsigned j    = 1;
    for (unsigned i = 1; i < size; i += 2) {
        if (j > i) {
            std::swap(v[i - 1], v[j - 1]);
            std::swap(v[i], v[j]);
        }
        unsigned m = half;
        while (m >= 2 && j > m) {
            j -= m;
            m >>= 1;
        }
        j += m;
    }
}

void complexForward(float* v, unsigned size) {
    bitReverseComplexPairs(v, size);
    for (unsigned span = 2; span < size; span <<= 1) {
        unsigned step  = span << 1;
        double   theta = kDPi / span;
        double   sh    = std::sin(0.5 * theta);
        double   wpR   = -2.0 * sh * sh;
        double   wpI   = std::sin(theta);
        double   wR = 1.0, wI = 0.0;
        for (unsigned m = 1; m < span; m += 2) {
            for (unsigned i = m; i <= size; i += step) {
                unsigned j  = i + span;
                float    tR = wR * v[j - 1] - wI * v[j];
                float    tI = wR * v[j] + wI * v[j - 1];
                v[j - 1]    = v[i - 1] - tR;
                v[j]        = v[i] - tI;
                v[i - 1] += tR;
                v[i] += tI;
            }
            double old = wR;
            wR         = wR * wpR - wI * wpI + wR;
            wI         = wI * wpR + old * wpI + wI;
        }
    }
}

void realForwardPacked(float* v, unsigned size) {
    const double theta = kPi / (size >> 1);
    const float  c     = -0.5f;
    complexForward(v, size);
    double sh  = std::sin(0.5 * theta);
    double wpR = -2.0 * sh * sh;
    double wpI = std::sin(theta);
    double wR  = wpR + 1;
    double wI  = wpI;
    for (unsigned i = 1; i < (size >> 2); ++i) {
        unsigned a = i + i, b = a + 1, p = size - a, q = p + 1;
        double   h1R = 0.5 * (v[a] + v[p]);
        double   h1I = 0.5 * (v[b] - v[q]);
        double   h2R = -c * (v[b] + v[q]);
        double   h2I = c * (v[a] - v[p]);
        v[a]         = h1R + wR * h2R - wI * h2I;
        v[b]         = h1I + wR * h2I + wI * h2R;
        v[p]         = h1R - wR * h2R + wI * h2I;
        v[q]         = -h1I + wR * h2I + wI * h2R;
        double old   = wR;
        wR           = wR * wpR - wI * wpI + wR;
        wI           = wI * wpR + old * wpI + wI;
    }
    float h = v[0];
    v[0]    = h + v[1];
    v[1]    = h - v[1];
}

/* Signal::FastFourierTransform::transform src/Signal/FastFourierTransform.cc:75-83: right zero padding
 * :51-55, RealFastFourierTransform::applyAlgorithm/unpack :88-100, estimateContinuous :66-73. */
void realFftNode(std::vector<float>& data, unsigned N, double sampleRate) {
    data.resize(N, 0.0f);
    realForwardPacked(data.data(), N);
    data.push_back(data[1]);
    data.push_back(0.0f);
    data[1] = 0.0f;
    if (sampleRate != 1) {
        const float scale = 1 / (float)sampleRate;
        for (size_t i = 0; i < data.size(); ++i)
            data[i] = data[i] * scale;
    }
}

/* ------------------------------------------------------------------ amplitude
 * alternatingComplexVectorAmplitude src/Signal/ComplexVectorFunction.hh:29-45, Math::pointerAbs
 * src/Math/Complex.hh:39-46 -> std::abs(std::complex<f32>) */
void amplitudeSpectrum(const std::vector<float>& x, std::vector<float>& out) {
    out.resize(x.size() / 2);
    for (size_t k = 0; k < out.size(); ++k)
        out[k] = std::abs(std::complex<float>(x[2 * k], x[2 * k + 1]));
}

/* ------------------------------------------------------------------ mel filter bank
 * FilterBankNode::init src/Signal/Filterbank.cc:765-797, createAnalyticFunction :799-820,
 * StretchToCover :523-569, FilterBank::init :640-664, FilterBuilder::{create,setStart,setEnd,setWeights}
 * :144-217, SymmetricalTriangularFilterBuilder::weight :236-244; analytic functions
 * src/Math/AnalyticFunction.hh:116-131, SimpleAnalyticFunctions.hh:107-123,
 * AcousticalAnalyticFunctions.hh:24-55, AnalyticFunctionFactory.cc:338-341 (continuous domain). */
struct MelFilterBank {
    std::vector<int>                start, end;
    std::vector<std::vector<float>> weights;
};

struct MelFunctions {
    double a;  // discrete-to-continuous scaling 1 / sampleRate_
    double d2c(double k) const { return a * k; }
    static double warp(double f) { return 2595.0 * log10(1.0 + f / 700.0); }
    double        index2mel(double k) const { return warp(d2c(k)); }
    /* invert(nest(warp, d2c)) = nest(d2c^-1, warp^-1);  warp^-1 = nest(mel-core^-1, scaling(1/2595)) */
    double mel2index(double m) const {
        double f = (pow(10, (1 / 2595.0) * m) - 1.0) * 700.0;
        return (1 / a) * f;
    }
    /* nest(derive(warp), d2c) with derive(nest(scaling(2595), mel-core)) =
     * nest(constant(2595), mel-core) * derived-mel-core */
    double derivative(double k) const { return 2595.0 * (1.0 / log(10) / (700.0 + d2c(k))); }
};

MelFilterBank buildMelFilterBank(double fbSampleRate, unsigned inputSize, double filterWidthParam) {
    MelFunctions fn;
    fn.a = 1 / fbSampleRate;
    const double minimumFrequency = 0;
    const double maximumFrequency = fn.index2mel(inputSize - 1);

    /* Boundary::init / setSpacing :423-448, StretchToCover::init :547-569 */
    const double centerPos   = 0.5;
    double       filterWidth = filterWidthParam;
    double       spacing     = centerPos * filterWidth;  // spacing parameter 0 => left flank
    double       nf          = (maximumFrequency - minimumFrequency - filterWidth) / spacing + 1;
    if (nf < 1)
        nf = 1;
    else if (almostInteger(nf))
        nf = std::round(nf);
    size_t nFilters = (size_t)std::floor(nf);
    double coverage = (spacing * (nFilters - 1) + filterWidth) / (maximumFrequency - minimumFrequency);
    bool   keep     = (nFilters == 1 && coverage > 1 && !almostEqual(coverage, 1));
    if (!keep) {
        filterWidth /= coverage;
        spacing /= coverage;
    }

    MelFilterBank fb;
    fb.start.resize(nFilters);
    fb.end.resize(nFilters);
    fb.weights.resize(nFilters);
    for (size_t i = 0; i < nFilters; ++i) {
        double center = minimumFrequency + spacing * i + centerPos * filterWidth;
        double s      = fn.mel2index(std::max(center - centerPos * filterWidth, minimumFrequency));
        s             = almostInteger(s) ? std::round(s) : std::ceil(s);
        double e      = fn.mel2index(std::min(center + (1.0 - centerPos) * filterWidth, maximumFrequency));
        e             = almostInteger(e) ? std::round(e) + 1 : std::ceil(e);
        fb.start[i]   = (int)(size_t)s;
        fb.end[i]     = (int)(size_t)e;
        fb.weights[i].resize(fb.end[i] - fb.start[i]);
        for (unsigned f = fb.start[i]; f < (unsigned)fb.end[i]; ++f) {
            float tri = (double)1 - std::fabs(fn.index2mel(f) - center) / (filterWidth / 2);
            if (!(tri >= 0))
                tri = 0;
            fb.weights[i][f - fb.start[i]] = tri * fn.derivative(f);
        }
    }
    return fb;
}

/* FilterBank::Filter::apply src/Signal/Filterbank.cc:65-71 */
void applyFilterBank(const MelFilterBank& fb, const std::vector<float>& in, std::vector<float>& out, bool fuse) {
    out.resize(fb.start.size());
    for (size_t f = 0; f < fb.start.size(); ++f) {
        float r = 0;
        for (int k = fb.start[f]; k < fb.end[f]; ++k)
            r = mulAdd(in[k], fb.weights[f][k - fb.start[f]], r, fuse);
        out[f] = r;
    }
}

/* ------------------------------------------------------------------ cosine transform
 * CosineTransform::initEvenAboutNminusHalf src/Signal/CosineTransform.cc:62-74 with the identity
 * warping function (derivative constant 1), apply :76-83 -> Math::Matrix * Vector
 * src/Math/Matrix.hh:487-494, src/Math/Vector.hh:95-101 */
std::vector<float> buildDct(unsigned nOut, unsigned nIn) {
    std::vector<float> m((size_t)nOut * nIn);
    for (size_t k = 0; k < nOut; ++k)
        for (size_t n = 0; n < nIn; ++n) {
            double omega   = M_PI * (n + 0.5) / nIn;
            m[k * nIn + n] = cos(omega * k) * 1.0;
        }
    return m;
}

void applyDct(const std::vector<float>& m, unsigned nOut, const std::vector<float>& in, std::vector<float>& out,
              bool fuse) {
    out.resize(nOut);
    size_t nIn = in.size();
    for (size_t k = 0; k < nOut; ++k) {
        float r = 0;
        for (size_t n = 0; n < nIn; ++n)
            r = mulAdd(m[k * nIn + n], in[n], r, fuse);
        out[k] = r;
    }
}

/* ------------------------------------------------------------------ derivatives
 * Signal::Regression::regressFirstOrder / regressSecondOrder src/Signal/Regression.cc:25-63 over the
 * 5-frame window of signal-delay (max-size 5, right 2, margin-policy copy, present-not-empty:
 * src/Signal/Delay.cc:30-41,137-183, SlidingWindow.hh:66-73 getClosest).  in[0] is the oldest frame. */
void regressFirst(const std::vector<const std::vector<float>*>& in, std::vector<float>& out, bool fuse) {
    std::fill(out.begin(), out.end(), 0.0f);
    float tm = 0.0f;
    for (unsigned i = 0; i < in.size(); ++i) {
        const std::vector<float>& f  = *in[i];
        float                     dt = float(i) - float(in.size() - 1) / 2.0;
        for (unsigned c = 0; c < out.size(); ++c)
            out[c] = mulAdd(dt, f[c], out[c], fuse);
        tm = mulAdd(dt, dt, tm, fuse);
    }
    for (unsigned c = 0; c < out.size(); ++c)
        out[c] /= tm;
}

void regressSecond(const std::vector<const std::vector<float>*>& in, std::vector<float>& out, bool fuse) {
    std::fill(out.begin(), out.end(), 0.0f);
    float tm = 0.0f, ns = 0.0f;
    for (unsigned i = 0; i < in.size(); ++i) {
        float dt = float(i) - float(in.size() - 1) / 2.0;
        tm       = mulAdd(dt, dt, tm, fuse);
        ns       = mulAdd(dt * dt * dt, dt, ns, fuse);
    }
    /* ns = tm*tm - n*ns : the second product is the one gcc fuses into the subtraction */
    ns = fuse ? std::fmaf(-float(in.size()), ns, tm * tm) : (tm * tm - float(in.size()) * ns);
    for (unsigned i = 0; i < in.size(); ++i) {
        const std::vector<float>& f  = *in[i];
        float                     dt = float(i) - float(in.size() - 1) / 2.0;
        for (unsigned c = 0; c < out.size(); ++c) {
            out[c] = mulAdd(f[c], tm, out[c], fuse);
            float t = f[c] * dt * dt;
            out[c]  = fuse ? std::fmaf(-t, float(in.size()), out[c]) : (out[c] - t * float(in.size()));
        }
    }
    for (unsigned c = 0; c < out.size(); ++c)
        out[c] *= 2.0 / ns;
}

}  // namespace

/* ====================================================================== C interface */

extern "C" int orc_frontend_get_geometry(const orc_frontend_cfg* cfg, orc_frontend_geometry* g) {
    if (!cfg || !g)
        return -1;
    Geometry geo = makeGeometry(*cfg);
    if (geo.N <= 0 || geo.L <= 0 || geo.S <= 0 || geo.L > geo.N)
        return -2;
    MelFilterBank fb = buildMelFilterBank(geo.fftOutRate, geo.N / 2 + 1, cfg->filter_width);
    g->win_length    = geo.L;
    g->win_shift     = geo.S;
    g->fft_length    = geo.N;
    g->n_bins        = geo.N / 2 + 1;
    g->n_filters     = (int)fb.start.size();
    g->n_weights     = 0;
    for (size_t i = 0; i < fb.weights.size(); ++i)
        g->n_weights += (int)fb.weights[i].size();
    g->feat_dim = cfg->n_cepstra * (cfg->derivatives ? 3 : 1);
    return 0;
}

extern "C" long orc_frontend_nframes(const orc_frontend_cfg* cfg, long n) {
    Geometry g = makeGeometry(*cfg);
    /* closed form of the get()/flush() protocol: frames start every S samples; the last frame is the
     * first whose remaining sample count is <= max(S, L) */
    if (n <= 0)
        return 0;
    long M = std::max(g.S, g.L);
    if (n <= M)
        return 1;
    return (n - M + g.S - 1) / g.S + 1;
}

extern "C" int orc_frontend_tables(const orc_frontend_cfg* cfg, float* window, int* fb_start, int* fb_end,
                                   float* fb_weights, float* dct) {
    Geometry      g  = makeGeometry(*cfg);
    MelFilterBank fb = buildMelFilterBank(g.fftOutRate, g.N / 2 + 1, cfg->filter_width);
    int           nb = g.N / 2 + 1;
    if (window) {
        std::vector<float> w = windowFunction(cfg->window_type, g.L);
        std::copy(w.begin(), w.end(), window);
    }
    for (size_t f = 0; f < fb.start.size(); ++f) {
        if (fb_start)
            fb_start[f] = fb.start[f];
        if (fb_end)
            fb_end[f] = fb.end[f];
        if (fb_weights) {
            for (int k = 0; k < nb; ++k)
                fb_weights[f * nb + k] = 0.0f;
            for (int k = fb.start[f]; k < fb.end[f]; ++k)
                fb_weights[f * nb + k] = fb.weights[f][k - fb.start[f]];
        }
    }
    if (dct) {
        std::vector<float> m = buildDct(cfg->n_cepstra, (unsigned)fb.start.size());
        std::copy(m.begin(), m.end(), dct);
    }
    return 0;
}

extern "C" void orc_fft_real_packed(float* v, int n) {
    realForwardPacked(v, (unsigned)n);
}

namespace {
long mfccImpl(const orc_frontend_cfg* cfg, const orc_dc_cfg* dcCfg, const float* samples, long n_samples, long chunk,
              float* feats, double* t_start, double* t_end, float* spectrum, float* amplitude, float* fbank,
              float* cepstra, long capacity, long* run_begin, long* run_end, double* run_start, long run_capacity,
              long* n_runs) {
    if (!cfg || (!samples && n_samples > 0))
        return -1;
    const bool fuse = cfg->use_fma != 0;
    Geometry   g    = makeGeometry(*cfg);
    if (g.N <= 0 || g.L <= 0 || g.S <= 0 || g.L > g.N)
        return -2;
    const unsigned     nBins  = g.N / 2 + 1;
    MelFilterBank      fb     = buildMelFilterBank(g.fftOutRate, nBins, cfg->filter_width);
    const unsigned     K      = cfg->n_cepstra;
    std::vector<float> dct    = buildDct(K, (unsigned)fb.start.size());
    std::vector<float> window = windowFunction(cfg->window_type, g.L);

    /* --- the pull loop of SlidingAlgorithmNode::work (src/Signal/SlidingAlgorithmNode.hh:60-80):
     * try get(); if it fails pull one more pre-emphasised packet; at end of stream flush(). */
    Preemphasis        pre(cfg->preemphasis_alpha, fuse);
    WindowBuffer       wb(g.L, g.S, g.sampleRate);
    std::vector<Frame> frames;
    long               fed = 0;
    if (chunk <= 0)
        chunk = n_samples > 0 ? n_samples : 1;
    if (!dcCfg) {
        bool eos = false;
        while (true) {
            Frame f;
            if (wb.get(f)) {
                frames.push_back(f);
                continue;
            }
            if (!eos && fed < n_samples) {
                long               n = std::min(chunk, n_samples - fed);
                std::vector<float> packet(samples + fed, samples + fed + n);
                Time               start = (Time)fed / g.sampleRate;
                pre.apply(packet);
                wb.put(packet, start);
                fed += n;
                continue;
            }
            eos = true;
            if (wb.flush(f)) {
                frames.push_back(f);
                continue;
            }
            break;
        }
    }
    else {
        /* samples.flow: ... -> signal-dc-detection -> (mfcc.flow) pre-emphasis -> window.  The detector's blocks are
         * time-contiguous while they belong to one non-DC segment; after a discarded DC stretch the next block
         * starts later than the previous one ended: Preemphasis re-initialises (Preemphasis.cc:54-55) and the window
         * node flushes like at end of stream before it accepts the block (SlidingAlgorithmNode.hh:83-99,
         * WindowBuffer.cc:56-63, flush-before-gap = true Window.cc:116). */
        DcDetection dc(g.sampleRate, dcCfg->min_dc_length_s, dcCfg->max_dc_increment,
                       dcCfg->min_non_dc_segment_length_s, (unsigned)dcCfg->maximal_output_size);
        Time        prevEnd = 0;
        bool        first   = true;
        long        nRuns   = 0;
        auto drain = [&](bool flushAll) {
            Frame f;
            while (wb.get(f))
                frames.push_back(f);
            if (flushAll)
                while (wb.flush(f))
                    frames.push_back(f);
        };
        auto feed = [&](DcBlock& b) {
            const bool gap = first || fabs(b.start - prevEnd) > 1e-9;
            if (gap) {
                drain(true);
                pre.needInit = true;
                if (run_begin && nRuns < run_capacity) {
                    run_begin[nRuns] = b.firstSample;
                    run_start[nRuns] = b.start;
                }
                ++nRuns;
            }
            if (run_end && nRuns <= run_capacity)
                run_end[nRuns - 1] = b.firstSample + (long)b.data.size();
            prevEnd = b.start + (Time)b.data.size() / g.sampleRate;
            first   = false;
            pre.apply(b.data);
            wb.put(b.data, b.start);
            drain(false);
        };
        DcBlock b;
        while (fed < n_samples) {
            long               n = std::min(chunk, n_samples - fed);
            std::vector<float> packet(samples + fed, samples + fed + n);
            dc.put(packet, (Time)fed / g.sampleRate);
            fed += n;
            while (dc.get(b))
                feed(b);
        }
        if (dc.flush(b))
            feed(b);
        drain(true);
        if (n_runs)
            *n_runs = nRuns;
    }

    const long                      T = (long)frames.size();
    if (capacity >= 0 && T > capacity)
        return -3;
    std::vector<std::vector<float>> cep(T);
    std::vector<float>              amp, fbOut;
    for (long t = 0; t < T; ++t) {
        std::vector<float>& x = frames[t].data;
        applyWindow(x, window);
        realFftNode(x, g.N, g.sampleRate);
        if (spectrum)
            std::copy(x.begin(), x.end(), spectrum + (size_t)t * (g.N + 2));
        amplitudeSpectrum(x, amp);
        if (amplitude)
            std::copy(amp.begin(), amp.end(), amplitude + (size_t)t * nBins);
        applyFilterBank(fb, amp, fbOut, fuse);
        if (fbank)
            std::copy(fbOut.begin(), fbOut.end(), fbank + (size_t)t * fbOut.size());
        /* Flow::VectorLogFunction<f32> src/Flow/SimpleFunction.hh:39-48: log10, no floor */
        for (size_t i = 0; i < fbOut.size(); ++i)
            fbOut[i] = std::log10(fbOut[i]);
        applyDct(dct, K, fbOut, cep[t], fuse);
        if (cepstra)
            std::copy(cep[t].begin(), cep[t].end(), cepstra + (size_t)t * K);
    }

    const int dim = K * (cfg->derivatives ? 3 : 1);
    for (long t = 0; t < T; ++t) {
        float* o = feats ? feats + (size_t)t * dim : 0;
        if (o)
            std::copy(cep[t].begin(), cep[t].end(), o);
        Time s = frames[t].start, e = frames[t].end;
        if (cfg->derivatives) {
            std::vector<const std::vector<float>*> win(5);
            for (int i = 0; i < 5; ++i) {
                long u = std::min<long>(std::max<long>(t + i - 2, 0), T - 1);
                win[i] = &cep[u];
                /* merged packets carry [min start, max end] of their inputs
                 * (src/Flow/Collector.hh:141-165, src/Flow/Merger.hh:79-101) */
                s = std::min(s, frames[u].start);
                e = std::max(e, frames[u].end);
            }
            if (o) {
                std::vector<float> d(K), dd(K);
                regressFirst(win, d, fuse);
                regressSecond(win, dd, fuse);
                std::copy(d.begin(), d.end(), o + K);
                std::copy(dd.begin(), dd.end(), o + 2 * K);
            }
        }
        if (t_start)
            t_start[t] = s;
        if (t_end)
            t_end[t] = e;
    }
    return T;
}
}  // namespace

extern "C" long orc_mfcc(const orc_frontend_cfg* cfg, const float* samples, long n_samples, long chunk, float* feats,
                         double* t_start, double* t_end, float* spectrum, float* amplitude, float* fbank,
                         float* cepstra) {
    return mfccImpl(cfg, 0, samples, n_samples, chunk, feats, t_start, t_end, spectrum, amplitude, fbank, cepstra, -1, 0,
                    0, 0, 0, 0);
}

extern "C" long orc_mfcc_dc(const orc_frontend_cfg* cfg, const orc_dc_cfg* dc, const float* samples, long n_samples,
                            long chunk, float* feats, double* t_start, double* t_end, long capacity, long* run_begin,
                            long* run_end, double* run_start, long run_capacity, long* n_runs) {
    if (!dc)
        return -1;
    return mfccImpl(cfg, dc, samples, n_samples, chunk, feats, t_start, t_end, 0, 0, 0, 0, capacity, run_begin, run_end,
                    run_start, run_capacity, n_runs);
}

extern "C" const char* orc_version(void) {
    return "rasr-oracle 1 (restatement of rwth-i6/rasr @8fe741b1)";
}
