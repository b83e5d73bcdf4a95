/*
 * ref_fft_shim.cc -- thin C entry point around the REFERENCE's own FFT translation unit.
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it is compiled together with
 * /root/reference/src/Math/FastFourierTransform.cc (from where it lies, see oracle/Makefile target
 * _ref) into oracle/_ref/libref_fft*.so, which is git-ignored.  It lets the tests check the
 * restatement in frontend_oracle.cc against the reference's real object code.
 */
#include <Math/FastFourierTransform.hh>

#include <cstdio>
#include <cstdlib>
#include <vector>

/* the one symbol the reference TU needs from Core (src/Core/Assertions.hh:180) */
namespace AssertionsPrivate {
void assertionFailed(const char* type, const char* expr, const char* function, const char* filename,
                     unsigned int line) {
    std::fprintf(stderr, "reference assertion failed: %s %s in %s (%s:%u)\n", type, expr, function, filename, line);
    std::abort();
}
}  // namespace AssertionsPrivate

extern "C" void ref_fft_transform_real(float* data, int n) {
    std::vector<float>         v(data, data + n);
    Math::FastFourierTransform fft;
    fft.transformReal(v, false);
    for (int i = 0; i < n; ++i)
        data[i] = v[i];
}

extern "C" void ref_fft_transform_complex(float* data, int n) {
    std::vector<float>         v(data, data + n);
    Math::FastFourierTransform fft;
    fft.transform(v, false);
    for (int i = 0; i < n; ++i)
        data[i] = v[i];
}
