// include/dsp/convertion.h (sic) — ComplexToStereo, ComplexToReal, ComplexToImag, RealToComplex (reference
// src/dsp/convertion.h:5-170).
#pragma once
#include <dsp/audio.h>

namespace dsp {
    class ComplexToStereo : public detail::layout_block<ComplexToStereo, complex_t, stereo_t, QDSP_LAYOUT_COMPLEX_TO_STEREO> {
    public:
        ComplexToStereo() {}
        ComplexToStereo(stream<complex_t>* in) { init(in); }
        static_assert(sizeof(complex_t) == sizeof(stereo_t), "Can't convert complex to stereo: different sizes");  // convertion.h:11
    };
    class ComplexToReal : public detail::layout_block<ComplexToReal, complex_t, float, QDSP_LAYOUT_COMPLEX_TO_REAL> {
    public:
        ComplexToReal() {}
        ComplexToReal(stream<complex_t>* in) { init(in); }
    };
    class ComplexToImag : public detail::layout_block<ComplexToImag, complex_t, float, QDSP_LAYOUT_COMPLEX_TO_IMAG> {
    public:
        ComplexToImag() {}
        ComplexToImag(stream<complex_t>* in) { init(in); }
    };
    class RealToComplex : public detail::layout_block<RealToComplex, float, complex_t, QDSP_LAYOUT_REAL_TO_COMPLEX> {
    public:
        RealToComplex() {}
        RealToComplex(stream<float>* in) { init(in); }
    };
}
