// include/dsp/types.h — POD sample types of the dsp:: mirror.
// Same names, layout (sizeof == 8, interleaved) and operators as the reference's src/dsp/types.h:7-84,
// so host code that fills stream buffers or inspects results compiles unchanged. Device kernels see
// these as float2.
#pragma once
#include <math.h>

#define FL_M_PI 3.1415926535f  // the reference's float "pi" (src/dsp/types.h:4)

namespace dsp {
    struct complex_t {
        float re;
        float im;

        complex_t operator*(const float b) const { return complex_t{re * b, im * b}; }
        complex_t operator/(const float b) const { return complex_t{re / b, im / b}; }
        complex_t operator*(const complex_t& b) const { return complex_t{(re * b.re) - (im * b.im), (im * b.re) + (re * b.im)}; }
        complex_t operator+(const complex_t& b) const { return complex_t{re + b.re, im + b.im}; }
        complex_t operator-(const complex_t& b) const { return complex_t{re - b.re, im - b.im}; }
        inline complex_t conj() const { return complex_t{re, -im}; }
        inline float phase() const { return atan2f(im, re); }
        // linear-approximation arctangent, the formula the FM demodulators use (types.h:36-54)
        inline float fastPhase() const {
            const float mag_im = fabsf(im);
            if (re == 0.0f && im == 0.0f) { return 0.0f; }
            float ang;
            if (re >= 0.0f) { ang = (FL_M_PI / 4.0f) - (FL_M_PI / 4.0f) * ((re - mag_im) / (re + mag_im)); }
            else { ang = (3.0f * (FL_M_PI / 4.0f)) - (FL_M_PI / 4.0f) * ((re + mag_im) / (mag_im - re)); }
            return (im < 0.0f) ? -ang : ang;
        }
        inline float amplitude() const { return sqrtf((re * re) + (im * im)); }
        // NOTE: the reference computes |re| twice here (types.h:58-64); FeedForwardAGC parity depends on it
        inline float fastAmplitude() const {
            const float a = fabsf(re);
            return a + 0.4f * a;
        }
    };

    struct stereo_t {
        float l;
        float r;
        stereo_t operator*(const float b) const { return stereo_t{l * b, r * b}; }
        stereo_t operator+(const stereo_t& b) const { return stereo_t{l + b.l, r + b.r}; }
        stereo_t operator-(const stereo_t& b) const { return stereo_t{l - b.l, r - b.r}; }
    };

    static_assert(sizeof(complex_t) == 8 && sizeof(stereo_t) == 8, "samples must be 8-byte interleaved pairs");
}
