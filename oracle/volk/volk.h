// oracle/volk/volk.h — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Stand-in for the third-party VOLK library (https://github.com/gnuradio/volk, 2.x series:
// the by-value `phase_inc` rotator call at reference src/dsp/processing.h:64 predates
// `rotator2`). VOLK is neither vendored nor version-pinned by the reference
// (only `#include <volk/volk.h>` at src/dsp/stream.h:4 and `target_link_libraries(... volk)`
// at CMakeLists.txt:24), so this header restates the *published algorithm of VOLK's
// `_generic` kernels* for the 15 kernels + 3 allocator helpers the reference calls.
// Together with the unmodified reference headers (included by path from /root/reference/src)
// it forms the parity oracle. Parity is therefore pinned on the reference's own call sites,
// not on a VOLK binary.
//
// Compile-time variants (see oracle/Makefile):
//   (default)                 generic-C semantics: sequential float accumulation, recursive
//                             float rotator renormalised every 512 samples and at call end.
//   QDSP_ORACLE_ROTATOR_F64   rotator phasor evaluated in closed form in float64
//                             (phase_n = phase_0 * exp(j*n*arg(inc))); used to attribute the
//                             float-recurrence drift of the generic rotator (NCO).
//   QDSP_ORACLE_SIMD_ORDER    dot products accumulate in 8 interleaved lanes, the order real
//                             SIMD VOLK kernels use; lets GCC vectorise. TIMING baseline only.
#pragma once
#include <complex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <mutex>

typedef std::complex<float> lv_32fc_t;
#define lv_cmake(r, i) lv_32fc_t((r), (i))
#define lv_creal(x) ((x).real())
#define lv_cimag(x) ((x).imag())

// ---- allocation (reference: src/dsp/stream.h:25-31, filter.h:25,28,47) -------------------
static inline size_t volk_get_alignment(void) { return 64; }
static inline void* volk_malloc(size_t size, size_t alignment) {
    void* p = nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    if (posix_memalign(&p, alignment, size ? size : alignment) != 0) return nullptr;
    return p;
}
static inline void volk_free(void* p) { free(p); }

// ---- dot products (reference: src/dsp/filter.h:60,65; resampling.h:116,123) ---------------
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input,
                                               const float* taps, unsigned int num_points) {
    const float* a = reinterpret_cast<const float*>(input);
#ifdef QDSP_ORACLE_SIMD_ORDER
    float re[8] = {0, 0, 0, 0, 0, 0, 0, 0}, im[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int n8 = num_points & ~7u, k = 0;
    for (; k < n8; k += 8)
        for (int l = 0; l < 8; l++) {
            re[l] += a[2 * (k + l)] * taps[k + l];
            im[l] += a[2 * (k + l) + 1] * taps[k + l];
        }
    float r = ((re[0] + re[1]) + (re[2] + re[3])) + ((re[4] + re[5]) + (re[6] + re[7]));
    float i = ((im[0] + im[1]) + (im[2] + im[3])) + ((im[4] + im[5]) + (im[6] + im[7]));
    for (; k < num_points; k++) { r += a[2 * k] * taps[k]; i += a[2 * k + 1] * taps[k]; }
    *result = lv_32fc_t(r, i);
#else
    float r = 0.0f, i = 0.0f;
    for (unsigned int k = 0; k < num_points; k++) {
        r += a[2 * k] * taps[k];
        i += a[2 * k + 1] * taps[k];
    }
    *result = lv_32fc_t(r, i);
#endif
}

static inline void volk_32f_x2_dot_prod_32f(float* result, const float* input, const float* taps,
                                            unsigned int num_points) {
#ifdef QDSP_ORACLE_SIMD_ORDER
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int n8 = num_points & ~7u, k = 0;
    for (; k < n8; k += 8)
        for (int l = 0; l < 8; l++) acc[l] += input[k + l] * taps[k + l];
    float r = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    for (; k < num_points; k++) r += input[k] * taps[k];
    *result = r;
#else
    float r = 0.0f;
    for (unsigned int k = 0; k < num_points; k++) r += input[k] * taps[k];
    *result = r;
#endif
}

// ---- rotator (reference: src/dsp/processing.h:64, source.h:56, demodulator.h:479) ----------
#define QDSP_VOLK_ROTATOR_RELOAD 512
#ifndef QDSP_ORACLE_ROTATOR_F64
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t* out, const lv_32fc_t* in,
                                                   const lv_32fc_t phase_inc, lv_32fc_t* phase,
                                                   unsigned int num_points) {
    // Explicit (ac-bd, ad+bc) float arithmetic: what std::complex operator* evaluates to for
    // finite operands, spelled out so the rounding sequence is unambiguous.
    float pr = phase->real(), pi = phase->imag();
    const float cr = phase_inc.real(), ci = phase_inc.imag();
    unsigned int done = 0;
    for (unsigned int seg = 0; seg < num_points / QDSP_VOLK_ROTATOR_RELOAD; seg++) {
        for (int j = 0; j < QDSP_VOLK_ROTATOR_RELOAD; j++, done++) {
            const float xr = in[done].real(), xi = in[done].imag();
            out[done] = lv_32fc_t(xr * pr - xi * pi, xr * pi + xi * pr);
            const float nr = pr * cr - pi * ci, ni = pr * ci + pi * cr;
            pr = nr; pi = ni;
        }
        const float h = hypotf(pr, pi);
        pr /= h; pi /= h;
    }
    unsigned int rem = num_points % QDSP_VOLK_ROTATOR_RELOAD;
    for (unsigned int j = 0; j < rem; j++, done++) {
        const float xr = in[done].real(), xi = in[done].imag();
        out[done] = lv_32fc_t(xr * pr - xi * pi, xr * pi + xi * pr);
        const float nr = pr * cr - pi * ci, ni = pr * ci + pi * cr;
        pr = nr; pi = ni;
    }
    if (rem) {
        const float h = hypotf(pr, pi);
        pr /= h; pi /= h;
    }
    *phase = lv_32fc_t(pr, pi);
}
#else
// Drift-free variant: the phasor for sample n of the stream is phase0 * exp(j*n*theta),
// theta = atan2 of the (float-rounded) increment, everything in float64, rounded to float
// once per sample before the float complex multiply. The running sample count is kept per
// `phase` pointer (the float state alone cannot carry 1e-9 rad precision across calls).
struct qdsp_rot64_state { double ang; lv_32fc_t last; };
static inline std::map<lv_32fc_t*, qdsp_rot64_state>& qdsp_rot64_table() {
    static std::map<lv_32fc_t*, qdsp_rot64_state> t;
    return t;
}
static inline std::mutex& qdsp_rot64_mutex() { static std::mutex m; return m; }
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t* out, const lv_32fc_t* in,
                                                   const lv_32fc_t phase_inc, lv_32fc_t* phase,
                                                   unsigned int num_points) {
    const double theta = atan2((double)phase_inc.imag(), (double)phase_inc.real());
    double ang;
    {
        std::lock_guard<std::mutex> lk(qdsp_rot64_mutex());
        auto& tab = qdsp_rot64_table();
        auto it = tab.find(phase);
        if (it == tab.end() || it->second.last != *phase) {
            ang = atan2((double)phase->imag(), (double)phase->real());  // (re)seed from float state
        } else {
            ang = it->second.ang;
        }
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (unsigned int n = 0; n < num_points; n++) {
        const double a = ang + theta * (double)n;
        const float pr = (float)cos(a), pi = (float)sin(a);
        const float xr = in[n].real(), xi = in[n].imag();
        out[n] = lv_32fc_t(xr * pr - xi * pi, xr * pi + xi * pr);
    }
    ang = fmod(ang + theta * (double)num_points, two_pi);
    *phase = lv_32fc_t((float)cos(ang), (float)sin(ang));
    {
        std::lock_guard<std::mutex> lk(qdsp_rot64_mutex());
        qdsp_rot64_table()[phase] = qdsp_rot64_state{ang, *phase};
    }
}
#endif

// ---- element-wise helpers (reference call sites in SURVEY.md §2.3) -------------------------
static inline void volk_32f_s32f_multiply_32f(float* out, const float* in, const float scalar,
                                              unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = in[k] * scalar;
}
static inline void volk_32f_x2_multiply_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = a[k] * b[k];
}
static inline void volk_32f_x2_add_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = a[k] + b[k];
}
static inline void volk_32f_x2_subtract_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = a[k] - b[k];
}
static inline void volk_32fc_x2_add_32fc(lv_32fc_t* out, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = a[k] + b[k];
}
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* out, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) {
        const float ar = a[k].real(), ai = a[k].imag(), br = b[k].real(), bi = b[k].imag();
        out[k] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
    }
}
static inline void volk_32f_x2_interleave_32fc(lv_32fc_t* out, const float* i, const float* q, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = lv_32fc_t(i[k], q[k]);
}
static inline void volk_32fc_deinterleave_32f_x2(float* i, float* q, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) { i[k] = in[k].real(); q[k] = in[k].imag(); }
}
static inline void volk_32fc_deinterleave_real_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = in[k].real();
}
static inline void volk_32fc_deinterleave_imag_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++) out[k] = in[k].imag();
}
static inline void volk_32fc_magnitude_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int k = 0; k < n; k++)
        out[k] = sqrtf(in[k].real() * in[k].real() + in[k].imag() * in[k].imag());
}
static inline void volk_32f_accumulator_s32f(float* result, const float* in, unsigned int n) {
    float acc = 0.0f;
    for (unsigned int k = 0; k < n; k++) acc += in[k];
    *result = acc;
}
