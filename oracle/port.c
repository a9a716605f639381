/* oracle/port.c — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * Plain-C, single-threaded restatement of the reference's algorithm for the hot path
 * (SURVEY.md §8a), one function per reference block, each citing the reference file:line it
 * follows. Float expression ORDER is kept identical to the reference (compiled with
 * -ffp-contract=off, no -ffast-math, SSE2 float evaluation) so that this port is bit-identical
 * to oracle/_ref/libqdsp_ref.so (the unmodified reference headers + VOLK-generic shim);
 * tests/test_oracle.py pins that equality, and pins both against tests/golden/.
 *
 * VOLK (third-party, un-vendored, unpinned by the reference; 2.x API) is restated with the
 * semantics of its `_generic` kernels: sequential float accumulation for dot products and a
 * recursive float phasor renormalised every 512 samples + at end of call for the rotator.
 *
 * Streams are processed per caller-supplied block partition (`blocks[nblocks]`) wherever the
 * reference's result depends on how the stream was cut into run() calls.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FL_M_PI 3.1415926535f /* src/dsp/types.h:4 */

typedef struct { float re, im; } cf32; /* == dsp::complex_t, src/dsp/types.h:7-67 */

/* ------------------------------------------------------------------------------------------
 * Tap design — src/dsp/window.h
 * ---------------------------------------------------------------------------------------- */

/* BlackmanWindow::getTapCount, src/dsp/window.h:36-50 (BlackmanBandpassWindow :104-118 identical) */
int port_blackman_tap_count(float cutoff, float transWidth, float sampleRate) {
    (void)cutoff;
    int M = (int)(4.0f / (transWidth / sampleRate));
    if (M < 4) M = 4;
    if (M % 2 == 0) M++;
    return M;
}

/* BlackmanWindow::createTaps, src/dsp/window.h:52-70. sin/cos resolve to the float overloads in
 * the reference (libstdc++), the "window" factor has no i in it (a constant), and the centre is
 * tc/2 (non-integer for odd tc) — all reproduced verbatim. */
void port_blackman_taps(float cutoff, float transWidth, float sampleRate, float* taps, int tapCount, float factor) {
    (void)transWidth;
    float fc = cutoff / sampleRate;
    if (fc > 1.0f) fc = 1.0f;
    float tc = (float)tapCount;
    float sum = 0.0f;
    for (int i = 0; i < tapCount; i++) {
        float val = (sinf(2.0f * FL_M_PI * fc * ((float)i - (tc / 2))) / ((float)i - (tc / 2))) *
                    (0.42f - (0.5f * cosf(2.0f * FL_M_PI / tc)) + (0.8f * cosf(4.0f * FL_M_PI / tc)));
        taps[i] = val;
        sum += val;
    }
    for (int i = 0; i < tapCount; i++) {
        taps[i] *= factor;
        taps[i] /= sum;
    }
}

/* BlackmanBandpassWindow::createTaps, src/dsp/window.h:120-141 */
void port_blackman_bandpass_taps(float cutoff, float transWidth, float offset, float sampleRate, float* taps,
                                 int tapCount, float factor) {
    (void)transWidth;
    float fc = cutoff / sampleRate;
    if (fc > 1.0f) fc = 1.0f;
    float tc = (float)tapCount;
    float sum = 0.0f;
    for (int i = 0; i < tapCount; i++) {
        float val = (sinf(2.0f * FL_M_PI * fc * ((float)i - (tc / 2))) / ((float)i - (tc / 2))) *
                    (0.42f - (0.5f * cosf(2.0f * FL_M_PI / tc)) + (0.8f * cosf(4.0f * FL_M_PI / tc)));
        taps[i] = val;
        sum += val;
    }
    for (int i = 0; i < tapCount; i++) {
        taps[i] *= cosf(2.0f * (offset / sampleRate) * FL_M_PI * (float)i);
        taps[i] *= factor;
        taps[i] /= sum;
    }
}

/* RRCTaps::createTaps, src/dsp/window.h:184-229 (double arithmetic, float parameters) */
void port_rrc_taps(int tapCount, float sampleRate, float baudRate, float alpha, float* taps) {
    tapCount |= 1;
    double spb = sampleRate / baudRate; /* float division, then widened */
    double scale = 0;
    for (int i = 0; i < tapCount; i++) {
        double x1, x2, x3, num, den;
        double xindx = i - tapCount / 2;
        x1 = FL_M_PI * xindx / spb;
        x2 = 4 * alpha * xindx / spb;
        x3 = x2 * x2 - 1;
        if (fabs(x3) >= 0.000001) {
            if (i != tapCount / 2)
                num = cos((1 + alpha) * x1) + sin((1 - alpha) * x1) / (4 * alpha * xindx / spb);
            else
                num = cos((1 + alpha) * x1) + (1 - alpha) * FL_M_PI / (4 * alpha);
            den = x3 * FL_M_PI;
        } else {
            if (alpha == 1) {
                taps[i] = -1;
                scale += taps[i];
                continue;
            }
            x3 = (1 - alpha) * x1;
            x2 = (1 + alpha) * x1;
            num = (sin(x2) * (1 + alpha) * FL_M_PI - cos(x3) * ((1 - alpha) * FL_M_PI * spb) / (4 * alpha * xindx) +
                   sin(x3) * spb * spb / (4 * alpha * xindx * xindx));
            den = -32 * FL_M_PI * alpha * alpha * xindx / spb;
        }
        taps[i] = (float)(4 * alpha * num / den);
        scale += taps[i];
    }
    for (int i = 0; i < tapCount; i++) taps[i] = (float)(taps[i] / scale);
}

/* ------------------------------------------------------------------------------------------
 * VOLK-generic kernels (restated; see file header)
 * ---------------------------------------------------------------------------------------- */
static cf32 dot_cf(const cf32* a, const float* t, int n) { /* volk_32fc_32f_dot_prod_32fc_generic */
    float r = 0.0f, i = 0.0f;
    for (int k = 0; k < n; k++) {
        r += a[k].re * t[k];
        i += a[k].im * t[k];
    }
    cf32 o = {r, i};
    return o;
}
static float dot_f(const float* a, const float* t, int n) { /* volk_32f_x2_dot_prod_32f_generic */
    float r = 0.0f;
    for (int k = 0; k < n; k++) r += a[k] * t[k];
    return r;
}

/* volk_32fc_s32fc_x2_rotator_32fc_generic: out=in*phase; phase*=inc; renormalise every 512
 * samples counted from the start of the call and once more at the end if a partial run happened.
 * Call sites: src/dsp/processing.h:64, src/dsp/source.h:56, src/dsp/demodulator.h:479. */
void port_rotator(const cf32* in, cf32* out, float inc_re, float inc_im, float* phase_re, float* phase_im, int count) {
    float pr = *phase_re, pi = *phase_im;
    int done = 0;
    for (int seg = 0; seg < count / 512; seg++) {
        for (int j = 0; j < 512; j++, done++) {
            float xr = in[done].re, xi = in[done].im;
            out[done].re = xr * pr - xi * pi;
            out[done].im = xr * pi + xi * pr;
            float nr = pr * inc_re - pi * inc_im, ni = pr * inc_im + pi * inc_re;
            pr = nr;
            pi = ni;
        }
        float h = hypotf(pr, pi);
        pr /= h;
        pi /= h;
    }
    int rem = count % 512;
    for (int j = 0; j < rem; j++, done++) {
        float xr = in[done].re, xi = in[done].im;
        out[done].re = xr * pr - xi * pi;
        out[done].im = xr * pi + xi * pr;
        float nr = pr * inc_re - pi * inc_im, ni = pr * inc_im + pi * inc_re;
        pr = nr;
        pi = ni;
    }
    if (rem) {
        float h = hypotf(pr, pi);
        pr /= h;
        pi /= h;
    }
    *phase_re = pr;
    *phase_im = pi;
}

/* Drift-free rotator used for NCO attribution (matches volk.h QDSP_ORACLE_ROTATOR_F64):
 * phasor(n) = exp(j*(ang0 + n*theta)) in float64, theta = atan2(inc) of the float-rounded inc. */
void port_rotator_f64(const cf32* in, cf32* out, float inc_re, float inc_im, double* ang, long long count) {
    const double theta = atan2((double)inc_im, (double)inc_re);
    const double two_pi = 6.283185307179586476925286766559;
    for (long long n = 0; n < count; n++) {
        double a = *ang + theta * (double)n;
        float pr = (float)cos(a), pi = (float)sin(a);
        float xr = in[n].re, xi = in[n].im;
        out[n].re = xr * pr - xi * pi;
        out[n].im = xr * pi + xi * pr;
    }
    *ang = fmod(*ang + theta * (double)count, two_pi);
}

/* FrequencyXlator::init phase increment, src/dsp/processing.h:20-21 */
void port_xlator_phase_delta(float sampleRate, float freq, float* re, float* im) {
    *re = cosf((freq / sampleRate) * 2.0f * FL_M_PI);
    *im = sinf((freq / sampleRate) * 2.0f * FL_M_PI);
}

/* ------------------------------------------------------------------------------------------
 * FIR — src/dsp/filter.h:51-74.  y[i] = sum_j taps[j] * x[i-(T-1)+j]; history = last T samples
 * of the stream (zeros before its start; the reference leaves it uninitialised, filter.h:28).
 * Result is independent of the block partition, so the port works on the whole stream.
 * ---------------------------------------------------------------------------------------- */
void port_fir_cf32(const float* taps, int T, const cf32* x, long long n, cf32* y) {
    cf32* buf = (cf32*)calloc((size_t)(n + T), sizeof(cf32)); /* [T zeros | stream] */
    memcpy(buf + T, x, (size_t)n * sizeof(cf32));
    for (long long i = 0; i < n; i++) y[i] = dot_cf(&buf[i + 1], taps, T); /* filter.h:65 */
    free(buf);
}
void port_fir_f32(const float* taps, int T, const float* x, long long n, float* y) {
    float* buf = (float*)calloc((size_t)(n + T), sizeof(float));
    memcpy(buf + T, x, (size_t)n * sizeof(float));
    for (long long i = 0; i < n; i++) y[i] = dot_f(&buf[i + 1], taps, T); /* filter.h:60 */
    free(buf);
}

/* ------------------------------------------------------------------------------------------
 * PolyphaseResampler — src/dsp/resampling.h:95-166
 * ---------------------------------------------------------------------------------------- */
/* buildTapPhases, resampling.h:137-166: tapPhases[(I-1)-p][t] = taps[t*I+p], zero padded.
 * phases must hold I*TPP floats, row-major [phase][t]. Returns TPP. */
int port_build_tap_phases(const float* taps, int T, int I, float* phases) {
    int tpp = (T + I - 1) / I;
    int cur = 0;
    for (int t = 0; t < tpp; t++)
        for (int p = 0; p < I; p++) phases[((I - 1) - p) * tpp + t] = (cur < T) ? taps[cur++] : 0.0f;
    return tpp;
}
void port_rates_to_ratio(float inSR, float outSR, int* interp, int* decim) { /* resampling.h:28-30 */
    int a = (int)inSR, b = (int)outSR;
    while (b) { int t = a % b; a = b; b = t; }
    int g = a < 0 ? -a : a;
    *interp = (int)(outSR / (float)g);
    *decim = (int)(inSR / (float)g);
}
/* run(), resampling.h:99-132: the (i, phase, index) schedule restarts at i=0 on every block. */
long long port_resamp_cf32(const float* taps, int T, int I, int D, const cf32* x, const int* blocks, int nblocks,
                           cf32* y, int* out_counts) {
    int tpp = (T + I - 1) / I;
    float* ph = (float*)malloc((size_t)I * tpp * sizeof(float));
    port_build_tap_phases(taps, T, I, ph);
    int maxb = 0;
    for (int b = 0; b < nblocks; b++) if (blocks[b] > maxb) maxb = blocks[b];
    cf32* buf = (cf32*)calloc((size_t)(maxb + 2 * tpp), sizeof(cf32)); /* [tpp history | block] */
    long long in_off = 0, out_off = 0;
    for (int b = 0; b < nblocks; b++) {
        int count = blocks[b];
        int outCount = (int)(((long long)count * I) / D); /* calcOutSize :95-97 */
        memcpy(&buf[tpp], x + in_off, (size_t)count * sizeof(cf32));
        int outIndex = 0;
        for (int i = 0; outIndex < outCount; i += D) {
            int phase = i % I;
            y[out_off + outIndex] = dot_cf(&buf[i / I], &ph[phase * tpp], tpp); /* :123 */
            outIndex++;
        }
        memmove(buf, &buf[count], (size_t)tpp * sizeof(cf32)); /* :129 */
        if (out_counts) out_counts[b] = outCount;
        in_off += count;
        out_off += outCount;
    }
    free(buf);
    free(ph);
    return out_off;
}
long long port_resamp_f32(const float* taps, int T, int I, int D, const float* x, const int* blocks, int nblocks,
                          float* y, int* out_counts) {
    int tpp = (T + I - 1) / I;
    float* ph = (float*)malloc((size_t)I * tpp * sizeof(float));
    port_build_tap_phases(taps, T, I, ph);
    int maxb = 0;
    for (int b = 0; b < nblocks; b++) if (blocks[b] > maxb) maxb = blocks[b];
    float* buf = (float*)calloc((size_t)(maxb + 2 * tpp), sizeof(float));
    long long in_off = 0, out_off = 0;
    for (int b = 0; b < nblocks; b++) {
        int count = blocks[b];
        int outCount = (int)(((long long)count * I) / D);
        memcpy(&buf[tpp], x + in_off, (size_t)count * sizeof(float));
        int outIndex = 0;
        for (int i = 0; outIndex < outCount; i += D) {
            int phase = i % I;
            y[out_off + outIndex] = dot_f(&buf[i / I], &ph[phase * tpp], tpp); /* :116 */
            outIndex++;
        }
        memmove(buf, &buf[count], (size_t)tpp * sizeof(float));
        if (out_counts) out_counts[b] = outCount;
        in_off += count;
        out_off += outCount;
    }
    free(buf);
    free(ph);
    return out_off;
}
/* The integer schedule alone (bit-exact gate): for output k of a block, i=k*D, phase=i%I,
 * index=i/I. Writes (phase,index) pairs for one block of `count` inputs; returns outCount. */
int port_resamp_schedule(int I, int D, int count, int* phase, int* index) {
    int outCount = (int)(((long long)count * I) / D);
    int k = 0;
    for (int i = 0; k < outCount; i += D, k++) {
        if (phase) phase[k] = i % I;
        if (index) index[k] = i / I;
    }
    return outCount;
}

/* PowerDecimator::run, src/dsp/resampling.h:220-249, INCLUDING its quirk for power>1: the
 * power-1 extra passes all re-read the *input* buffer (after flush()), so the output is the
 * first count/2^(power-1) pair averages, not a 2^power decimation. */
long long port_power_decim(unsigned int power, const cf32* x, const int* blocks, int nblocks, cf32* y, int* out_counts) {
    long long in_off = 0, out_off = 0;
    for (int b = 0; b < nblocks; b++) {
        int count = blocks[b];
        const cf32* in = x + in_off;
        cf32* out = y + out_off;
        if (power == 0) {
            memcpy(out, in, (size_t)count * sizeof(cf32));
        } else if (power == 1) {
            for (int j = 0; j < count; j += 2) {
                out[j / 2].re = (in[j].re + in[j + 1].re) * 0.5f;
                out[j / 2].im = (in[j].im + in[j + 1].im) * 0.5f;
            }
            count /= 2;
        }
        if (power > 1) {
            for (unsigned int i = 1; i < power; i++) {
                for (int j = 0; j < count; j += 2) {
                    out[j / 2].re = (in[j].re + in[j + 1].re) * 0.5f;
                    out[j / 2].im = (in[j].im + in[j + 1].im) * 0.5f;
                }
                count /= 2;
            }
        }
        if (out_counts) out_counts[b] = count;
        in_off += blocks[b];
        out_off += count;
    }
    return out_off;
}

/* ------------------------------------------------------------------------------------------
 * FM demodulation — src/dsp/demodulator.h:11-30 (fast_arctan2), :81-99 (FloatFMDemod::run)
 * ---------------------------------------------------------------------------------------- */
float port_fast_arctan2(float y, float x) {
    const float c1 = FL_M_PI / 4.0f;         /* FAST_ATAN2_COEF1 */
    float abs_y = fabsf(y);
    float r, angle;
    if (x == 0.0f && y == 0.0f) return 0.0f;
    if (x >= 0.0f) {
        r = (x - abs_y) / (x + abs_y);
        angle = c1 - c1 * r;
    } else {
        r = (x + abs_y) / (abs_y - x);
        angle = 3.0f * c1 - c1 * r;          /* FAST_ATAN2_COEF2 expands to 3.0f * FL_M_PI / 4.0f */
    }
    if (y < 0.0f) return -angle;
    return angle;
}
float port_fm_phasor_speed(float sampleRate, float deviation) { /* demodulator.h:43 */
    return (2 * FL_M_PI) / (sampleRate / deviation);
}
void port_fm_demod(const cf32* x, long long n, float phasorSpeed, float* phase_state, float* out) {
    float phase = *phase_state;
    for (long long i = 0; i < n; i++) {
        float cur = port_fast_arctan2(x[i].im, x[i].re);
        float diff = cur - phase;
        if (diff > 3.1415926535f) diff -= 2 * 3.1415926535f;
        else if (diff <= -3.1415926535f) diff += 2 * 3.1415926535f;
        out[i] = diff / phasorSpeed;
        phase = cur;
    }
    *phase_state = phase;
}

/* ------------------------------------------------------------------------------------------
 * VFO — src/dsp/vfo.h:19-36: xlator(-offset) -> resampler with the auto window, re-rated to
 * inSR*interp (gain = interp). Returns tapCount; taps may be NULL to query the size.
 * ---------------------------------------------------------------------------------------- */
int port_vfo_design(float inSR, float outSR, float bandWidth, float* taps, int maxTaps, int* interp, int* decim) {
    float m = inSR < outSR ? inSR : outSR;
    float realCutoff = (bandWidth < m ? bandWidth : m) / 2.0f; /* vfo.h:26 */
    int I, D;
    port_rates_to_ratio(inSR, outSR, &I, &D);
    float winSR = inSR * (float)I;                              /* vfo.h:32 (int promoted to float) */
    int tc = port_blackman_tap_count(realCutoff, realCutoff, winSR);
    if (interp) *interp = I;
    if (decim) *decim = D;
    if (taps && tc <= maxTaps) port_blackman_taps(realCutoff, realCutoff, winSR, taps, tc, (float)I);
    return tc;
}

/* The reference composition VFO -> FloatFMDemod on one stream, block partition honoured by the
 * rotator (renormalisation grid restarts per call) and the resampler (schedule restarts). */
long long port_vfo_fm(float offset, float inSR, float outSR, float bandWidth, float deviation, int nco_f64,
                      const cf32* x, const int* blocks, int nblocks, float* audio, int* out_counts, cf32* iq_out) {
    int I, D;
    int T = port_vfo_design(inSR, outSR, bandWidth, NULL, 0, &I, &D);
    float* taps = (float*)malloc((size_t)T * sizeof(float));
    port_vfo_design(inSR, outSR, bandWidth, taps, T, &I, &D);
    float inc_re, inc_im;
    port_xlator_phase_delta(inSR, -offset, &inc_re, &inc_im); /* vfo.h:28 */
    long long n = 0;
    for (int b = 0; b < nblocks; b++) n += blocks[b];
    cf32* mixed = (cf32*)malloc((size_t)n * sizeof(cf32));
    float pr = 1.0f, pi = 0.0f;
    double ang = 0.0;
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        if (nco_f64) port_rotator_f64(x + off, mixed + off, inc_re, inc_im, &ang, blocks[b]);
        else port_rotator(x + off, mixed + off, inc_re, inc_im, &pr, &pi, blocks[b]);
        off += blocks[b];
    }
    long long cap = (long long)((double)n * I / D) + nblocks + 16;
    cf32* iq = iq_out ? iq_out : (cf32*)malloc((size_t)cap * sizeof(cf32));
    long long m = port_resamp_cf32(taps, T, I, D, mixed, blocks, nblocks, iq, out_counts);
    float st = 0.0f;
    if (audio) port_fm_demod(iq, m, port_fm_phasor_speed(outSR, deviation), &st, audio);
    if (!iq_out) free(iq);
    free(mixed);
    free(taps);
    return m;
}

/* Same composition over a WINDOW of a longer stream (bench / full-size parity checks): the drift-free rotator starts at
 * angle start_ang (= theta * absolute index of x[0], reduced by the caller), the resampler from zero history. Outputs whose
 * window reaches before x[0] (the first ceil(T/D) + 1) are start-up and must be discarded by the caller. */
long long port_vfo_fm_window(float offset, float inSR, float outSR, float bandWidth, float deviation, double start_ang,
                             const cf32* x, const int* blocks, int nblocks, float* audio, int* out_counts, cf32* iq_out) {
    int I, D;
    int T = port_vfo_design(inSR, outSR, bandWidth, NULL, 0, &I, &D);
    float* taps = (float*)malloc((size_t)T * sizeof(float));
    port_vfo_design(inSR, outSR, bandWidth, taps, T, &I, &D);
    float inc_re, inc_im;
    port_xlator_phase_delta(inSR, -offset, &inc_re, &inc_im);
    long long n = 0;
    for (int b = 0; b < nblocks; b++) n += blocks[b];
    cf32* mixed = (cf32*)malloc((size_t)(n ? n : 1) * sizeof(cf32));
    double ang = start_ang;
    port_rotator_f64(x, mixed, inc_re, inc_im, &ang, n);
    long long cap = (long long)((double)n * I / D) + nblocks + 16;
    cf32* iq = iq_out ? iq_out : (cf32*)malloc((size_t)cap * sizeof(cf32));
    long long m = port_resamp_cf32(taps, T, I, D, mixed, blocks, nblocks, iq, out_counts);
    float st = 0.0f;
    if (audio) port_fm_demod(iq, m, port_fm_phasor_speed(outSR, deviation), &st, audio);
    if (!iq_out) free(iq);
    free(mixed);
    free(taps);
    return m;
}
/* theta of the float-rounded phase increment (what the f64 rotator and the CUDA closed-form NCO both use) */
double port_xlator_theta(float sampleRate, float freq) {
    float re, im;
    port_xlator_phase_delta(sampleRate, freq, &re, &im);
    return atan2((double)im, (double)re);
}

/* The rotator's phase STATE at the start of every 512-sample run of every call (block), i.e. what
 * volk_32fc_s32fc_x2_rotator_32fc_generic holds right after each renormalisation (and at the start of each call). The
 * sequence does not depend on the samples. ckpt[2*k], ckpt[2*k+1] = (re, im) for run k; runs are numbered block by
 * block, ceil(count_b / 512) per block. Returns the number of runs; *phase_re/_im carry the state across calls. */
long long port_rotator_checkpoints(float inc_re, float inc_im, float* phase_re, float* phase_im, const int* blocks,
                                   int nblocks, float* ckpt) {
    float pr = *phase_re, pi = *phase_im;
    long long k = 0;
    for (int b = 0; b < nblocks; b++) {
        int count = blocks[b];
        for (int seg = 0; seg < count / 512; seg++) {
            if (ckpt) { ckpt[2 * k] = pr; ckpt[2 * k + 1] = pi; }
            k++;
            for (int j = 0; j < 512; j++) {
                float nr = pr * inc_re - pi * inc_im, ni = pr * inc_im + pi * inc_re;
                pr = nr;
                pi = ni;
            }
            float h = hypotf(pr, pi);
            pr /= h;
            pi /= h;
        }
        int rem = count % 512;
        if (rem) {
            if (ckpt) { ckpt[2 * k] = pr; ckpt[2 * k + 1] = pi; }
            k++;
            for (int j = 0; j < rem; j++) {
                float nr = pr * inc_re - pi * inc_im, ni = pr * inc_im + pi * inc_re;
                pr = nr;
                pi = ni;
            }
            float h = hypotf(pr, pi);
            pr /= h;
            pi /= h;
        }
    }
    *phase_re = pr;
    *phase_im = pi;
    return k;
}

void port_agc(float fallRate, float sampleRate, const float* x, const int* blocks, int nblocks, float* y, float* level_state);

/* StereoFMDemod::run, src/dsp/demodulator.h:255-277, with its sub-blocks composed as init() wires them
 * (:203-216): mpx = FloatFMDemod(x); pilot = AGC(20, fs)(FIR<float>(BlackmanBandpassWindow(1000, 1000, 19000, fs))(mpx));
 * doubled = pilot*pilot; amb = mpx*doubled; out = {mpx + amb, mpx - amb}. Every sub-block sees the same run()
 * partition. The pilot FIR's first-call history is uninitialised in the reference (filter.h:28): zeros here. */
long long port_stereo_fm(float sampleRate, float deviation, const cf32* x, const int* blocks, int nblocks, float* out_lr) {
    long long n = 0;
    for (int b = 0; b < nblocks; b++) n += blocks[b];
    float* mpx = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
    float* pil = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
    float* agc = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
    float ph = 0.0f;
    port_fm_demod(x, n, port_fm_phasor_speed(sampleRate, deviation), &ph, mpx);
    int T = port_blackman_tap_count(1000.0f, 1000.0f, sampleRate);
    float* taps = (float*)malloc(sizeof(float) * (size_t)T);
    port_blackman_bandpass_taps(1000.0f, 1000.0f, 19000.0f, sampleRate, taps, T, 1.0f);
    port_fir_f32(taps, T, mpx, n, pil);
    float level = 0.0f;
    port_agc(20.0f, sampleRate, pil, blocks, nblocks, agc, &level);
    for (long long i = 0; i < n; i++) {
        float doubled = agc[i] * agc[i];   /* volk_32f_x2_multiply_32f */
        float amb = mpx[i] * doubled;
        out_lr[2 * i] = mpx[i] + amb;
        out_lr[2 * i + 1] = mpx[i] - amb;
    }
    free(mpx); free(pil); free(agc); free(taps);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Recurrent blocks
 * ---------------------------------------------------------------------------------------- */
/* BFMDeemp::run, src/dsp/filter.h:129-158 (alpha from :101-103). x,y interleaved (l,r). */
void port_deemp(float sampleRate, float tau, const float* x, long long n, float* y, float* lastL, float* lastR) {
    float dt = 1.0f / sampleRate;
    float alpha = dt / (tau + dt);
    float l = *lastL, r = *lastR;
    if (isnan(l)) l = 0.0f;
    if (isnan(r)) r = 0.0f;
    for (long long i = 0; i < n; i++) {
        l = (alpha * x[2 * i]) + ((1 - alpha) * l);
        r = (alpha * x[2 * i + 1]) + ((1 - alpha) * r);
        y[2 * i] = l;
        y[2 * i + 1] = r;
    }
    *lastL = l;
    *lastR = r;
}

/* AGC::run, src/dsp/processing.h:119-134: per-BLOCK decay in dB (double pow), running max of the
 * RAW samples (no fabs), scale by 1/level. level starts at 0 (log10f(0) = -inf -> pow -> 0). */
void port_agc(float fallRate, float sampleRate, const float* x, const int* blocks, int nblocks, float* y, float* level_state) {
    float corrected = fallRate / sampleRate; /* :94 */
    float level = *level_state;
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        int count = blocks[b];
        level = (float)pow(10, (double)(((10.0f * log10f(level)) - (corrected * count)) / 10.0f));
        for (int i = 0; i < count; i++)
            if (x[off + i] > level) level = x[off + i];
        float s = 1.0f / level;
        for (int i = 0; i < count; i++) y[off + i] = x[off + i] * s; /* volk_32f_s32f_multiply_32f */
        off += count;
    }
    *level_state = level;
}

/* ComplexAGC::run, src/dsp/processing.h:271-286 */
void port_complex_agc(float setPoint, float maxGain, float rate, const cf32* x, long long n, cf32* y, float* gain_state) {
    float g = *gain_state;
    for (long long i = 0; i < n; i++) {
        cf32 v = {x[i].re * g, x[i].im * g};
        y[i] = v;
        g += (setPoint - sqrtf((v.re * v.re) + (v.im * v.im))) * rate;
        if (g > maxGain) g = maxGain;
    }
    *gain_state = g;
}

/* FeedForwardAGC<complex_t>::run, src/dsp/processing.h:175-223, with complex_t::fastAmplitude's
 * bug (im_abs = fabsf(re), src/dsp/types.h:58-64) kept: amplitude == |re| + 0.4f*|re|.
 * Stream-level form: output i (i < n-1023) = x[i] / max(1e-4, max_{j<1024} amp(x[i+j])). The block
 * partition only decides when outputs are emitted, not their values. Returns n-1023 (or 0). */
long long port_ff_agc_cf32(const cf32* x, long long n, cf32* y) {
    const int W = 1024;
    if (n < W) return 0;
    long long m = n - W + 1;
    for (long long i = 0; i < m; i++) {
        float level = (float)1e-4;
        for (int j = 0; j < W; j++) {
            float re_abs = fabsf(x[i + j].re);
            float im_abs = fabsf(x[i + j].re);
            float val = (re_abs > im_abs) ? (re_abs + 0.4f * im_abs) : (im_abs + 0.4f * re_abs);
            if (val > level) level = val;
        }
        y[i].re = x[i].re / level;
        y[i].im = x[i].im / level;
    }
    return m;
}

/* CostasLoop<ORDER>::init/run, src/dsp/pll.h:14-27,47-102. state = {vcoFrequency, vcoPhase,
 * lastVCO.re, lastVCO.im}; initial {0,0,1,0}. */
void port_costas_coeffs(float bw, float* alpha, float* beta) {
    float damp = sqrtf(2.0f) / 2.0f;
    float denominator = (float)(1.0 + 2.0 * damp * bw + bw * bw);
    *alpha = (4 * damp * bw) / denominator;
    *beta = (4 * bw * bw) / denominator;
}
#define STEP(n) (((n) > 0.0f) ? 1.0f : -1.0f) /* src/dsp/utils/macros.h:6 */
void port_costas(int order, float bw, const cf32* x, long long n, cf32* y, float* state) {
    float alpha, beta;
    port_costas_coeffs(bw, &alpha, &beta);
    float vcoFrequency = state[0], vcoPhase = state[1], vr = state[2], vi = state[3];
    for (long long i = 0; i < n; i++) {
        cf32 o;
        o.re = (vr * x[i].re) - (vi * x[i].im);
        o.im = (vi * x[i].re) + (vr * x[i].im);
        y[i] = o;
        float error = 0.0f;
        if (order == 2) {
            error = o.re * o.im;
        } else if (order == 4) {
            error = (STEP(o.re) * o.im) - (STEP(o.im) * o.re);
        } else {
            const float K = (sqrtf(2.0) - 1);
            if (fabsf(o.re) >= fabsf(o.im))
                error = ((o.re > 0.0f ? 1.0f : -1.0f) * o.im - (o.im > 0.0f ? 1.0f : -1.0f) * o.re * K);
            else
                error = ((o.re > 0.0f ? 1.0f : -1.0f) * o.im * K - (o.im > 0.0f ? 1.0f : -1.0f) * o.re);
        }
        if (error > 1.0f) error = 1.0f;
        else if (error < -1.0f) error = -1.0f;
        vcoFrequency += beta * error;
        if (vcoFrequency > 1.0f) vcoFrequency = 1.0f;
        else if (vcoFrequency < -1.0f) vcoFrequency = -1.0f;
        vcoPhase += vcoFrequency + (alpha * error);
        while (vcoPhase > (2.0f * FL_M_PI)) vcoPhase -= (2.0f * FL_M_PI);
        while (vcoPhase < (-2.0f * FL_M_PI)) vcoPhase += (2.0f * FL_M_PI);
        vr = cosf(-vcoPhase);
        vi = sinf(-vcoPhase);
    }
    state[0] = vcoFrequency;
    state[1] = vcoPhase;
    state[2] = vr;
    state[3] = vi;
}

/* ------------------------------------------------------------------------------------------
 * Element-wise / layout / per-block-statistic blocks ("next" rows of the scope table)
 * ---------------------------------------------------------------------------------------- */

/* Add / Substract / Multiply<T>::run, src/dsp/math.h:32-37, 79-84, 126-131 over one matched pair of
 * blocks. op 0/1/2; dtype 0 = float, 1 = complex_t (add/subtract of complex or stereo streams are
 * float add/subtract over 2n floats; Multiply<complex_t> is volk_32fc_x2_multiply_32fc) */
void port_math(int op, int dtype, const float* a, const float* b, float* out, long long n) {
    if (dtype == 1 && op == 2) {
        for (long long k = 0; k < n; k++) {
            const float ar = a[2 * k], ai = a[2 * k + 1], br = b[2 * k], bi = b[2 * k + 1];
            out[2 * k] = ar * br - ai * bi;
            out[2 * k + 1] = ar * bi + ai * br;
        }
        return;
    }
    const long long nf = dtype == 1 ? 2 * n : n;
    for (long long k = 0; k < nf; k++) out[k] = op == 0 ? a[k] + b[k] : (op == 1 ? a[k] - b[k] : a[k] * b[k]);
}

/* audio.h / convertion.h run() bodies; op numbering = QDSP_LAYOUT_* (include/qdsp_b200.h):
 * 0 MonoToStereo audio.h:30, 1 ChannelsToStereo audio.h:80, 2 StereoToMono audio.h:129-131,
 * 3 StereoToChannels audio.h:173, 4 ComplexToStereo convertion.h:32, 5 ComplexToReal convertion.h:71,
 * 6 ComplexToImag convertion.h:110, 7 RealToComplex convertion.h:157 (nullBuffer is zeros, :137-140) */
void port_layout(int op, const float* in0, const float* in1, float* out0, float* out1, long long n) {
    for (long long k = 0; k < n; k++) {
        switch (op) {
            case 0: out0[2 * k] = in0[k]; out0[2 * k + 1] = in0[k]; break;
            case 1: out0[2 * k] = in0[k]; out0[2 * k + 1] = in1[k]; break;
            case 2: out0[k] = (in0[2 * k] + in0[2 * k + 1]) * 0.5f; break;
            case 3: out0[k] = in0[2 * k]; out1[k] = in0[2 * k + 1]; break;
            case 4: out0[2 * k] = in0[2 * k]; out0[2 * k + 1] = in0[2 * k + 1]; break;
            case 5: out0[k] = in0[2 * k]; break;
            case 6: out0[k] = in0[2 * k + 1]; break;
            case 7: out0[2 * k] = in0[k]; out0[2 * k + 1] = 0.0f; break;
        }
    }
}

/* Volume<T>: level = powf(volume, 2) only once setVolume() ran (processing.h:371-374; init :355-359 leaves
 * level = 1.0f); run :388-411 over nf floats */
float port_volume_level(float volume) { return powf(volume, 2); }
void port_volume(float level, int muted, const float* in, float* out, long long nf) {
    if (muted) {
        memset(out, 0, sizeof(float) * (size_t)nf);
        return;
    }
    for (long long k = 0; k < nf; k++) out[k] = in[k] * level;
}

/* Threshold::run, processing.h:593-595 */
void port_threshold(const float* in, unsigned char* out, long long n) {
    for (long long k = 0; k < n; k++) out[k] = (in[k] > 0.0f);
}

/* DelayImag::run, processing.h:325-330; *lastIm carried across calls */
void port_delay_imag(const cf32* in, cf32* out, long long n, float* lastIm) {
    for (long long k = 0; k < n; k++) {
        const cf32 v = in[k];
        out[k].re = v.re;
        out[k].im = *lastIm;
        *lastIm = v.im;
    }
}

static float port_mean_mag(const cf32* x, int n, float* mag) {
    /* volk_32fc_magnitude_32f + volk_32f_accumulator_s32f (generic: sequential float sum), then / (float)count */
    float acc = 0.0f;
    for (int k = 0; k < n; k++) {
        const float m = sqrtf(x[k].re * x[k].re + x[k].im * x[k].im);
        if (mag) mag[k] = m;
        acc += m;
    }
    return acc / (float)n;
}

/* AMDemod::run, demodulator.h:355-374, per run() block */
void port_amdemod(const cf32* x, const int* blocks, int nblocks, float* out) {
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        const float avg = port_mean_mag(x + off, blocks[b], out + off);
        for (int k = 0; k < blocks[b]; k++) out[off + k] -= avg;
        off += blocks[b];
    }
}

/* Squelch::run, processing.h:460-479, per run() block */
void port_squelch(float level, const cf32* x, const int* blocks, int nblocks, cf32* out) {
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        const float sum = port_mean_mag(x + off, blocks[b], NULL);
        if (10.0f * log10f(sum) >= level) memcpy(out + off, x + off, sizeof(cf32) * (size_t)blocks[b]);
        else memset(out + off, 0, sizeof(cf32) * (size_t)blocks[b]);
        off += blocks[b];
    }
}

/* SSBDemod: phaseDelta per mode (demodulator.h:402-412, float cos/sin overloads), run :479-480 = VOLK rotator per
 * run() block (phase carried) + real part */
void port_ssb_phase_delta(float sampleRate, float bandWidth, int mode, float* re, float* im) {
    if (mode == 0) {
        *re = cosf((bandWidth / sampleRate) * FL_M_PI);
        *im = sinf((bandWidth / sampleRate) * FL_M_PI);
    } else if (mode == 1) {
        *re = cosf(-(bandWidth / sampleRate) * FL_M_PI);
        *im = sinf(-(bandWidth / sampleRate) * FL_M_PI);
    } else {
        *re = 1.0f;
        *im = 0.0f;
    }
}
void port_ssbdemod(float sampleRate, float bandWidth, int mode, const cf32* x, const int* blocks, int nblocks, float* out) {
    float ir, ii, pr = 1.0f, pi = 0.0f;
    port_ssb_phase_delta(sampleRate, bandWidth, mode, &ir, &ii);
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        cf32* tmp = (cf32*)malloc(sizeof(cf32) * (size_t)(blocks[b] > 0 ? blocks[b] : 1));
        port_rotator(x + off, tmp, ir, ii, &pr, &pi, blocks[b]);
        for (int k = 0; k < blocks[b]; k++) out[off + k] = tmp[k].re;
        free(tmp);
        off += blocks[b];
    }
}

/* ------------------------------------------------------------------------------------------
 * MMClockRecovery<T>::run, src/dsp/clock_recovery.h:127-215 (T = float / complex_t), per run() block.
 * `taps` is the caller's INTERP_TAPS[129][8] (src/dsp/interpolation_taps.h, a baked MMSE table that
 * this repository does not reproduce). State layout (floats): [0] mu, [1] dynOmega, [2] lastOutput,
 * [3..8] p_0T p_1T p_2T (re, im each), [9..14] c_0T c_1T c_2T, [15] nextOffset (as float-encoded int),
 * [16..29] delay[0..6] (re, im; float streams use the re slots). The reference leaves delay[]
 * uninitialised for the first block; the oracle (and the product) define it as zeros.
 * ---------------------------------------------------------------------------------------- */
#define PORT_STEP(n) (((n) > 0.0f) ? 1.0f : -1.0f)
long long port_mm(int dtype, float omega, float gainOmega, float muGain, float omegaRelLimit, const float* taps,
                  const float* x, const int* blocks, int nblocks, float* out, int* out_counts, float* state) {
    const int es = dtype == 1 ? 2 : 1;                        /* floats per element */
    const float omegaMin = omega - (omega * omegaRelLimit);   /* :83-84 */
    const float omegaMax = omega + (omega * omegaRelLimit);
    float mu = state[0], dynOmega = state[1], lastOutput = state[2];
    cf32 p0 = {state[3], state[4]}, p1 = {state[5], state[6]}, p2 = {state[7], state[8]};
    cf32 c0 = {state[9], state[10]}, c1 = {state[11], state[12]}, c2 = {state[13], state[14]};
    int nextOffset = (int)state[15];
    float delay[2 * 14];
    for (int k = 0; k < 14; k++) { delay[2 * k] = state[16 + 2 * k]; delay[2 * k + 1] = state[17 + 2 * k]; }
    long long off = 0, total = 0;
    for (int b = 0; b < nblocks; b++) {
        const int count = blocks[b];
        const float* in = x + off * es;
        int outCount = 0;
        const int maxOut = (int)(2.0f * omega * (float)count);                  /* :135 */
        for (int k = 0; k < 7 && k < count; k++)                                  /* :138 (the reference copies 7 blindly) */
            for (int e = 0; e < es; e++) delay[es * (7 + k) + e] = in[es * k + e];
        int i = nextOffset;
        for (; i < count && outCount < maxOut;) {
            const float* t8 = taps + 8 * (int)roundf(mu * 128.0f);
            const float* src = (i < 7) ? &delay[es * i] : &in[es * (i - 7)];
            float phaseError;
            if (dtype == 0) {
                float outVal = 0.0f;
                for (int k = 0; k < 8; k++) outVal += src[k] * t8[k];           /* volk_32f_x2_dot_prod_32f, generic */
                out[total + outCount] = outVal;
                outCount++;
                phaseError = (PORT_STEP(lastOutput) * outVal) - (lastOutput * PORT_STEP(outVal));   /* :156 */
                lastOutput = outVal;
            } else {
                p2 = p1; p1 = p0;                                                /* :161-165 */
                c2 = c1; c1 = c0;
                float re = 0.0f, im = 0.0f;
                for (int k = 0; k < 8; k++) { re += src[2 * k] * t8[k]; im += src[2 * k + 1] * t8[k]; }
                p0.re = re; p0.im = im;
                out[2 * (total + outCount)] = re;
                out[2 * (total + outCount) + 1] = im;
                outCount++;
                c0.re = PORT_STEP(p0.re); c0.im = PORT_STEP(p0.im);              /* :178 */
                /* :181  (((p0 - p2) * conj(c1)) - ((c0 - c2) * conj(p1))).re with complex_t's operators (types.h:17-27) */
                const cf32 a = {p0.re - p2.re, p0.im - p2.im}, bq = {c1.re, -c1.im};
                const cf32 d = {c0.re - c2.re, c0.im - c2.im}, e = {p1.re, -p1.im};
                const float ab = (a.re * bq.re) - (a.im * bq.im);
                const float de = (d.re * e.re) - (d.im * e.im);
                phaseError = ab - de;
            }
            if (phaseError > 1.0f) phaseError = 1.0f;                           /* :185-186 */
            if (phaseError < -1.0f) phaseError = -1.0f;
            dynOmega = dynOmega + (gainOmega * phaseError);                     /* :190-192 */
            if (dynOmega > omegaMax) dynOmega = omegaMax;
            else if (dynOmega < omegaMin) dynOmega = omegaMin;
            mu = mu + dynOmega + (muGain * phaseError);                         /* :197-198 */
            const float roundedStep = floorf(mu);
            i += (int)roundedStep;                                               /* :201-202 */
            if (i < 0) i = 0;
            mu -= roundedStep;                                                   /* :205 */
        }
        nextOffset = i - count;                                                  /* :208 */
        for (int k = 0; k < 7; k++)                                              /* :211 */
            for (int e = 0; e < es; e++) delay[es * k + e] = (count - 7 + k >= 0) ? in[es * (count - 7 + k) + e] : 0.0f;
        if (out_counts) out_counts[b] = outCount;
        total += outCount;
        off += count;
    }
    state[0] = mu; state[1] = dynOmega; state[2] = lastOutput;
    state[3] = p0.re; state[4] = p0.im; state[5] = p1.re; state[6] = p1.im; state[7] = p2.re; state[8] = p2.im;
    state[9] = c0.re; state[10] = c0.im; state[11] = c1.re; state[12] = c1.im; state[13] = c2.re; state[14] = c2.im;
    state[15] = (float)nextOffset;
    for (int k = 0; k < 14; k++) { state[16 + 2 * k] = delay[2 * k]; state[17 + 2 * k] = delay[2 * k + 1]; }
    return total;
}
