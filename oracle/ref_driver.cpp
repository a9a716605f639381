// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Runs the UNMODIFIED reference blocks (headers included by path from /root/reference/src,
// never copied) through their own thread-per-block / stream<T> plumbing and exposes the
// results through a plain C ABI so pytest (ctypes) and bench.py's cpu_baseline leg can call
// them. Built only where /root/reference exists (this container) by oracle/Makefile into
// oracle/_ref/; the built .so travels to the GPU box, the sources of the reference do not.
//
// Every entry point feeds `in` in the caller-given block partition (`blocks[nblocks]`,
// each <= STREAM_BUFFER_SIZE) because several reference blocks have block-size-dependent
// semantics (resampler schedule restart src/dsp/resampling.h:121, AGC src/dsp/processing.h:123).
#include <dsp/block.h>
#include <dsp/stream.h>
#include <dsp/types.h>
#include <dsp/window.h>
#include <dsp/filter.h>
#include <dsp/resampling.h>
#include <dsp/processing.h>
#include <dsp/demodulator.h>
#include <dsp/pll.h>
#include <dsp/vfo.h>
#include <dsp/routing.h>
#include <dsp/math.h>
#include <dsp/audio.h>
#include <dsp/convertion.h>
#include <dsp/clock_recovery.h>
#include <new>

#include <chrono>
#include <thread>
#include <vector>
#include <cstring>

using namespace dsp;

namespace {

template <class T>
void feed(stream<T>* s, const T* data, const int* blocks, int nblocks) {
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        memcpy(s->writeBuf, data + off, (size_t)blocks[b] * sizeof(T));
        if (!s->swap(blocks[b])) return;
        off += blocks[b];
    }
}

// Drain exactly `nblocks` swaps from `s` into `out`; returns total elements, per-swap counts in
// out_counts (may be null).
template <class T>
long long drain(stream<T>* s, T* out, int nblocks, int* out_counts) {
    long long total = 0;
    for (int b = 0; b < nblocks; b++) {
        int n = s->read();
        if (n < 0) break;
        if (out) memcpy(out + total, s->readBuf, (size_t)n * sizeof(T));
        s->flush();
        if (out_counts) out_counts[b] = n;
        total += n;
    }
    return total;
}

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double seconds() const {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};

template <class Tin, class Tout>
long long pump(stream<Tin>* in, stream<Tout>* out, const Tin* data, const int* blocks, int nblocks,
               Tout* result, int* out_counts, double* seconds, int out_swaps = -1) {
    Timer t;
    std::thread feeder([&] { feed(in, data, blocks, nblocks); });
    long long total = drain(out, result, out_swaps < 0 ? nblocks : out_swaps, out_counts);
    feeder.join();
    if (seconds) *seconds = t.seconds();
    return total;
}

template <int ORDER>
long long costas_impl(float bw, const float* in, const int* blocks, int nblocks, float* out, double* seconds) {
    stream<complex_t> src;
    CostasLoop<ORDER> c(&src, bw);
    c.start();
    long long n = pump(&src, &c.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, nullptr, seconds);
    c.stop();
    return n;
}

}  // namespace

extern "C" {

int ref_stream_buffer_size() { return STREAM_BUFFER_SIZE; }

// ---- tap design: src/dsp/window.h ----------------------------------------------------------
int ref_blackman_tap_count(float cutoff, float transWidth, float sampleRate) {
    filter_window::BlackmanWindow w(cutoff, transWidth, sampleRate);
    return w.getTapCount();
}
void ref_blackman_taps(float cutoff, float transWidth, float sampleRate, float* taps, int tapCount, float factor) {
    filter_window::BlackmanWindow w(cutoff, transWidth, sampleRate);
    w.createTaps(taps, tapCount, factor);
}
int ref_blackman_bandpass_tap_count(float cutoff, float transWidth, float offset, float sampleRate) {
    filter_window::BlackmanBandpassWindow w(cutoff, transWidth, offset, sampleRate);
    return w.getTapCount();
}
void ref_blackman_bandpass_taps(float cutoff, float transWidth, float offset, float sampleRate, float* taps,
                                int tapCount, float factor) {
    filter_window::BlackmanBandpassWindow w(cutoff, transWidth, offset, sampleRate);
    w.createTaps(taps, tapCount, factor);
}
void ref_rrc_taps(int tapCount, float sampleRate, float baudRate, float alpha, float* taps) {
    RRCTaps w(tapCount, sampleRate, baudRate, alpha);
    w.createTaps(taps, tapCount);
}

// ---- FIR: src/dsp/filter.h:9-90 --------------------------------------------------------------
// NOTE the reference leaves the first-block history uninitialised (filter.h:28); glibc returns
// fresh zero pages for an allocation this large, and callers exclude the first tapCount-1
// outputs from parity anyway.
long long ref_fir_cf32(float cutoff, float transWidth, float sampleRate, const float* in, const int* blocks,
                       int nblocks, float* out, double* seconds) {
    filter_window::BlackmanWindow win(cutoff, transWidth, sampleRate);
    stream<complex_t> src;
    FIR<complex_t> fir(&src, &win);
    fir.start();
    long long n = pump(&src, &fir.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, nullptr, seconds);
    fir.stop();
    return n;
}
long long ref_fir_f32(float cutoff, float transWidth, float sampleRate, const float* in, const int* blocks,
                      int nblocks, float* out, double* seconds) {
    filter_window::BlackmanWindow win(cutoff, transWidth, sampleRate);
    stream<float> src;
    FIR<float> fir(&src, &win);
    fir.start();
    long long n = pump(&src, &fir.out, in, blocks, nblocks, out, nullptr, seconds);
    fir.stop();
    return n;
}

// ---- PolyphaseResampler: src/dsp/resampling.h:9-190 -----------------------------------------
// Window is BlackmanWindow(cutoff, transWidth, winSampleRate). `vfo_style` != 0 repeats what
// VFO::init does (src/dsp/vfo.h:29-33): re-rate the window to inSR*interp and updateWindow.
long long ref_resamp_cf32(float cutoff, float transWidth, float winSampleRate, float inSR, float outSR,
                          int vfo_style, const float* in, const int* blocks, int nblocks, float* out,
                          int* out_counts, int* interp, int* decim, double* seconds) {
    filter_window::BlackmanWindow win(cutoff, transWidth, winSampleRate);
    stream<complex_t> src;
    PolyphaseResampler<complex_t> rs(&src, &win, inSR, outSR);
    if (vfo_style) {
        win.setSampleRate(inSR * rs.getInterpolation());
        rs.updateWindow(&win);
    }
    if (interp) *interp = rs.getInterpolation();
    if (decim) *decim = rs.getDecimation();
    rs.start();
    long long n = pump(&src, &rs.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, out_counts, seconds);
    rs.stop();
    return n;
}
long long ref_resamp_f32(float cutoff, float transWidth, float winSampleRate, float inSR, float outSR,
                         const float* in, const int* blocks, int nblocks, float* out, int* out_counts,
                         int* interp, int* decim, double* seconds) {
    filter_window::BlackmanWindow win(cutoff, transWidth, winSampleRate);
    stream<float> src;
    PolyphaseResampler<float> rs(&src, &win, inSR, outSR);
    if (interp) *interp = rs.getInterpolation();
    if (decim) *decim = rs.getDecimation();
    rs.start();
    long long n = pump(&src, &rs.out, in, blocks, nblocks, out, out_counts, seconds);
    rs.stop();
    return n;
}

// ---- PowerDecimator: src/dsp/resampling.h:192-258 --------------------------------------------
long long ref_power_decim(unsigned int power, const float* in, const int* blocks, int nblocks, float* out,
                          int* out_counts) {
    stream<complex_t> src;
    PowerDecimator pd(&src, power);
    pd.start();
    long long n = pump(&src, &pd.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, out_counts, nullptr);
    pd.stop();
    return n;
}

// ---- FrequencyXlator: src/dsp/processing.h:9-82 ---------------------------------------------
long long ref_xlator(float sampleRate, float freq, const float* in, const int* blocks, int nblocks, float* out,
                     double* seconds) {
    stream<complex_t> src;
    FrequencyXlator<complex_t> x(&src, sampleRate, freq);
    x.start();
    long long n = pump(&src, &x.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, nullptr, seconds);
    x.stop();
    return n;
}
// phase increment exactly as the reference derives it (processing.h:21), for the GPU side to reuse
void ref_xlator_phase_delta(float sampleRate, float freq, float* re, float* im) {
    lv_32fc_t d = lv_cmake(std::cos((freq / sampleRate) * 2.0f * FL_M_PI), std::sin((freq / sampleRate) * 2.0f * FL_M_PI));
    *re = d.real();
    *im = d.imag();
}
// raw rotator call (shim semantics) with explicit phase in/out, one call per block
void ref_rotator(const float* in, float* out, float inc_re, float inc_im, float* phase_re, float* phase_im,
                 const int* blocks, int nblocks) {
    lv_32fc_t ph = lv_cmake(*phase_re, *phase_im);
    long long off = 0;
    for (int b = 0; b < nblocks; b++) {
        volk_32fc_s32fc_x2_rotator_32fc((lv_32fc_t*)out + off, (const lv_32fc_t*)in + off, lv_cmake(inc_re, inc_im), &ph,
                                        blocks[b]);
        off += blocks[b];
    }
    *phase_re = ph.real();
    *phase_im = ph.imag();
}

// ---- VFO: src/dsp/vfo.h ----------------------------------------------------------------------
long long ref_vfo(float offset, float inSR, float outSR, float bandWidth, const float* in, const int* blocks,
                  int nblocks, float* out, int* out_counts, double* seconds) {
    stream<complex_t> src;
    VFO vfo(&src, offset, inSR, outSR, bandWidth);
    vfo.start();
    long long n = pump(&src, vfo.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, out_counts, seconds);
    return n;  // VFO::stop() is a no-op in the reference (vfo.h:44-48); member dtors stop the workers
}
// taps/I/D a VFO ends up with (vfo.h:26-33), re-derived through the same public calls
int ref_vfo_design(float inSR, float outSR, float bandWidth, float* taps, int maxTaps, int* interp, int* decim) {
    float realCutoff = std::min<float>(bandWidth, std::min<float>(inSR, outSR)) / 2.0f;
    filter_window::BlackmanWindow win;
    win.init(realCutoff, realCutoff, inSR);
    stream<complex_t> dummy;
    PolyphaseResampler<complex_t> rs(&dummy, &win, inSR, outSR);
    win.setSampleRate(inSR * rs.getInterpolation());
    int tc = win.getTapCount();
    if (interp) *interp = rs.getInterpolation();
    if (decim) *decim = rs.getDecimation();
    if (taps && tc <= maxTaps) win.createTaps(taps, tc, rs.getInterpolation());
    return tc;
}

// ---- FM demod: src/dsp/demodulator.h:14-30,32-182 --------------------------------------------
long long ref_fm_demod(float sampleRate, float deviation, const float* in, const int* blocks, int nblocks,
                       float* out, double* seconds) {
    stream<complex_t> src;
    FloatFMDemod d(&src, sampleRate, deviation);
    d.start();
    long long n = pump(&src, &d.out, (const complex_t*)in, blocks, nblocks, out, nullptr, seconds);
    d.stop();
    return n;
}
long long ref_fm_demod_stereo(float sampleRate, float deviation, const float* in, const int* blocks, int nblocks,
                              float* out) {
    stream<complex_t> src;
    FMDemod d(&src, sampleRate, deviation);
    d.start();
    long long n = pump(&src, &d.out, (const complex_t*)in, blocks, nblocks, (stereo_t*)out, nullptr, nullptr);
    d.stop();
    return n;
}
float ref_fast_arctan2(float y, float x) { return fast_arctan2(y, x); }

// ---- StereoFMDemod: src/dsp/demodulator.h:189-330 (FloatFMDemod -> Splitter -> pilot FIR<float> with
// BlackmanBandpassWindow(1000, 1000, 19000, fs) -> AGC(20, fs); L/R = mpx +- mpx * pilot^2) ---------------
long long ref_stereo_fm(float sampleRate, float deviation, const float* in, const int* blocks, int nblocks,
                        float* out) {
    stream<complex_t> src;
    StereoFMDemod d(&src, sampleRate, deviation);
    d.start();
    long long n = pump(&src, &d.out, (const complex_t*)in, blocks, nblocks, (stereo_t*)out, nullptr, nullptr);
    d.stop();
    return n;
}

// ---- the fused chain as the reference composes it: VFO -> FloatFMDemod (3 worker threads) ----
long long ref_vfo_fm(float offset, float inSR, float outSR, float bandWidth, float deviation, const float* in,
                     const int* blocks, int nblocks, float* audio, int* out_counts, double* seconds) {
    stream<complex_t> src;
    VFO vfo(&src, offset, inSR, outSR, bandWidth);
    FloatFMDemod dem(vfo.out, outSR, deviation);
    vfo.start();
    dem.start();
    long long n = pump(&src, &dem.out, (const complex_t*)in, blocks, nblocks, audio, out_counts, seconds);
    dem.stop();
    return n;
}

// ---- channelizer: one Splitter fanning the same wideband stream to nch VFO+FloatFMDemod -------
// (src/dsp/routing.h:47-57). audio is [nch][outPerChannel] row-major; returns outputs/channel.
long long ref_channelizer_fm(int nch, const float* offsets, float inSR, float outSR, float bandWidth,
                             float deviation, const float* in, const int* blocks, int nblocks, float* audio,
                             long long audio_stride, double* seconds) {
    stream<complex_t> src;
    Splitter<complex_t> split(&src);
    std::vector<stream<complex_t>*> legs(nch);
    std::vector<VFO*> vfos(nch);
    std::vector<FloatFMDemod*> dems(nch);
    for (int c = 0; c < nch; c++) {
        legs[c] = new stream<complex_t>();
        split.bindStream(legs[c]);
        vfos[c] = new VFO(legs[c], offsets[c], inSR, outSR, bandWidth);
        dems[c] = new FloatFMDemod(vfos[c]->out, outSR, deviation);
    }
    for (int c = 0; c < nch; c++) { vfos[c]->start(); dems[c]->start(); }
    split.start();
    Timer t;
    std::thread feeder([&] { feed(&src, (const complex_t*)in, blocks, nblocks); });
    std::vector<std::thread> drains;
    std::vector<long long> totals(nch, 0);
    for (int c = 0; c < nch; c++)
        drains.emplace_back([&, c] { totals[c] = drain(&dems[c]->out, audio + (long long)c * audio_stride, nblocks, nullptr); });
    feeder.join();
    for (auto& d : drains) d.join();
    if (seconds) *seconds = t.seconds();
    split.stop();
    for (int c = 0; c < nch; c++) { dems[c]->stop(); }
    long long per = totals.empty() ? 0 : totals[0];
    for (int c = 0; c < nch; c++) { delete dems[c]; delete vfos[c]; delete legs[c]; }
    return per;
}

// ---- recurrent blocks ---------------------------------------------------------------------------
// BFMDeemp: src/dsp/filter.h:92-172
long long ref_deemp(float sampleRate, float tau, const float* in, const int* blocks, int nblocks, float* out,
                    double* seconds) {
    stream<stereo_t> src;
    BFMDeemp d(&src, sampleRate, tau);
    d.start();
    long long n = pump(&src, &d.out, (const stereo_t*)in, blocks, nblocks, (stereo_t*)out, nullptr, seconds);
    d.stop();
    return n;
}
// AGC: src/dsp/processing.h:84-147
long long ref_agc(float fallRate, float sampleRate, const float* in, const int* blocks, int nblocks, float* out,
                  double* seconds) {
    stream<float> src;
    AGC a(&src, fallRate, sampleRate);
    a.start();
    long long n = pump(&src, &a.out, in, blocks, nblocks, out, nullptr, seconds);
    a.stop();
    return n;
}
// ComplexAGC: src/dsp/processing.h:236-297
long long ref_complex_agc(float setPoint, float maxGain, float rate, const float* in, const int* blocks,
                          int nblocks, float* out, double* seconds) {
    stream<complex_t> src;
    ComplexAGC a(&src, setPoint, maxGain, rate);
    a.start();
    long long n = pump(&src, &a.out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, nullptr, seconds);
    a.stop();
    return n;
}
// FeedForwardAGC: src/dsp/processing.h:149-234. A run() that has not yet collected 1024 samples
// returns without swapping (:188-191), so the number of swaps can be < nblocks; out_counts[b] is
// the swap size (== input block size, :221) and valid_counts[b] the number of meaningful outputs.
long long ref_ff_agc_cf32(const float* in, const int* blocks, int nblocks, float* out, int* valid_counts,
                          int* nswaps) {
    stream<complex_t> src;
    FeedForwardAGC<complex_t> a(&src);
    // replicate the block's own bookkeeping to know how many swaps to expect
    int inBuffer = 0, swaps = 0;
    std::vector<int> valid;
    for (int b = 0; b < nblocks; b++) {
        inBuffer += blocks[b];
        if (inBuffer < 1024) continue;
        int toProcess = inBuffer - 1024 + 1;
        inBuffer -= toProcess;
        valid.push_back(toProcess);
        swaps++;
    }
    a.start();
    std::thread feeder([&] { feed(&src, (const complex_t*)in, blocks, nblocks); });
    long long total = 0;
    for (int s = 0; s < swaps; s++) {
        int n = a.out.read();
        if (n < 0) break;
        memcpy((complex_t*)out + total, a.out.readBuf, (size_t)valid[s] * sizeof(complex_t));
        a.out.flush();
        if (valid_counts) valid_counts[s] = valid[s];
        total += valid[s];
    }
    feeder.join();
    // the last run() may still be blocked in read(); stop() unblocks it
    a.stop();
    if (nswaps) *nswaps = swaps;
    return total;
}
// CostasLoop<ORDER>: src/dsp/pll.h
long long ref_costas(int order, float loopBandwidth, const float* in, const int* blocks, int nblocks, float* out,
                     double* seconds) {
    switch (order) {
        case 2: return costas_impl<2>(loopBandwidth, in, blocks, nblocks, out, seconds);
        case 4: return costas_impl<4>(loopBandwidth, in, blocks, nblocks, out, seconds);
        case 8: return costas_impl<8>(loopBandwidth, in, blocks, nblocks, out, seconds);
    }
    return -1;
}

}  // extern "C"

// ---- element-wise / layout / per-block-statistic blocks ("next" rows) --------------------------------------------
namespace {
template <class Tin, class Tout, class B>
long long run1(stream<Tin>* src, B& blk, stream<Tout>* out_s, const void* in, const int* blocks, int nblocks, void* out) {
    blk.start();
    long long n = pump(src, out_s, (const Tin*)in, blocks, nblocks, (Tout*)out, nullptr, nullptr);
    blk.stop();
    return n;
}
template <class T, class B>
long long run2(stream<T>* sa, stream<T>* sb, B& blk, const void* a, const void* b, const int* blocks, int nblocks, void* out) {
    blk.start();
    std::thread fa([&] { feed(sa, (const T*)a, blocks, nblocks); });
    std::thread fb([&] { feed(sb, (const T*)b, blocks, nblocks); });
    long long n = drain(&blk.out, (T*)out, nblocks, nullptr);
    fa.join();
    fb.join();
    blk.stop();
    return n;
}
}  // namespace

extern "C" {

// math.h: Add / Substract / Multiply. op 0/1/2; dtype 0 = float, 1 = complex_t
long long ref_math(int op, int dtype, const float* a, const float* b, const int* blocks, int nblocks, float* out) {
    if (dtype == 1) {
        stream<complex_t> sa, sb;
        if (op == 0) { Add<complex_t> k(&sa, &sb); return run2(&sa, &sb, k, a, b, blocks, nblocks, out); }
        if (op == 1) { Substract<complex_t> k(&sa, &sb); return run2(&sa, &sb, k, a, b, blocks, nblocks, out); }
        Multiply<complex_t> k(&sa, &sb);
        return run2(&sa, &sb, k, a, b, blocks, nblocks, out);
    }
    stream<float> sa, sb;
    if (op == 0) { Add<float> k(&sa, &sb); return run2(&sa, &sb, k, a, b, blocks, nblocks, out); }
    if (op == 1) { Substract<float> k(&sa, &sb); return run2(&sa, &sb, k, a, b, blocks, nblocks, out); }
    Multiply<float> k(&sa, &sb);
    return run2(&sa, &sb, k, a, b, blocks, nblocks, out);
}
// audio.h / convertion.h; op numbering = QDSP_LAYOUT_* of include/qdsp_b200.h
long long ref_layout(int op, const float* in0, const float* in1, const int* blocks, int nblocks, float* out0, float* out1) {
    switch (op) {
        case 0: { stream<float> s; MonoToStereo k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
        case 1: {
            stream<float> sl, sr;
            ChannelsToStereo k(&sl, &sr);
            k.start();
            std::thread fa([&] { feed(&sl, in0, blocks, nblocks); });
            std::thread fb([&] { feed(&sr, in1, blocks, nblocks); });
            long long n = drain(&k.out, (stereo_t*)out0, nblocks, nullptr);
            fa.join();
            fb.join();
            k.stop();
            return n;
        }
        case 2: { stream<stereo_t> s; StereoToMono k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
        case 3: {
            stream<stereo_t> s;
            StereoToChannels k(&s);
            k.start();
            std::thread f([&] { feed(&s, (const stereo_t*)in0, blocks, nblocks); });
            long long n = 0;
            for (int b = 0; b < nblocks; b++) {   // out_left and out_right are swapped one after the other (audio.h:176-177)
                int c = k.out_left.read();
                if (c < 0) break;
                memcpy(out0 + n, k.out_left.readBuf, (size_t)c * sizeof(float));
                k.out_left.flush();
                int c2 = k.out_right.read();
                if (c2 < 0) break;
                memcpy(out1 + n, k.out_right.readBuf, (size_t)c2 * sizeof(float));
                k.out_right.flush();
                n += c;
            }
            f.join();
            k.stop();
            return n;
        }
        case 4: { stream<complex_t> s; ComplexToStereo k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
        case 5: { stream<complex_t> s; ComplexToReal k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
        case 6: { stream<complex_t> s; ComplexToImag k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
        case 7: { stream<float> s; RealToComplex k(&s); return run1(&s, k, &k.out, in0, blocks, nblocks, out0); }
    }
    return -1;
}
// Volume<T>: processing.h:348-421. call_set: whether setVolume(volume) is called after construction
long long ref_volume(int dtype, float volume, int call_set, int muted, const float* in, const int* blocks, int nblocks, float* out) {
    if (dtype == 1) {
        stream<stereo_t> s;
        Volume<stereo_t> k(&s, volume);
        if (call_set) k.setVolume(volume);
        k.setMuted(muted != 0);
        return run1(&s, k, &k.out, in, blocks, nblocks, out);
    }
    stream<float> s;
    Volume<float> k(&s, volume);
    if (call_set) k.setVolume(volume);
    k.setMuted(muted != 0);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}
long long ref_threshold(const float* in, const int* blocks, int nblocks, unsigned char* out) {
    stream<float> s;
    Threshold k(&s);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}
long long ref_delay_imag(const float* in, const int* blocks, int nblocks, float* out) {
    stream<complex_t> s;
    DelayImag k(&s);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}
long long ref_amdemod(const float* in, const int* blocks, int nblocks, float* out) {
    stream<complex_t> s;
    AMDemod k(&s);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}
long long ref_squelch(float level, const float* in, const int* blocks, int nblocks, float* out) {
    stream<complex_t> s;
    Squelch k(&s, level);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}
long long ref_ssbdemod(float sampleRate, float bandWidth, int mode, const float* in, const int* blocks, int nblocks,
                       float* out) {
    stream<complex_t> s;
    SSBDemod k(&s, sampleRate, bandWidth, mode);
    return run1(&s, k, &k.out, in, blocks, nblocks, out);
}

// the baked MMSE interpolator table, src/dsp/interpolation_taps.h:6-136 (129 x 8 floats)
void ref_interp_taps(float* out) { memcpy(out, INTERP_TAPS, sizeof(float) * (INTERP_STEPS + 1) * INTERP_TAP_COUNT); }

}  // extern "C"
namespace {
// MMClockRecovery keeps `T delay[1024]` uninitialised (clock_recovery.h:218): construct it in zeroed storage so the
// first seven outputs are deterministic (the oracle and the product define that history as zeros)
template <class T>
long long mm_impl(float omega, float gainOmega, float muGain, float omegaRelLimit, const void* in, const int* blocks,
                  int nblocks, void* out, int* out_counts) {
    stream<T> src;
    void* mem = calloc(1, sizeof(MMClockRecovery<T>));
    MMClockRecovery<T>* k = new (mem) MMClockRecovery<T>(&src, omega, gainOmega, muGain, omegaRelLimit);
    k->start();
    long long n = pump(&src, &k->out, (const T*)in, blocks, nblocks, (T*)out, out_counts, nullptr);
    k->stop();
    k->~MMClockRecovery<T>();
    free(mem);
    return n;
}
template <int ORDER, bool OFFSET>
long long psk_impl(float sampleRate, float baudRate, const float* in, const int* blocks, int nblocks, float* out,
                   int* out_counts) {
    stream<complex_t> src;
    typedef PSKDemod<ORDER, OFFSET> D;
    void* mem = calloc(1, sizeof(D));
    D* d = new (mem) D(&src, sampleRate, baudRate);
    d->start();
    long long n = pump(&src, d->out, (const complex_t*)in, blocks, nblocks, (complex_t*)out, out_counts, nullptr);
    d->stop();
    d->~D();
    free(mem);
    return n;
}
}  // namespace
extern "C" {

// MMClockRecovery<float | complex_t>: src/dsp/clock_recovery.h:68-243
long long ref_mm(int dtype, float omega, float gainOmega, float muGain, float omegaRelLimit, const float* in,
                 const int* blocks, int nblocks, float* out, int* out_counts) {
    return dtype == 1 ? mm_impl<complex_t>(omega, gainOmega, muGain, omegaRelLimit, in, blocks, nblocks, out, out_counts)
                      : mm_impl<float>(omega, gainOmega, muGain, omegaRelLimit, in, blocks, nblocks, out, out_counts);
}
// MSKDemod / PSKDemod<ORDER, OFFSET> hier blocks: src/dsp/demodulator.h:499-682 (defaults of their init())
long long ref_msk_demod(float sampleRate, float deviation, float baudRate, const float* in, const int* blocks, int nblocks,
                        float* out, int* out_counts) {
    stream<complex_t> src;
    void* mem = calloc(1, sizeof(MSKDemod));
    MSKDemod* d = new (mem) MSKDemod(&src, sampleRate, deviation, baudRate);
    d->start();
    long long n = pump(&src, d->out, (const complex_t*)in, blocks, nblocks, out, out_counts, nullptr);
    d->stop();
    d->~MSKDemod();
    free(mem);
    return n;
}
long long ref_psk_demod(int order, int offset, float sampleRate, float baudRate, const float* in, const int* blocks,
                        int nblocks, float* out, int* out_counts) {
    if (order == 2) return psk_impl<2, false>(sampleRate, baudRate, in, blocks, nblocks, out, out_counts);
    if (order == 4 && !offset) return psk_impl<4, false>(sampleRate, baudRate, in, blocks, nblocks, out, out_counts);
    if (order == 4) return psk_impl<4, true>(sampleRate, baudRate, in, blocks, nblocks, out, out_counts);
    if (order == 8) return psk_impl<8, false>(sampleRate, baudRate, in, blocks, nblocks, out, out_counts);
    return -1;
}

}  // extern "C"
