// include/dsp/processing.h — FrequencyXlator<T>, AGC, FeedForwardAGC<T>, ComplexAGC (reference
// src/dsp/processing.h:8-300) and the element-wise helpers of that header: DelayImag, Volume<T>, Squelch, Threshold
// (:300-610). Packer<T> (host-side re-blocking) is not part of this library.
#pragma once
#include <type_traits>
#include <dsp/block.h>

namespace dsp {
    template <class T>
    class FrequencyXlator : public generic_block<FrequencyXlator<T>> {
        using base = generic_block<FrequencyXlator<T>>;
        static_assert(std::is_same<T, complex_t>::value, "the reference implements FrequencyXlator for complex_t only");

    public:
        FrequencyXlator() {}
        FrequencyXlator(stream<complex_t>* in, float sampleRate, float freq) { init(in, sampleRate, freq); }
        ~FrequencyXlator() {
            base::stop();
            if (h) { qdsp_xlator_destroy(h); }
        }
        void init(stream<complex_t>* in, float sampleRate, float freq) {
            _in = in;
            _sampleRate = sampleRate;
            _freq = freq;
            if (h) { qdsp_xlator_destroy(h); }
            h = qdsp_xlator_create(_sampleRate, _freq);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        // (sic) the reference names its setInput "setInputSize" (processing.h:26); both spellings work here
        void setInputSize(stream<complex_t>* in) { setInput(in); }
        void setInput(stream<complex_t>* in) { base::rebindInput(_in, in); }
        void setSampleRate(float sampleRate) {
            _sampleRate = sampleRate;
            qdsp_xlator_set_frequency(h, _sampleRate, _freq);
        }
        float getSampleRate() { return _sampleRate; }
        void setFrequency(float freq) {
            _freq = freq;
            qdsp_xlator_set_frequency(h, _sampleRate, _freq);
        }
        float getFrequency() { return _freq; }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_xlator_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, base::cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        float _sampleRate = 1, _freq = 0;
        stream<complex_t>* _in = nullptr;
        qdsp_xlator* h = nullptr;
    };

    class AGC : public generic_block<AGC> {
    public:
        AGC() {}
        AGC(stream<float>* in, float fallRate, float sampleRate) { init(in, fallRate, sampleRate); }
        ~AGC() {
            generic_block<AGC>::stop();
            if (h) { qdsp_agc_destroy(h); }
        }
        void init(stream<float>* in, float fallRate, float sampleRate) {
            _in = in;
            _sampleRate = sampleRate;
            _fallRate = fallRate;
            rebuild();
            generic_block<AGC>::registerInput(_in);
            generic_block<AGC>::registerOutput(&out);
        }
        void setInput(stream<float>* in) { generic_block<AGC>::rebindInput(_in, in); }
        void setSampleRate(float sampleRate) {
            std::lock_guard<std::mutex> lck(generic_block<AGC>::ctrlMtx);
            _sampleRate = sampleRate;
            if (h) { qdsp_agc_set_params(h, _fallRate, _sampleRate); }   // scalar only (processing.h:101-106): state survives
        }
        void setFallRate(float fallRate) {
            std::lock_guard<std::mutex> lck(generic_block<AGC>::ctrlMtx);
            _fallRate = fallRate;
            if (h) { qdsp_agc_set_params(h, _fallRate, _sampleRate); }
        }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const int one = count;  // one run() call == one block of the decay schedule (processing.h:123)
            const long long n = qdsp_agc_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<float> out;

    private:
        void rebuild() {   // init() only: no worker thread exists yet
            if (h) { qdsp_agc_set_params(h, _fallRate, _sampleRate); return; }
            h = qdsp_agc_create(_fallRate, _sampleRate);
        }
        float _fallRate = 0, _sampleRate = 1;
        stream<float>* _in = nullptr;
        qdsp_agc* h = nullptr;
    };

    template <class T>
    class FeedForwardAGC : public generic_block<FeedForwardAGC<T>> {
        using base = generic_block<FeedForwardAGC<T>>;

    public:
        FeedForwardAGC() {}
        FeedForwardAGC(stream<T>* in) { init(in); }
        ~FeedForwardAGC() {
            base::stop();
            if (h) { qdsp_ffagc_destroy(h); }
        }
        void init(stream<T>* in) {
            _in = in;
            h = qdsp_ffagc_create(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32);
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        // emits the valid (toProcess) outputs; the reference swaps `count` elements of which only toProcess are
        // fresh (processing.h:221) -- consumers here see exactly the fresh ones
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_ffagc_process(h, _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (n == 0) { return count; }
            if (!out.swapDevice((int)n, base::cuStream)) { return -1; }
            return (int)n;
        }

        stream<T> out;

    private:
        stream<T>* _in = nullptr;
        qdsp_ffagc* h = nullptr;
    };

    class ComplexAGC : public generic_block<ComplexAGC> {
    public:
        ComplexAGC() {}
        ComplexAGC(stream<complex_t>* in, float setPoint, float maxGain, float rate) { init(in, setPoint, maxGain, rate); }
        ~ComplexAGC() {
            generic_block<ComplexAGC>::stop();
            if (h) { qdsp_cagc_destroy(h); }
        }
        void init(stream<complex_t>* in, float setPoint, float maxGain, float rate) {
            _in = in;
            _setPoint = setPoint;
            _maxGain = maxGain;
            _rate = rate;
            rebuild();
            generic_block<ComplexAGC>::registerInput(_in);
            generic_block<ComplexAGC>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<ComplexAGC>::rebindInput(_in, in); }
        // like the reference (processing.h:258-269) the setters only change scalars: legal while the worker is in run()
        void setSetPoint(float setPoint) { _setPoint = setPoint; if (h) { qdsp_cagc_set_params(h, _setPoint, _maxGain, _rate); } }
        void setMaxGain(float maxGain) { _maxGain = maxGain; if (h) { qdsp_cagc_set_params(h, _setPoint, _maxGain, _rate); } }
        void setRate(float rate) { _rate = rate; if (h) { qdsp_cagc_set_params(h, _setPoint, _maxGain, _rate); } }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_cagc_process(h, _in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        void rebuild() {   // init() only: no worker thread exists yet
            if (h) { qdsp_cagc_set_params(h, _setPoint, _maxGain, _rate); return; }
            h = qdsp_cagc_create(_setPoint, _maxGain, _rate);
        }
        float _setPoint = 1.0f, _maxGain = 65535.0f, _rate = 1e-3f;
        stream<complex_t>* _in = nullptr;
        qdsp_cagc* h = nullptr;
    };

    // DelayImag (reference processing.h:300-346): out[i] = {in[i].re, in[i-1].im}
    class DelayImag : public generic_block<DelayImag> {
    public:
        DelayImag() {}
        DelayImag(stream<complex_t>* in) { init(in); }
        ~DelayImag() {
            generic_block<DelayImag>::stop();
            if (h) { qdsp_delayimag_destroy(h); }
        }
        void init(stream<complex_t>* in) {
            _in = in;
            if (!h) { h = qdsp_delayimag_create(); }
            generic_block<DelayImag>::registerInput(_in);
            generic_block<DelayImag>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<DelayImag>::rebindInput(_in, in); }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_delayimag_process(h, _in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        stream<complex_t>* _in = nullptr;
        qdsp_delayimag* h = nullptr;
    };

    // Volume<T> (reference processing.h:348-421), T = float or stereo_t. As in the reference, init() stores the
    // volume but the applied level stays 1.0 until setVolume() is called (:355-359 vs :371-374).
    template <class T>
    class Volume : public generic_block<Volume<T>> {
        using base = generic_block<Volume<T>>;
        static_assert(std::is_same<T, float>::value || std::is_same<T, stereo_t>::value, "Volume<float> or Volume<stereo_t>");

    public:
        Volume() {}
        Volume(stream<T>* in, float volume) { init(in, volume); }
        ~Volume() { base::stop(); }
        void init(stream<T>* in, float volume) {
            _in = in;
            _volume = volume;
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        // (sic) the reference names its setInput "setInputSize" (processing.h:362); both spellings work here
        void setInputSize(stream<T>* in) { setInput(in); }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        void setVolume(float volume) {
            _volume = volume;
            level = qdsp_volume_level(_volume);
        }
        float getVolume() { return _volume; }
        void setMuted(bool muted) { _muted = muted; }
        bool getMuted() { return _muted; }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const long long n = qdsp_volume_process(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32, level, _muted ? 1 : 0,
                                                    _in->readDev(), out.writeDev(), count, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, base::cuStream)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        float level = 1.0f;
        float _volume = 1.0f;
        bool _muted = false;
        stream<T>* _in = nullptr;
    };

    // Squelch (reference processing.h:424-489): a run() block passes iff 10*log10f(mean |x|) >= level
    class Squelch : public generic_block<Squelch> {
    public:
        Squelch() {}
        Squelch(stream<complex_t>* in, float level) { init(in, level); }
        ~Squelch() {
            generic_block<Squelch>::stop();
            if (h) { qdsp_squelch_destroy(h); }
        }
        void init(stream<complex_t>* in, float level) {
            _in = in;
            _level = level;
            if (!h) { h = qdsp_squelch_create(_level); }
            qdsp_squelch_set_level(h, _level);
            generic_block<Squelch>::registerInput(_in);
            generic_block<Squelch>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<Squelch>::rebindInput(_in, in); }
        void setLevel(float level) {
            _level = level;
            qdsp_squelch_set_level(h, _level);
        }
        float getLevel() { return _level; }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const int one = count;  // one run() call == one block of the mean (processing.h:465-468)
            const long long n = qdsp_squelch_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<complex_t> out;

    private:
        float _level = -50.0f;
        stream<complex_t>* _in = nullptr;
        qdsp_squelch* h = nullptr;
    };

    // Threshold (reference processing.h:554-610): uint8 stream of (x > 0); setLevel/getLevel exist but run() ignores them
    class Threshold : public generic_block<Threshold> {
    public:
        Threshold() {}
        Threshold(stream<float>* in) { init(in); }
        ~Threshold() { generic_block<Threshold>::stop(); }
        void init(stream<float>* in) {
            _in = in;
            generic_block<Threshold>::registerInput(_in);
            generic_block<Threshold>::registerOutput(&out);
        }
        void setInput(stream<float>* in) { generic_block<Threshold>::rebindInput(_in, in); }
        void setLevel(float level) { _level = level; }
        float getLevel() { return _level; }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_threshold_process(_in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<uint8_t> out;

    private:
        float _level = -50.0f;
        stream<float>* _in = nullptr;
    };
}
