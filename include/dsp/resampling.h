// include/dsp/resampling.h — PolyphaseResampler<T> and PowerDecimator (reference src/dsp/resampling.h:9-258).
#pragma once
#include <numeric>
#include <type_traits>
#include <vector>
#include <dsp/block.h>
#include <dsp/window.h>

namespace dsp {
    template <class T>
    class PolyphaseResampler : public generic_block<PolyphaseResampler<T>> {
        using base = generic_block<PolyphaseResampler<T>>;

    public:
        PolyphaseResampler() {}
        PolyphaseResampler(stream<T>* in, dsp::filter_window::generic_window* window, float inSampleRate, float outSampleRate) {
            init(in, window, inSampleRate, outSampleRate);
        }
        ~PolyphaseResampler() {
            base::stop();
            if (h) { qdsp_resamp_destroy(h); }
        }

        void init(stream<T>* in, dsp::filter_window::generic_window* window, float inSampleRate, float outSampleRate) {
            _in = in;
            _window = window;
            _inSampleRate = inSampleRate;
            _outSampleRate = outSampleRate;
            qdsp_rates_to_ratio(_inSampleRate, _outSampleRate, &_interp, &_decim);
            rebuild();
            base::registerInput(_in);
            base::registerOutput(&out);
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        // like the reference (resampling.h:53-73) the rate setters recompute I/D but keep the current taps
        // until updateWindow() is called
        void setInSampleRate(float inSampleRate) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            _inSampleRate = inSampleRate;
            qdsp_rates_to_ratio(_inSampleRate, _outSampleRate, &_interp, &_decim);
            rebuild();
            base::tempStart();
        }
        void setOutSampleRate(float outSampleRate) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            _outSampleRate = outSampleRate;
            qdsp_rates_to_ratio(_inSampleRate, _outSampleRate, &_interp, &_decim);
            rebuild();
            base::tempStart();
        }
        int getInterpolation() { return _interp; }
        int getDecimation() { return _decim; }
        void updateWindow(dsp::filter_window::generic_window* window) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            _window = window;
            rebuild();
            base::tempStart();
        }
        int calcOutSize(int in) override { return (int)qdsp_resamp_out_count(h, in); }

        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(base::cuStream);
            const int one = count;
            const long long n = qdsp_resamp_process(h, _in->readDev(), out.writeDev(), count, &one, 1, 0, nullptr, base::cuStream);
            _in->flushDevice(base::cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice((int)n, base::cuStream)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        void rebuild() {
            const int tapCount = _window->getTapCount();
            // RRCTaps::createTaps rounds an even count up and writes tapCount | 1 taps (window.h:186): the reference
            // overflows its buffer by one float there and filters with the first tapCount of them; same taps, no overflow
            std::vector<float> taps((size_t)tapCount | 1);
            _window->createTaps(taps.data(), tapCount, (float)_interp);
            if (h) { qdsp_resamp_destroy(h); }
            h = qdsp_resamp_create(std::is_same<T, float>::value ? QDSP_F32 : QDSP_CF32, taps.data(), tapCount, _interp, _decim);
        }
        stream<T>* _in = nullptr;
        dsp::filter_window::generic_window* _window = nullptr;
        int _interp = 1, _decim = 1;
        float _inSampleRate = 1, _outSampleRate = 1;
        qdsp_resamp* h = nullptr;
    };

    class PowerDecimator : public generic_block<PowerDecimator> {
    public:
        PowerDecimator() {}
        PowerDecimator(stream<complex_t>* in, unsigned int power) { init(in, power); }
        void init(stream<complex_t>* in, unsigned int power) {
            _in = in;
            _power = power;
            generic_block<PowerDecimator>::registerInput(_in);
            generic_block<PowerDecimator>::registerOutput(&out);
        }
        void setInput(stream<complex_t>* in) { generic_block<PowerDecimator>::rebindInput(_in, in); }
        void setPower(unsigned int power) {
            std::lock_guard<std::mutex> lck(generic_block<PowerDecimator>::ctrlMtx);
            generic_block<PowerDecimator>::tempStop();
            _power = power;
            generic_block<PowerDecimator>::tempStart();
        }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_power_decim_process(_power, _in->readDev(), out.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice((int)n, cuStream)) { return -1; }
            return (int)n;
        }

        stream<complex_t> out;

    private:
        unsigned int _power = 0;
        stream<complex_t>* _in = nullptr;
    };
}
