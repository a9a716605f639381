// include/dsp/window.h — filter_window::{generic_window, BlackmanWindow, BlackmanBandpassWindow, RRCTaps}.
// Same interface as the reference (src/dsp/window.h:7-231); the arithmetic lives in libqdsp_b200.so's host-side
// tap designer, which is bit-exact with the reference's float expressions (tests/test_host.py).
#pragma once
#include <qdsp_b200.h>

namespace dsp {
    namespace filter_window {
        class generic_window {
        public:
            virtual ~generic_window() {}
            virtual int getTapCount() { return -1; }
            virtual void createTaps(float* taps, int tapCount, float factor = 1.0f) {}
        };

        class BlackmanWindow : public filter_window::generic_window {
        public:
            BlackmanWindow() {}
            BlackmanWindow(float cutoff, float transWidth, float sampleRate) { init(cutoff, transWidth, sampleRate); }
            void init(float cutoff, float transWidth, float sampleRate) {
                _cutoff = cutoff;
                _transWidth = transWidth;
                _sampleRate = sampleRate;
            }
            void setSampleRate(float sampleRate) { _sampleRate = sampleRate; }
            void setCutoff(float cutoff) { _cutoff = cutoff; }
            void setTransWidth(float transWidth) { _transWidth = transWidth; }
            int getTapCount() override { return qdsp_blackman_tap_count(_cutoff, _transWidth, _sampleRate); }
            void createTaps(float* taps, int tapCount, float factor = 1.0f) override {
                qdsp_blackman_taps(_cutoff, _transWidth, _sampleRate, taps, tapCount, factor);
            }

        protected:
            float _cutoff = 0, _transWidth = 0, _sampleRate = 1;
        };

        class BlackmanBandpassWindow : public BlackmanWindow {
        public:
            BlackmanBandpassWindow() {}
            BlackmanBandpassWindow(float cutoff, float transWidth, float offset, float sampleRate) { init(cutoff, transWidth, offset, sampleRate); }
            void init(float cutoff, float transWidth, float offset, float sampleRate) {
                BlackmanWindow::init(cutoff, transWidth, sampleRate);
                _offset = offset;
            }
            void setOffset(float offset) { _offset = offset; }
            void createTaps(float* taps, int tapCount, float factor = 1.0f) override {
                qdsp_blackman_bandpass_taps(_cutoff, _transWidth, _offset, _sampleRate, taps, tapCount, factor);
            }

        private:
            float _offset = 0;
        };
    }

    // (the reference declares RRCTaps in namespace dsp, outside filter_window: window.h:149)
    class RRCTaps : public filter_window::generic_window {
    public:
        RRCTaps() {}
        RRCTaps(int tapCount, float sampleRate, float baudRate, float alpha) { init(tapCount, sampleRate, baudRate, alpha); }
        void init(int tapCount, float sampleRate, float baudRate, float alpha) {
            _tapCount = tapCount;
            _sampleRate = sampleRate;
            _baudRate = baudRate;
            _alpha = alpha;
        }
        int getTapCount() override { return _tapCount; }
        void setSampleRate(float sampleRate) { _sampleRate = sampleRate; }
        void setTapCount(int count) { _tapCount = count; }
        void setBaudRate(float baudRate) { _baudRate = baudRate; }
        void setAlpha(float alpha) { _alpha = alpha; }
        void createTaps(float* taps, int tapCount, float factor = 1.0f) override {
            qdsp_rrc_taps(tapCount, _sampleRate, _baudRate, _alpha, taps);
        }

    private:
        int _tapCount = 0;
        float _sampleRate = 1, _baudRate = 1, _alpha = 0.35f;
    };

    namespace filter_window {
        using dsp::RRCTaps;   // the spelling earlier versions of this header offered
    }
}
