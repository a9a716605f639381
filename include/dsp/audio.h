// include/dsp/audio.h — MonoToStereo, ChannelsToStereo, StereoToMono, StereoToChannels (reference
// src/dsp/audio.h:5-187): layout shuffles between float and stereo_t streams, one streaming kernel each.
#pragma once
#include <dsp/block.h>

namespace dsp {
    namespace detail {
        // one input stream, one output stream, one QDSP_LAYOUT_* op
        template <class SELF, class TIN, class TOUT, int OP>
        class layout_block : public generic_block<SELF> {
            using base = generic_block<SELF>;

        public:
            ~layout_block() { base::stop(); }
            void init(stream<TIN>* in) {
                _in = in;
                base::registerInput(_in);
                base::registerOutput(&out);
            }
            void setInput(stream<TIN>* in) { base::rebindInput(_in, in); }
            int run() override {
                const int count = _in->readDevice(base::cuStream);
                if (count < 0) { return -1; }
                out.acquireWriteDev(base::cuStream);
                const long long n = qdsp_layout_process(OP, _in->readDev(), nullptr, out.writeDev(), nullptr, count, base::cuStream);
                _in->flushDevice(base::cuStream);
                if (n < 0) { return -1; }
                if (!out.swapDevice(count, base::cuStream)) { return -1; }
                return count;
            }

            stream<TOUT> out;

        private:
            stream<TIN>* _in = nullptr;
        };
    }

    class MonoToStereo : public detail::layout_block<MonoToStereo, float, stereo_t, QDSP_LAYOUT_MONO_TO_STEREO> {
    public:
        MonoToStereo() {}
        MonoToStereo(stream<float>* in) { init(in); }
    };

    class StereoToMono : public detail::layout_block<StereoToMono, stereo_t, float, QDSP_LAYOUT_STEREO_TO_MONO> {
    public:
        StereoToMono() {}
        StereoToMono(stream<stereo_t>* in) { init(in); }
    };

    class ChannelsToStereo : public generic_block<ChannelsToStereo> {
    public:
        ChannelsToStereo() {}
        ChannelsToStereo(stream<float>* in_left, stream<float>* in_right) { init(in_left, in_right); }
        ~ChannelsToStereo() { generic_block<ChannelsToStereo>::stop(); }
        void init(stream<float>* in_left, stream<float>* in_right) {
            _in_left = in_left;
            _in_right = in_right;
            generic_block<ChannelsToStereo>::registerInput(_in_left);
            generic_block<ChannelsToStereo>::registerInput(_in_right);
            generic_block<ChannelsToStereo>::registerOutput(&out);
        }
        void setInput(stream<float>* in_left, stream<float>* in_right) {
            std::lock_guard<std::mutex> lck(generic_block<ChannelsToStereo>::ctrlMtx);
            generic_block<ChannelsToStereo>::tempStop();
            generic_block<ChannelsToStereo>::unregisterInput(_in_left);
            generic_block<ChannelsToStereo>::unregisterInput(_in_right);
            _in_left = in_left;
            _in_right = in_right;
            generic_block<ChannelsToStereo>::registerInput(_in_left);
            generic_block<ChannelsToStereo>::registerInput(_in_right);
            generic_block<ChannelsToStereo>::tempStart();
        }
        int run() override {
            const int count_l = _in_left->readDevice(cuStream);
            if (count_l < 0) { return -1; }
            const int count_r = _in_right->readDevice(cuStream);
            if (count_r < 0) { return -1; }
            // the reference only warns on a size mismatch and interleaves count_l elements (audio.h:76-80)
            out.acquireWriteDev(cuStream);
            const long long n = qdsp_layout_process(QDSP_LAYOUT_CHANNELS_TO_STEREO, _in_left->readDev(), _in_right->readDev(),
                                                    out.writeDev(), nullptr, count_l, cuStream);
            _in_left->flushDevice(cuStream);
            _in_right->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out.swapDevice(count_l, cuStream)) { return -1; }
            return count_l;
        }

        stream<stereo_t> out;

    private:
        stream<float>* _in_left = nullptr;
        stream<float>* _in_right = nullptr;
    };

    class StereoToChannels : public generic_block<StereoToChannels> {
    public:
        StereoToChannels() {}
        StereoToChannels(stream<stereo_t>* in) { init(in); }
        ~StereoToChannels() { generic_block<StereoToChannels>::stop(); }
        void init(stream<stereo_t>* in) {
            _in = in;
            generic_block<StereoToChannels>::registerInput(_in);
            generic_block<StereoToChannels>::registerOutput(&out_left);
            generic_block<StereoToChannels>::registerOutput(&out_right);
        }
        void setInput(stream<stereo_t>* in) { generic_block<StereoToChannels>::rebindInput(_in, in); }
        int run() override {
            const int count = _in->readDevice(cuStream);
            if (count < 0) { return -1; }
            out_left.acquireWriteDev(cuStream);
            out_right.acquireWriteDev(cuStream);
            const long long n = qdsp_layout_process(QDSP_LAYOUT_STEREO_TO_CHANNELS, _in->readDev(), nullptr, out_left.writeDev(),
                                                    out_right.writeDev(), count, cuStream);
            _in->flushDevice(cuStream);
            if (n < 0) { return -1; }
            if (!out_left.swapDevice(count, cuStream)) { return -1; }
            if (!out_right.swapDevice(count, cuStream)) { return -1; }
            return count;
        }

        stream<float> out_left;
        stream<float> out_right;

    private:
        stream<stereo_t>* _in = nullptr;
    };
}
