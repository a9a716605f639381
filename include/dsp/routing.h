// include/dsp/routing.h — Splitter<T> (reference src/dsp/routing.h:9-66): fan one stream out to every bound
// stream. Here the fan-out is a device-to-device copy per bound stream (the channelizer entry point
// qdsp_channelizer_process avoids even that: all channels read the same device buffer).
#pragma once
#include <algorithm>
#include <vector>
#include <dsp/block.h>

namespace dsp {
    template <class T>
    class Splitter : public generic_block<Splitter<T>> {
        using base = generic_block<Splitter<T>>;

    public:
        Splitter() {}
        Splitter(stream<T>* in) { init(in); }
        ~Splitter() { base::stop(); }
        void init(stream<T>* in) {
            _in = in;
            base::registerInput(_in);
        }
        void setInput(stream<T>* in) { base::rebindInput(_in, in); }
        void bindStream(stream<T>* stream) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            out.push_back(stream);
            base::registerOutput(stream);
            base::tempStart();
        }
        void unbindStream(stream<T>* stream) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            base::unregisterOutput(stream);
            out.erase(std::remove(out.begin(), out.end(), stream), out.end());
            base::tempStart();
        }
        int run() override {
            const int count = _in->readDevice(base::cuStream);
            if (count < 0) { return -1; }
            for (stream<T>* s : out) {
                s->acquireWriteDev(base::cuStream);
                qdsp_copy_d2d(s->writeDev(), _in->readDev(), (size_t)count * sizeof(T), base::cuStream);
                if (!s->swapDevice(count, base::cuStream)) { return -1; }
            }
            _in->flushDevice(base::cuStream);
            return count;
        }

    private:
        stream<T>* _in = nullptr;
        std::vector<stream<T>*> out;
    };
}
