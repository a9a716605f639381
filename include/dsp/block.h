// include/dsp/block.h — generic_block / generic_hier_block: the thread-per-block driver of the dsp:: mirror.
// Public surface as in the reference (src/dsp/block.h:13-208): start(), stop(), run(), calcOutSize(),
// `ctrlMtx`-guarded setters with tempStop()/tempStart() bracketing. Each block additionally owns one CUDA
// stream (`cuStream`) on which its run() enqueues kernels; hand-off to the next block is stream<T>'s job.
#pragma once
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <dsp/stream.h>
#include <dsp/types.h>

namespace dsp {
    class generic_unnamed_block {
    public:
        virtual ~generic_unnamed_block() {}
        virtual void start() {}
        virtual void stop() {}
        virtual int calcOutSize(int inSize) { return inSize; }
        virtual int run() { return -1; }
    };

    template <class BLOCK>
    class generic_block : public generic_unnamed_block {
    public:
        generic_block() {
            cuDevice = qdsp_get_device();      // streams, buffers and handles live on the constructing thread's device
            cuStream = qdsp_stream_create();
        }
        virtual ~generic_block() {
            stop();
            qdsp_stream_destroy(cuStream);
        }
        virtual void init() {}

        virtual void start() override {
            std::lock_guard<std::mutex> lck(ctrlMtx);
            if (running) { return; }
            running = true;
            doStart();
        }
        virtual void stop() override {
            std::lock_guard<std::mutex> lck(ctrlMtx);
            if (!running) { return; }
            doStop();
            running = false;
        }
        virtual int calcOutSize(int inSize) override { return inSize; }
        virtual int run() override = 0;

        friend BLOCK;

    protected:
        // setInput() of every block: pause the worker, swap the registered input stream, resume
        template <class S>
        void rebindInput(S*& slot, S* in) {
            std::lock_guard<std::mutex> lck(ctrlMtx);
            tempStop();
            unregisterInput(slot);
            slot = in;
            registerInput(slot);
            tempStart();
        }
        void registerInput(untyped_steam* s) { inputs.push_back(s); }
        void unregisterInput(untyped_steam* s) { inputs.erase(std::remove(inputs.begin(), inputs.end(), s), inputs.end()); }
        void registerOutput(untyped_steam* s) { outputs.push_back(s); }
        void unregisterOutput(untyped_steam* s) { outputs.erase(std::remove(outputs.begin(), outputs.end(), s), outputs.end()); }

        virtual void doStart() {
            worker = std::thread([this] {
                if (cuDevice >= 0) { qdsp_set_device(cuDevice); }   // a new host thread starts on device 0
                while (run() >= 0) {}
            });
        }
        // cooperative stop: wake the worker out of read()/swap(), join, re-arm the streams
        virtual void doStop() {
            for (untyped_steam* s : inputs) { s->stopReader(); }
            for (untyped_steam* s : outputs) { s->stopWriter(); }
            if (worker.joinable()) { worker.join(); }
            qdsp_stream_sync(cuStream);
            for (untyped_steam* s : inputs) { s->clearReadStop(); }
            for (untyped_steam* s : outputs) { s->clearWriteStop(); }
        }
        void tempStart() {
            if (!paused) { return; }
            doStart();
            paused = false;
        }
        void tempStop() {
            if (!running || paused) { return; }
            doStop();
            paused = true;
        }

        std::vector<untyped_steam*> inputs, outputs;
        bool running = false;
        bool paused = false;
        std::thread worker;
        qdsp_stream_t cuStream = nullptr;
        int cuDevice = -1;
        std::mutex ctrlMtx;
    };

    template <class BLOCK>
    class generic_hier_block {
    public:
        virtual ~generic_hier_block() { stop(); }
        virtual void init() {}
        virtual void start() {
            std::lock_guard<std::mutex> lck(ctrlMtx);
            if (running) { return; }
            running = true;
            for (generic_unnamed_block* b : blocks) { b->start(); }
        }
        virtual void stop() {
            std::lock_guard<std::mutex> lck(ctrlMtx);
            if (!running) { return; }
            for (generic_unnamed_block* b : blocks) { b->stop(); }
            running = false;
        }
        virtual int calcOutSize(int inSize) { return inSize; }

        friend BLOCK;

    protected:
        void registerBlock(generic_unnamed_block* b) { blocks.push_back(b); }
        void unregisterBlock(generic_unnamed_block* b) { blocks.erase(std::remove(blocks.begin(), blocks.end(), b), blocks.end()); }
        std::vector<generic_unnamed_block*> blocks;
        bool running = false;
        std::mutex ctrlMtx;
    };
}
