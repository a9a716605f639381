// include/dsp/source.h — HandlerSource<T> (reference src/dsp/source.h:66-107): a user callback fills
// `out.writeBuf` (pinned host memory); the first device block downstream uploads it once.
#pragma once
#include <dsp/block.h>

namespace dsp {
    template <class T>
    class HandlerSource : public generic_block<HandlerSource<T>> {
        using base = generic_block<HandlerSource<T>>;

    public:
        HandlerSource() {}
        HandlerSource(int (*handler)(T* data, void* ctx), void* ctx) { init(handler, ctx); }
        ~HandlerSource() { base::stop(); }
        void init(int (*handler)(T* data, void* ctx), void* ctx) {
            _handler = handler;
            _ctx = ctx;
            base::registerOutput(&out);
        }
        void setHandler(int (*handler)(T* data, void* ctx), void* ctx) {
            std::lock_guard<std::mutex> lck(base::ctrlMtx);
            base::tempStop();
            _handler = handler;
            _ctx = ctx;
            base::tempStart();
        }
        int run() override {
            const int count = _handler(out.writeBuf, _ctx);
            if (count < 0) { return -1; }
            if (!out.swap(count)) { return -1; }
            return count;
        }

        stream<T> out;

    private:
        int (*_handler)(T* data, void* ctx) = nullptr;
        void* _ctx = nullptr;
    };
}
