// qdsp_b200/csrc/runtime.cu — runtime plumbing of libqdsp_b200.so: error text, device memory and
// streams (the replacement for volk_malloc-backed stream buffers, reference src/dsp/stream.h:25-31),
// block-partition tables, history tails, NCO bookkeeping and the host-side tap designer.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <numeric>
#include "internal.cuh"

namespace qdsp {

static thread_local char t_err[512] = "";
std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// ---- Partition ------------------------------------------------------------------------------
Partition::~Partition() {
    if (dev) cudaFree(dev);
}

int Partition::build(long long count, const int* blocks, int nblocks, int block_size, int interp, int decim,
                     cudaStream_t s) {
    if (count < 0 || interp <= 0 || decim <= 0) {
        set_last_error("partition: bad arguments");
        return -1;
    }
    view.total = count;
    view.interp = interp;
    view.decim = decim;
    total_out = 0;
    max_out = 0;
    max_count = 0;
    if (blocks == nullptr) {
        if (block_size <= 0) block_size = count > 0 ? (int)(count < 1000000 ? count : 1000000) : 1;
        const long long nb = count == 0 ? 0 : (count + block_size - 1) / block_size;
        view.table = nullptr;
        view.nblocks = (int)nb;
        view.block_size = block_size;
        if (nb > 0) {
            const long long per = ((long long)block_size * interp) / decim;
            const long long last = count - (nb - 1) * block_size;
            const long long last_out = (last * interp) / decim;
            total_out = (nb - 1) * per + last_out;
            max_out = (int)(nb > 1 ? per : last_out);
            if (last_out > max_out) max_out = (int)last_out;
            max_count = (int)(nb > 1 ? block_size : last);
        }
        sizes.clear();
        view.total_out = total_out;
        return 0;
    }
    long long sum = 0;
    for (int b = 0; b < nblocks; b++) {
        if (blocks[b] < 0) {
            set_last_error("partition: negative block size");
            return -1;
        }
        sum += blocks[b];
    }
    if (sum != count) {
        set_last_error("partition: block sizes sum to %lld, count is %lld", sum, count);
        return -1;
    }
    host.resize(nblocks);
    long long in0 = 0, out0 = 0;
    for (int b = 0; b < nblocks; b++) {
        BlkInfo& bi = host[b];
        bi.in_start = in0;
        bi.out_start = out0;
        bi.count = blocks[b];
        bi.out_count = (int)(((long long)blocks[b] * interp) / decim);
        in0 += bi.count;
        out0 += bi.out_count;
        if (bi.out_count > max_out) max_out = bi.out_count;
        if (bi.count > max_count) max_count = bi.count;
    }
    total_out = out0;
    if ((size_t)nblocks > dev_cap) {
        if (dev) cudaFree(dev);
        dev = nullptr;
        dev_cap = 0;
        QDSP_CUDA_OK(cudaMalloc(&dev, sizeof(BlkInfo) * (size_t)nblocks));
        dev_cap = (size_t)nblocks;
    }
    if (nblocks > 0)
        QDSP_CUDA_OK(cudaMemcpyAsync(dev, host.data(), sizeof(BlkInfo) * (size_t)nblocks, cudaMemcpyHostToDevice, s));
    view.table = dev;
    view.total_out = total_out;
    view.nblocks = nblocks;
    view.block_size = 0;
    return 0;
}

// ---- History ----------------------------------------------------------------------------------
int History::init(int H_, int elem_bytes) {
    release();
    H = H_;
    elem = elem_bytes;
    cur = 0;
    const size_t bytes = (size_t)(H > 0 ? H : 1) * elem + 64;
    for (int i = 0; i < 2; i++) {
        QDSP_CUDA_OK(cudaMalloc(&buf[i], bytes));
        QDSP_CUDA_OK(cudaMemset(buf[i], 0, bytes));
    }
    return 0;
}
void History::release() {
    for (int i = 0; i < 2; i++) {
        if (buf[i]) cudaFree(buf[i]);
        buf[i] = nullptr;
    }
}
int History::reset(cudaStream_t s) {
    for (int i = 0; i < 2; i++)
        if (buf[i]) QDSP_CUDA_OK(cudaMemsetAsync(buf[i], 0, (size_t)(H > 0 ? H : 1) * elem, s));
    return 0;
}

// new_hist[j] = virtual[count - H + j], virtual = old_hist ++ in
__global__ void history_advance_kernel(const uint32_t* __restrict__ old_hist, const uint32_t* __restrict__ in,
                                       uint32_t* __restrict__ new_hist, int H, long long count, int words) {
    const long long total = (long long)H * words;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long j = i / words;
        const int w = (int)(i - j * words);
        const long long v = count - H + j;
        new_hist[i] = v >= 0 ? in[v * words + w] : old_hist[(H + v) * words + w];
    }
}
int History::advance(const void* in_dev, long long count, cudaStream_t s) {
    if (H <= 0 || count <= 0) return 0;
    const int words = elem / 4;
    const int nxt = cur ^ 1;
    const long long total = (long long)H * words;
    int grid = (int)((total + 255) / 256);
    if (grid > 1024) grid = 1024;
    history_advance_kernel<<<grid, 256, 0, s>>>((const uint32_t*)buf[cur], (const uint32_t*)in_dev,
                                                (uint32_t*)buf[nxt], H, count, words);
    QDSP_LAUNCH_OK();
    cur = nxt;
    return 0;
}

// ---- NCO ----------------------------------------------------------------------------------------
static const double kTwoPi = 6.283185307179586476925286766559;
static uint64_t turns_from_angle(double ang) {
    double t = ang / kTwoPi;
    t -= floor(t);  // [0,1)
    long double scaled = (long double)t * 18446744073709551616.0L;
    if (scaled >= 18446744073709551615.0L) return 0;
    return (uint64_t)scaled;
}
void Nco::set_freq(float sampleRate, float freq) {
    // reference src/dsp/processing.h:21 — float expression, float cos/sin overloads
    const float a = (freq / sampleRate) * 2.0f * QDSP_FL_M_PI;
    inc_re = cosf(a);
    inc_im = sinf(a);
    step = turns_from_angle(atan2((double)inc_im, (double)inc_re));
}
void Nco::set_inc(float re, float im) {
    inc_re = re;
    inc_im = im;
    step = turns_from_angle(atan2((double)im, (double)re));
}
void Nco::set_phase(float re, float im) { phase = turns_from_angle(atan2((double)im, (double)re)); }
void Nco::get_phase(float* re, float* im) const {
    const double ang = (double)(int64_t)phase * (kTwoPi / 18446744073709551616.0);
    *re = (float)cos(ang);
    *im = (float)sin(ang);
}

}  // namespace qdsp

using namespace qdsp;

// =================================================================================================
// C ABI: runtime
// =================================================================================================
extern "C" {

int qdsp_abi_version(void) { return QDSP_ABI_VERSION; }
const char* qdsp_last_error(void) { return t_err; }
long long qdsp_launch_count(void) { return g_launches.load(); }

int qdsp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        set_last_error("cudaGetDeviceCount failed: no usable CUDA device (this library has no CPU fallback)");
        return 0;
    }
    return n;
}
int qdsp_set_device(int device) {
    QDSP_CUDA_OK(cudaSetDevice(device));
    return 0;
}
int qdsp_get_device(void) {
    int d = -1;
    QDSP_CUDA_OK(cudaGetDevice(&d));
    return d;
}
void* qdsp_malloc_device(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void qdsp_free_device(void* p) {
    if (p) cudaFree(p);
}
void* qdsp_malloc_pinned(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_last_error("cudaHostAlloc(%zu) -> %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void qdsp_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}
int qdsp_memset_device(void* dst, int value, size_t bytes, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return 0;
}
int qdsp_copy_h2d(void* dst, const void* src, size_t bytes, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return 0;
}
int qdsp_copy_d2h(void* dst, const void* src, size_t bytes, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return 0;
}
int qdsp_copy_d2d(void* dst, const void* src, size_t bytes, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return 0;
}
int qdsp_copy_peer(void* dst, int dst_device, const void* src, int src_device, size_t bytes, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, bytes, as_stream(s)));
    return 0;
}
int qdsp_enable_peer_access(int device, int peer) {
    int can = 0;
    QDSP_CUDA_OK(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) {
        set_last_error("device %d cannot access peer %d", device, peer);
        return -1;
    }
    int prev = 0;
    QDSP_CUDA_OK(cudaGetDevice(&prev));
    QDSP_CUDA_OK(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
    }
    cudaSetDevice(prev);
    QDSP_CUDA_OK(e);
    return 0;
}
int qdsp_ipc_export(const void* dev_ptr, void* handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == QDSP_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t hnd;
    QDSP_CUDA_OK(cudaIpcGetMemHandle(&hnd, const_cast<void*>(dev_ptr)));
    memcpy(handle_out, &hnd, sizeof(hnd));
    return 0;
}
void* qdsp_ipc_open(const void* handle) {
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle, sizeof(hnd));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
int qdsp_ipc_close(void* mapped) {
    if (mapped) QDSP_CUDA_OK(cudaIpcCloseMemHandle(mapped));
    return 0;
}
qdsp_stream_t qdsp_stream_create(void) {
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_last_error("cudaStreamCreate -> %s", cudaGetErrorString(e));
        return nullptr;
    }
    return (qdsp_stream_t)s;
}
void qdsp_stream_destroy(qdsp_stream_t s) {
    if (s) cudaStreamDestroy(as_stream(s));
}
int qdsp_stream_sync(qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaStreamSynchronize(as_stream(s)));
    return 0;
}

void* qdsp_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
        set_last_error("cudaEventCreate failed");
        return nullptr;
    }
    return (void*)e;
}
void qdsp_event_destroy(void* ev) {
    if (ev) cudaEventDestroy((cudaEvent_t)ev);
}
int qdsp_event_record(void* ev, qdsp_stream_t s) {
    QDSP_CUDA_OK(cudaEventRecord((cudaEvent_t)ev, as_stream(s)));
    return 0;
}
int qdsp_stream_wait_event(qdsp_stream_t s, void* ev) {
    QDSP_CUDA_OK(cudaStreamWaitEvent(as_stream(s), (cudaEvent_t)ev, 0));
    return 0;
}
int qdsp_event_sync(void* ev) {
    QDSP_CUDA_OK(cudaEventSynchronize((cudaEvent_t)ev));
    return 0;
}

// =================================================================================================
// C ABI: tap design (host). Float expression order follows the reference so the taps are
// bit-identical; this file is compiled for baseline x86-64 (no FMA contraction on the host).
// =================================================================================================

// BlackmanWindow::getTapCount — reference src/dsp/window.h:36-50
int qdsp_blackman_tap_count(float cutoff, float transWidth, float sampleRate) {
    (void)cutoff;
    const float fc = transWidth / sampleRate;
    int count = (int)(4.0f / fc);
    if (count < 4) count = 4;
    if ((count & 1) == 0) count += 1;
    return count;
}

// The reference's "Blackman" factor has no sample index in it (window.h:61-62): one constant.
static float window_constant(float tc) {
    return 0.42f - (0.5f * cosf(2.0f * QDSP_FL_M_PI / tc)) + (0.8f * cosf(4.0f * QDSP_FL_M_PI / tc));
}

// BlackmanWindow::createTaps — reference src/dsp/window.h:52-70
void qdsp_blackman_taps(float cutoff, float transWidth, float sampleRate, float* taps, int tapCount, float factor) {
    (void)transWidth;
    float fc = cutoff / sampleRate;
    if (fc > 1.0f) fc = 1.0f;
    const float tc = (float)tapCount;
    const float half = tc / 2;
    const float w = window_constant(tc);
    float sum = 0.0f;
    for (int k = 0; k < tapCount; k++) {
        const float d = (float)k - half;
        const float v = (sinf(2.0f * QDSP_FL_M_PI * fc * d) / d) * w;
        taps[k] = v;
        sum += v;
    }
    for (int k = 0; k < tapCount; k++) {
        float v = taps[k] * factor;
        taps[k] = v / sum;
    }
}

// BlackmanBandpassWindow::createTaps — reference src/dsp/window.h:120-141
void qdsp_blackman_bandpass_taps(float cutoff, float transWidth, float offset, float sampleRate, float* taps,
                                 int tapCount, float factor) {
    (void)transWidth;
    float fc = cutoff / sampleRate;
    if (fc > 1.0f) fc = 1.0f;
    const float tc = (float)tapCount;
    const float half = tc / 2;
    const float w = window_constant(tc);
    float sum = 0.0f;
    for (int k = 0; k < tapCount; k++) {
        const float d = (float)k - half;
        const float v = (sinf(2.0f * QDSP_FL_M_PI * fc * d) / d) * w;
        taps[k] = v;
        sum += v;
    }
    const float rel = offset / sampleRate;
    for (int k = 0; k < tapCount; k++) {
        float v = taps[k] * cosf(2.0f * rel * QDSP_FL_M_PI * (float)k);
        v *= factor;
        taps[k] = v / sum;
    }
}

// RRCTaps::createTaps — reference src/dsp/window.h:184-229 (double arithmetic on float parameters)
void qdsp_rrc_taps(int tapCount, float sampleRate, float baudRate, float alpha, float* taps) {
    tapCount |= 1;
    const double spb = sampleRate / baudRate;
    const int mid = tapCount / 2;
    double scale = 0;
    for (int k = 0; k < tapCount; k++) {
        const double xi = k - mid;
        const double x1 = QDSP_FL_M_PI * xi / spb;
        double x2 = 4 * alpha * xi / spb;
        double x3 = x2 * x2 - 1;
        double num, den;
        if (fabs(x3) >= 0.000001) {
            num = (k != mid) ? cos((1 + alpha) * x1) + sin((1 - alpha) * x1) / (4 * alpha * xi / spb)
                             : cos((1 + alpha) * x1) + (1 - alpha) * QDSP_FL_M_PI / (4 * alpha);
            den = x3 * QDSP_FL_M_PI;
        } else {
            if (alpha == 1) {
                taps[k] = -1;
                scale += taps[k];
                continue;
            }
            x3 = (1 - alpha) * x1;
            x2 = (1 + alpha) * x1;
            num = (sin(x2) * (1 + alpha) * QDSP_FL_M_PI - cos(x3) * ((1 - alpha) * QDSP_FL_M_PI * spb) / (4 * alpha * xi) +
                   sin(x3) * spb * spb / (4 * alpha * xi * xi));
            den = -32 * QDSP_FL_M_PI * alpha * alpha * xi / spb;
        }
        taps[k] = (float)(4 * alpha * num / den);
        scale += taps[k];
    }
    for (int k = 0; k < tapCount; k++) taps[k] = (float)(taps[k] / scale);
}

// PolyphaseResampler::init — reference src/dsp/resampling.h:28-30
void qdsp_rates_to_ratio(float inSampleRate, float outSampleRate, int* interp, int* decim) {
    const int a = (int)inSampleRate, b = (int)outSampleRate;
    const int g = std::gcd(a, b);
    *interp = (int)(outSampleRate / (float)g);
    *decim = (int)(inSampleRate / (float)g);
}

// VFO::init — reference src/dsp/vfo.h:26-33 + resampling.h:83-93 (updateWindow: gain factor = interp)
int qdsp_vfo_design(float inSampleRate, float outSampleRate, float bandWidth, float* taps, int maxTaps, int* interp,
                    int* decim) {
    float lim = inSampleRate < outSampleRate ? inSampleRate : outSampleRate;
    if (bandWidth < lim) lim = bandWidth;
    const float realCutoff = lim / 2.0f;
    int I, D;
    qdsp_rates_to_ratio(inSampleRate, outSampleRate, &I, &D);
    const float winRate = inSampleRate * (float)I;
    const int n = qdsp_blackman_tap_count(realCutoff, realCutoff, winRate);
    if (interp) *interp = I;
    if (decim) *decim = D;
    if (taps && maxTaps >= n) qdsp_blackman_taps(realCutoff, realCutoff, winRate, taps, n, (float)I);
    return n;
}

// PolyphaseResampler::run index schedule — reference src/dsp/resampling.h:121-125
int qdsp_resamp_schedule(int interp, int decim, int count, int* phase, int* index) {
    const int outCount = (int)(((long long)count * interp) / decim);
    long long i = 0;
    for (int k = 0; k < outCount; k++, i += decim) {
        if (phase) phase[k] = (int)(i % interp);
        if (index) index[k] = (int)(i / interp);
    }
    return outCount;
}

}  // extern "C"
