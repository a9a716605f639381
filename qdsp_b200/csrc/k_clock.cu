// qdsp_b200/csrc/k_clock.cu — MMClockRecovery<float | complex_t>::run (reference src/dsp/clock_recovery.h:127-215):
// Mueller & Mueller symbol-timing recovery with an 8-tap, 129-phase interpolator table.
//
// The recurrence is data dependent in its INDEX (which input sample the next symbol is interpolated at depends on the
// phase error of the previous symbol), so this first implementation keeps the reference's order exactly: one warp per
// stream stages windows of the input in shared memory with coalesced loads, lane 0 walks the symbols. Every float
// operation is rounded where the x86 reference rounds it (sequential dot product, no contraction): outputs, per-block
// output counts and the carried state are bit-identical to the reference. Throughput is latency bound (about one
// symbol per 150-200 cycles); independent streams scale across CTAs (one per stream), a chunk-speculative variant for
// a single long stream is the follow-up.
#include "internal.cuh"
#include "kernels.cuh"

namespace qdsp {

constexpr int kMmWin = 4096;     // samples staged per window
constexpr int kMmHalo = 8;       // an output at sample i reads x[i-7 .. i]

__device__ __forceinline__ float mm_step(float v) { return v > 0.0f ? 1.0f : -1.0f; }   // DSP_STEP, utils/macros.h:6

// state layout == oracle/port.c: [0] mu [1] dynOmega [2] lastOutput [3..8] p0 p1 p2 [9..14] c0 c1 c2 [15] nextOffset
// [16..29] delay[0..6] as (re, im) pairs (float streams use the re slots)
template <bool CPLX>
__global__ void __launch_bounds__(32) mm_kernel(const float* __restrict__ in, PartitionDev part, const float* __restrict__ taps_g,
                                                float omega, float gainOmega, float muGain, float omegaMin, float omegaMax,
                                                float* __restrict__ state, float* __restrict__ out, int* __restrict__ out_counts,
                                                long long* __restrict__ total_out) {
    constexpr int ES = CPLX ? 2 : 1;
    __shared__ float s_taps[129 * 8];
    __shared__ float s_x[(kMmWin + kMmHalo) * ES];
    const int lane = threadIdx.x;
    for (int k = lane; k < 129 * 8; k += 32) s_taps[k] = taps_g[k];

    float mu = state[0], dynOmega = state[1], lastOutput = state[2];
    float2 p0 = make_float2(state[3], state[4]), p1 = make_float2(state[5], state[6]), p2 = make_float2(state[7], state[8]);
    float2 c0 = make_float2(state[9], state[10]), c1 = make_float2(state[11], state[12]), c2 = make_float2(state[13], state[14]);
    int nextOffset = (int)state[15];
    long long total = 0;

    for (int b = 0; b < part.nblocks; b++) {
        const BlkInfo bi = part.get(b);
        const int count = bi.count;
        const long long S = bi.in_start;
        const int maxOut = (int)__fmul_rn(__fmul_rn(2.0f, omega), (float)count);   // clock_recovery.h:135
        int outCount = 0;
        int i = nextOffset < 0 ? 0 : nextOffset;   // negative only after the maxOut cap fired (the reference then reads delay[i < 0])
        // windows of the block: samples [w0 - 7, w0 + kMmWin) of the block-relative stream; what lies before the call's
        // first sample comes from the carried delay[] (the previous call's last 7 samples)
        for (int w0 = i / kMmWin * kMmWin; w0 < count; w0 += kMmWin) {
            __syncwarp();
            // 8 loads per lane in flight together (a load-store loop would expose the DRAM latency 128 times per window)
            for (int e0 = lane; e0 < kMmWin + kMmHalo - 1; e0 += 8 * 32) {
                float2 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = e0 + 32 * j;
                    const long long g = S + w0 - 7 + e;             // call-relative sample index
                    v[j] = make_float2(0.f, 0.f);
                    if (e < kMmWin + kMmHalo - 1) {
                        if (g >= 0) {
                            if (g < part.total) {
                                if (CPLX) v[j] = reinterpret_cast<const float2*>(in)[g];
                                else v[j].x = in[g];
                            }
                        } else {
                            v[j] = make_float2(state[16 + 2 * (int)(g + 7)], state[17 + 2 * (int)(g + 7)]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = e0 + 32 * j;
                    if (e < kMmWin + kMmHalo - 1) {
                        if (CPLX) reinterpret_cast<float2*>(s_x)[e] = v[j];
                        else s_x[e] = v[j].x;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                const int wend = w0 + kMmWin < count ? w0 + kMmWin : count;
                while (i < wend && outCount < maxOut) {
                    const float* t8 = s_taps + 8 * (int)roundf(__fmul_rn(mu, 128.0f));
                    const float* src = s_x + (i - w0) * ES;        // x[i-7 .. i]
                    float phaseError;
                    if (!CPLX) {
                        float outVal = 0.0f;
#pragma unroll
                        for (int k = 0; k < 8; k++) outVal = __fadd_rn(outVal, __fmul_rn(src[k], t8[k]));
                        out[total + outCount] = outVal;
                        outCount++;
                        phaseError = __fsub_rn(__fmul_rn(mm_step(lastOutput), outVal), __fmul_rn(lastOutput, mm_step(outVal)));
                        lastOutput = outVal;
                    } else {
                        p2 = p1;
                        p1 = p0;
                        c2 = c1;
                        c1 = c0;
                        float re = 0.0f, im = 0.0f;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            re = __fadd_rn(re, __fmul_rn(src[2 * k], t8[k]));
                            im = __fadd_rn(im, __fmul_rn(src[2 * k + 1], t8[k]));
                        }
                        p0 = make_float2(re, im);
                        reinterpret_cast<float2*>(out)[total + outCount] = p0;
                        outCount++;
                        c0 = make_float2(mm_step(p0.x), mm_step(p0.y));
                        // (((p0 - p2) * conj(c1)) - ((c0 - c2) * conj(p1))).re, clock_recovery.h:181 with types.h:17-27
                        const float ar = __fsub_rn(p0.x, p2.x), ai = __fsub_rn(p0.y, p2.y);
                        const float dr = __fsub_rn(c0.x, c2.x), di = __fsub_rn(c0.y, c2.y);
                        const float ab = __fsub_rn(__fmul_rn(ar, c1.x), __fmul_rn(ai, -c1.y));
                        const float de = __fsub_rn(__fmul_rn(dr, p1.x), __fmul_rn(di, -p1.y));
                        phaseError = __fsub_rn(ab, de);
                    }
                    if (phaseError > 1.0f) phaseError = 1.0f;
                    if (phaseError < -1.0f) phaseError = -1.0f;
                    dynOmega = __fadd_rn(dynOmega, __fmul_rn(gainOmega, phaseError));
                    if (dynOmega > omegaMax) dynOmega = omegaMax;
                    else if (dynOmega < omegaMin) dynOmega = omegaMin;
                    mu = __fadd_rn(__fadd_rn(mu, dynOmega), __fmul_rn(muGain, phaseError));
                    const float roundedStep = floorf(mu);
                    i += (int)roundedStep;
                    if (i < 0) i = 0;
                    mu = __fsub_rn(mu, roundedStep);
                }
            }
            i = __shfl_sync(0xffffffffu, i, 0);
            outCount = __shfl_sync(0xffffffffu, outCount, 0);
            if (outCount >= maxOut) break;
            if (i >= w0 + 2 * kMmWin) w0 = i / kMmWin * kMmWin - kMmWin;   // a step longer than a window (huge omega)
        }
        nextOffset = i - count;                                    // clock_recovery.h:208
        if (lane == 0 && out_counts) out_counts[b] = outCount;
        total += outCount;
    }
    __syncwarp();
    // the call's last 7 samples become delay[0..6] (clock_recovery.h:211); shorter calls shift the old tail
    if (lane < 7) {
        const long long g = part.total - 7 + lane;
        float2 v;
        if (g >= 0) {
            if (CPLX) v = reinterpret_cast<const float2*>(in)[g];
            else v = make_float2(in[g], 0.f);
        } else {
            v = make_float2(state[16 + 2 * (int)(g + 7)], state[17 + 2 * (int)(g + 7)]);
        }
        __syncwarp(0x7f);
        state[16 + 2 * lane] = v.x;
        state[17 + 2 * lane] = v.y;
    }
    if (lane == 0) {
        state[0] = mu;
        state[1] = dynOmega;
        state[2] = lastOutput;
        state[3] = p0.x; state[4] = p0.y; state[5] = p1.x; state[6] = p1.y; state[7] = p2.x; state[8] = p2.y;
        state[9] = c0.x; state[10] = c0.y; state[11] = c1.x; state[12] = c1.y; state[13] = c2.x; state[14] = c2.y;
        state[15] = (float)nextOffset;
        *total_out = total;
    }
}

// =================================================================================================
// Speculate and verify (single long stream): one THREAD per chunk of kSpecChunk samples. Chunk 0 continues the
// carried state; every later chunk starts kSpecWarm samples before its boundary from the default loop state (the
// timing loop is contractive: after the warm-up its sampling instants have merged with the true trajectory's), and
// records the loop state it ENTERS its chunk with. A single thread then walks the boundaries: a chunk is accepted
// iff that entry state (symbol index + every float of the loop state) is bit-equal to the state its -- already
// verified -- predecessor LEAVES with; otherwise the chunk is re-walked from the true state. The accepted outputs
// are exactly the sequential ones by induction, whatever the speculation did. A compaction pass moves the chunk
// outputs to their final positions and counts the symbols per run() block from their source indices.
// =================================================================================================
struct MmLoop {            // the loop state BEFORE the symbol at sample index i is produced
    long long i;
    float mu, dynOmega, lastOutput;
    float2 p0, p1, p2, c0, c1, c2;
};
__device__ __forceinline__ bool mm_same(const MmLoop& a, const MmLoop& b) {
    auto eq = [](float x, float y) { return __float_as_uint(x) == __float_as_uint(y); };
    return a.i == b.i && eq(a.mu, b.mu) && eq(a.dynOmega, b.dynOmega) && eq(a.lastOutput, b.lastOutput) &&
           eq(a.p0.x, b.p0.x) && eq(a.p0.y, b.p0.y) && eq(a.p1.x, b.p1.x) && eq(a.p1.y, b.p1.y) && eq(a.p2.x, b.p2.x) &&
           eq(a.p2.y, b.p2.y) && eq(a.c0.x, b.c0.x) && eq(a.c0.y, b.c0.y) && eq(a.c1.x, b.c1.x) && eq(a.c1.y, b.c1.y) &&
           eq(a.c2.x, b.c2.x) && eq(a.c2.y, b.c2.y);
}
struct MmParams {
    float gainOmega, muGain, omegaMin, omegaMax;
};
#ifdef QDSP_MM_SPECULATION   // experimental speculate-and-verify variant: exact, measured, does not pay; not built by default
// one symbol: interpolate at L.i from x[i-7 .. i] (`src`, 8 elements), run the detector and the loop update
template <bool CPLX>
__device__ __forceinline__ float2 mm_symbol(MmLoop& L, const float* src, const float* taps, const MmParams& P) {
    const float* t8 = taps + 8 * (int)roundf(__fmul_rn(L.mu, 128.0f));
    float phaseError;
    float2 o;
    if (!CPLX) {
        float outVal = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) outVal = __fadd_rn(outVal, __fmul_rn(src[k], t8[k]));
        o = make_float2(outVal, 0.f);
        phaseError = __fsub_rn(__fmul_rn(mm_step(L.lastOutput), outVal), __fmul_rn(L.lastOutput, mm_step(outVal)));
        L.lastOutput = outVal;
    } else {
        L.p2 = L.p1;
        L.p1 = L.p0;
        L.c2 = L.c1;
        L.c1 = L.c0;
        float re = 0.0f, im = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            re = __fadd_rn(re, __fmul_rn(src[2 * k], t8[k]));
            im = __fadd_rn(im, __fmul_rn(src[2 * k + 1], t8[k]));
        }
        L.p0 = o = make_float2(re, im);
        L.c0 = make_float2(mm_step(re), mm_step(im));
        const float ar = __fsub_rn(L.p0.x, L.p2.x), ai = __fsub_rn(L.p0.y, L.p2.y);
        const float dr = __fsub_rn(L.c0.x, L.c2.x), di = __fsub_rn(L.c0.y, L.c2.y);
        const float ab = __fsub_rn(__fmul_rn(ar, L.c1.x), __fmul_rn(ai, -L.c1.y));
        const float de = __fsub_rn(__fmul_rn(dr, L.p1.x), __fmul_rn(di, -L.p1.y));
        phaseError = __fsub_rn(ab, de);
    }
    if (phaseError > 1.0f) phaseError = 1.0f;
    if (phaseError < -1.0f) phaseError = -1.0f;
    L.dynOmega = __fadd_rn(L.dynOmega, __fmul_rn(P.gainOmega, phaseError));
    if (L.dynOmega > P.omegaMax) L.dynOmega = P.omegaMax;
    else if (L.dynOmega < P.omegaMin) L.dynOmega = P.omegaMin;
    L.mu = __fadd_rn(__fadd_rn(L.mu, L.dynOmega), __fmul_rn(P.muGain, phaseError));
    const float roundedStep = floorf(L.mu);
    L.i += (long long)(int)roundedStep;
    if (L.i < 0) L.i = 0;
    L.mu = __fsub_rn(L.mu, roundedStep);
    return o;
}
// walk from L until the symbol index reaches `end`; symbols at i >= begin are stored (value + source index)
template <bool CPLX>
__device__ int mm_walk(MmLoop& L, MmLoop* entry, long long begin, long long end, const float* __restrict__ in,
                       const float* __restrict__ delay14, const float* taps, const MmParams& P, float* __restrict__ vals,
                       int* __restrict__ idx, int cap) {
    constexpr int ES = CPLX ? 2 : 1;
    int n = 0;
    bool entered = false;
    while (L.i < end) {
        if (!entered && L.i >= begin) {
            if (entry) *entry = L;
            entered = true;
        }
        const long long i = L.i;
        float tmp[16];
        const float* src;
        if (i >= 7) {
            src = in + (i - 7) * ES;
        } else {   // the call's first symbols look back into the carried delay line
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const long long g = i - 7 + k;
                for (int e = 0; e < ES; e++) tmp[ES * k + e] = g >= 0 ? in[g * ES + e] : delay14[2 * (int)(g + 7) + e];
            }
            src = tmp;
        }
        const float2 o = mm_symbol<CPLX>(L, src, taps, P);
        if (entered && n < cap) {
            if (CPLX) reinterpret_cast<float2*>(vals)[n] = o;
            else vals[n] = o.x;
            idx[n] = (int)i;
            n++;
        }
    }
    if (!entered && entry) *entry = L;
    return n;
}
__device__ __forceinline__ MmLoop mm_load_state(const float* st) {
    MmLoop L;
    L.i = (long long)(int)st[15];
    if (L.i < 0) L.i = 0;
    L.mu = st[0];
    L.dynOmega = st[1];
    L.lastOutput = st[2];
    L.p0 = make_float2(st[3], st[4]);
    L.p1 = make_float2(st[5], st[6]);
    L.p2 = make_float2(st[7], st[8]);
    L.c0 = make_float2(st[9], st[10]);
    L.c1 = make_float2(st[11], st[12]);
    L.c2 = make_float2(st[13], st[14]);
    return L;
}

template <bool CPLX>
__global__ void __launch_bounds__(64) mm_spec_kernel(const float* __restrict__ in, long long N, int nchunks, int chunk,
                                                     int warm, const float* __restrict__ taps_g, MmParams P,
                                                     const float* __restrict__ state, MmLoop* __restrict__ entry,
                                                     MmLoop* __restrict__ exitst, int* __restrict__ counts,
                                                     float* __restrict__ vals, int* __restrict__ idx, int cap) {
    constexpr int ES = CPLX ? 2 : 1;
    __shared__ float s_taps[129 * 8];
    for (int k = threadIdx.x; k < 129 * 8; k += blockDim.x) s_taps[k] = taps_g[k];
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const long long begin = (long long)c * chunk, end = begin + chunk < N ? begin + chunk : N;
    MmLoop L;
    if (c == 0) {
        L = mm_load_state(state);
    } else {   // speculative start: the default loop state (clock_recovery.h:232-240), the carried symbol period
        L.i = begin - warm;
        L.mu = 0.5f;
        L.dynOmega = state[1];
        L.lastOutput = 0.0f;
        L.p0 = L.p1 = L.p2 = L.c0 = L.c1 = L.c2 = make_float2(0.f, 0.f);
    }
    MmLoop E;
    counts[c] = mm_walk<CPLX>(L, &E, begin, end, in, state + 16, s_taps, P, vals + (size_t)c * cap * ES, idx + (size_t)c * cap, cap);
    entry[c] = E;
    exitst[c] = L;
}
// single thread: accept / re-walk the chunks in order, exclusive prefix of the counts, final state
template <bool CPLX>
__global__ void mm_verify_kernel(const float* __restrict__ in, long long N, int nchunks, int chunk,
                                 const float* __restrict__ taps_g, MmParams P, float* __restrict__ state,
                                 const MmLoop* __restrict__ entry, MmLoop* __restrict__ exitst, int* __restrict__ counts,
                                 long long* __restrict__ offsets, float* __restrict__ vals, int* __restrict__ idx, int cap,
                                 long long* __restrict__ total_out, int* __restrict__ rewalked) {
    constexpr int ES = CPLX ? 2 : 1;
    MmLoop truth = exitst[0];
    long long off = 0;
    int bad = 0;
    offsets[0] = 0;
    off += counts[0];
    for (int c = 1; c < nchunks; c++) {
        const long long begin = (long long)c * chunk, end = begin + chunk < N ? begin + chunk : N;
        if (mm_same(entry[c], truth)) {
            truth = exitst[c];
        } else {   // the speculation had not merged: this chunk again, from the true state
            bad++;
            MmLoop L = truth;
            counts[c] = mm_walk<CPLX>(L, nullptr, begin, end, in, state + 16, taps_g, P, vals + (size_t)c * cap * ES,
                                      idx + (size_t)c * cap, cap);
            exitst[c] = L;
            truth = L;
        }
        offsets[c] = off;
        off += counts[c];
    }
    offsets[nchunks] = off;
    *total_out = off;
    *rewalked = bad;
    // carried state: clock_recovery.h:208-211
    state[0] = truth.mu;
    state[1] = truth.dynOmega;
    state[2] = truth.lastOutput;
    state[3] = truth.p0.x; state[4] = truth.p0.y; state[5] = truth.p1.x; state[6] = truth.p1.y;
    state[7] = truth.p2.x; state[8] = truth.p2.y;
    state[9] = truth.c0.x; state[10] = truth.c0.y; state[11] = truth.c1.x; state[12] = truth.c1.y;
    state[13] = truth.c2.x; state[14] = truth.c2.y;
    state[15] = (float)(int)(truth.i - N);
    float tail[14];
    for (int k = 0; k < 7; k++) {
        const long long g = N - 7 + k;
        for (int e = 0; e < 2; e++)
            tail[2 * k + e] = g >= 0 ? (e < ES ? in[g * ES + e] : 0.0f) : state[16 + 2 * (int)(g + 7) + e];
    }
    for (int k = 0; k < 14; k++) state[16 + k] = tail[k];
}
// chunk outputs -> final positions; symbols per run() block from the source indices (monotone within and across chunks)
template <bool CPLX>
__global__ void __launch_bounds__(256) mm_compact_kernel(int nchunks, int cap, const int* __restrict__ counts,
                                                         const long long* __restrict__ offsets, const float* __restrict__ vals,
                                                         float* __restrict__ out) {
    constexpr int ES = CPLX ? 2 : 1;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int n = counts[c];
        const long long o = offsets[c];
        for (int k = threadIdx.x; k < n * ES; k += blockDim.x) out[o * ES + k] = vals[(size_t)c * cap * ES + k];
    }
}
__global__ void mm_blockcount_kernel(PartitionDev part, int nchunks, int chunk, int cap, const int* __restrict__ counts,
                                     const long long* __restrict__ offsets, const int* __restrict__ idx,
                                     int* __restrict__ out_counts) {
    // position (in the output stream) of the first symbol whose source index is >= g
    auto pos = [&](long long g) -> long long {
        if (g >= part.total) return offsets[nchunks];
        int c = (int)(g / chunk);
        if (c >= nchunks) return offsets[nchunks];
        int lo = 0, hi = counts[c];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (idx[(size_t)c * cap + mid] < g) lo = mid + 1;
            else hi = mid;
        }
        return offsets[c] + lo;
    };
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < part.nblocks; b += gridDim.x * blockDim.x) {
        const BlkInfo bi = part.get(b);
        out_counts[b] = (int)(pos(bi.in_start + bi.count) - pos(bi.in_start));
    }
}

size_t mm_spec_scratch_bytes(long long count, int chunk, int cap) {
    const long long nchunks = (count + chunk - 1) / chunk + 1;
    return (size_t)nchunks * ((size_t)cap * 12 + 2 * sizeof(MmLoop) + sizeof(int) + sizeof(long long)) + 256;
}
int launch_mm_spec(int cplx, const void* in, const Partition& part, const float* taps_dev, float gainOmega, float muGain,
                   float omegaMin, float omegaMax, float* state, void* out, int* out_counts_dev, long long* total_dev,
                   int* rewalked_dev, int chunk, int warm, int cap, void* scratch, cudaStream_t s) {
    const long long N = part.view.total;
    const int nchunks = (int)((N + chunk - 1) / chunk);
    char* p = (char*)scratch;
    MmLoop* entry = (MmLoop*)p;
    p += sizeof(MmLoop) * (size_t)(nchunks + 1);
    MmLoop* exitst = (MmLoop*)p;
    p += sizeof(MmLoop) * (size_t)(nchunks + 1);
    long long* offsets = (long long*)p;
    p += sizeof(long long) * (size_t)(nchunks + 1);
    int* counts = (int*)p;
    p += sizeof(int) * (size_t)(nchunks + 2) / 2 * 2;
    int* idx = (int*)p;
    p += sizeof(int) * (size_t)nchunks * cap;
    float* vals = (float*)p;
    const MmParams P{gainOmega, muGain, omegaMin, omegaMax};
    const float* x = (const float*)in;
    const int g1 = (nchunks + 63) / 64;
    int gc = nchunks < 1184 ? nchunks : 1184;
    if (cplx) {
        mm_spec_kernel<true><<<g1, 64, 0, s>>>(x, N, nchunks, chunk, warm, taps_dev, P, state, entry, exitst, counts, vals, idx, cap);
        QDSP_LAUNCH_OK();
        mm_verify_kernel<true><<<1, 1, 0, s>>>(x, N, nchunks, chunk, taps_dev, P, state, entry, exitst, counts, offsets, vals, idx,
                                               cap, total_dev, rewalked_dev);
        QDSP_LAUNCH_OK();
        mm_compact_kernel<true><<<gc, 256, 0, s>>>(nchunks, cap, counts, offsets, vals, (float*)out);
    } else {
        mm_spec_kernel<false><<<g1, 64, 0, s>>>(x, N, nchunks, chunk, warm, taps_dev, P, state, entry, exitst, counts, vals, idx, cap);
        QDSP_LAUNCH_OK();
        mm_verify_kernel<false><<<1, 1, 0, s>>>(x, N, nchunks, chunk, taps_dev, P, state, entry, exitst, counts, offsets, vals, idx,
                                                cap, total_dev, rewalked_dev);
        QDSP_LAUNCH_OK();
        mm_compact_kernel<false><<<gc, 256, 0, s>>>(nchunks, cap, counts, offsets, vals, (float*)out);
    }
    QDSP_LAUNCH_OK();
    if (out_counts_dev) {
        mm_blockcount_kernel<<<(part.view.nblocks + 127) / 128, 128, 0, s>>>(part.view, nchunks, chunk, cap, counts, offsets, idx,
                                                                             out_counts_dev);
        QDSP_LAUNCH_OK();
    }
    return 0;
}

#else
size_t mm_spec_scratch_bytes(long long, int, int) { return 0; }
int launch_mm_spec(int, const void*, const Partition&, const float*, float, float, float, float, float*, void*, int*, long long*,
                   int*, int, int, int, void*, cudaStream_t) {
    set_last_error("mm: the speculate-and-verify variant is not compiled in (build with -DQDSP_MM_SPECULATION)");
    return -1;
}
#endif

int launch_mm(int cplx, const void* in, const Partition& part, const float* taps_dev, float omega, float gainOmega,
              float muGain, float omegaMin, float omegaMax, float* state, void* out, int* out_counts_dev,
              long long* total_dev, cudaStream_t s) {
    if (cplx)
        mm_kernel<true><<<1, 32, 0, s>>>((const float*)in, part.view, taps_dev, omega, gainOmega, muGain, omegaMin, omegaMax,
                                         state, (float*)out, out_counts_dev, total_dev);
    else
        mm_kernel<false><<<1, 32, 0, s>>>((const float*)in, part.view, taps_dev, omega, gainOmega, muGain, omegaMin, omegaMax,
                                          state, (float*)out, out_counts_dev, total_dev);
    QDSP_LAUNCH_OK();
    return 0;
}

}  // namespace qdsp
