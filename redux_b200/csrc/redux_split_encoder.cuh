// redux_split_encoder.cuh -- intra-stream parallel ENCODER for a few long streams.
//
// A stream is bit-serial only in its coder half: low/high of symbol t+1 depend on symbol t
// (src/codec.rs:58-60).  The model half does not depend on the coder at all: the range a
// symbol is coded with is a function of the symbol PREFIX,
//     cum_lo(t) = s_t + #{ j < min(t, T) : s_j <  s_t },      T = freq_max - symbol_count (the freeze,
//     cum_hi(t) = cum_lo(t) + 1 + #{ j < min(t, T) : s_j == s_t }      adaptive_tree.rs:84)
// so BASELINE configs 1, 2 and 5 (one file, 29 files, a dozen 1 MiB blocks -- far too few streams to fill
// the machine one stream per lane) are encoded in two phases:
//   A. model, massively parallel: the stream is cut into chunks of kSplitChunk symbols;
//        split_hist_kernel   histogram of every chunk (positions < T only)
//        split_scan_kernel   exclusive prefix over the chunks of a stream, per symbol
//        split_model_kernel  one warp per chunk: the cumulative array of AdaptiveLinearModel
//                            (adaptive_linear.rs:21-70) in registers, seeded with the prefix, walks the
//                            chunk and writes (cum_lo, cum_hi) for every position
//   B. coder, one warp per stream (split_coder_kernel): only the range update, the closed-form
//      renormalisation and the bit packing remain on the serial chain -- no table, no search, no update;
//      the ranges and the per-position reciprocals arrive 32 at a time through coalesced loads.
// The bytes are those of the lane kernels and of the oracle (tests/test_gpu_parity.py runs all mappings).
// Decoding has no such split: the symbol itself comes out of the coder state.
#pragma once
#include "redux_common.cuh"
#include "redux_lane_al.cuh"
#include "redux_warp_codec.cuh"

namespace rdx {

constexpr uint32_t kSplitChunk = 2048;          // symbols per model chunk
constexpr int kSplitModelWarps = 4;             // warps (chunks) per CTA of the model kernel

struct SplitJob {
    const uint8_t *in; const uint64_t *in_off; uint64_t n_blocks;
    uint32_t chunks_per_stream;                 // ceil(max_block_len / kSplitChunk), same for every stream
    uint32_t *hist;                             // [n_blocks][chunks_per_stream][256]
    uint2 *pairs;                               // [n_blocks][pair_stride] (cum_lo, cum_hi) per position
    uint64_t pair_stride;
    uint32_t tcap;                              // T
};

// ---- A1: per-chunk histograms of the positions that still update the model
__global__ void __launch_bounds__(256)
split_hist_kernel(const SplitJob job)
{
    __shared__ uint32_t h[256];
    const uint64_t blk = blockIdx.x / job.chunks_per_stream;
    const uint32_t chunk = blockIdx.x % job.chunks_per_stream;
    const uint64_t off = job.in_off[blk];
    const uint64_t len = job.in_off[blk + 1] - off;
    const uint64_t lim = len < job.tcap ? len : job.tcap;
    const uint64_t beg = (uint64_t)chunk * kSplitChunk;
    const uint64_t end = beg + kSplitChunk < lim ? beg + kSplitChunk : lim;
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = beg + threadIdx.x; i < end; i += 256) atomicAdd(&h[job.in[off + i]], 1u);
    __syncthreads();
    job.hist[((size_t)blk * job.chunks_per_stream + chunk) * 256 + threadIdx.x] = h[threadIdx.x];
}

// ---- A2: hist[chunk][s] <- number of s in the chunks before it (per stream); thread = symbol
__global__ void __launch_bounds__(256)
split_scan_kernel(const SplitJob job)
{
    uint32_t *h = job.hist + (size_t)blockIdx.x * job.chunks_per_stream * 256 + threadIdx.x;
    uint32_t run = 0;
    for (uint32_t c = 0; c < job.chunks_per_stream; ++c) {
        const uint32_t v = h[(size_t)c * 256];
        h[(size_t)c * 256] = run;
        run += v;
    }
}

// ---- A3: one warp per chunk walks it with the cumulative array in registers
__global__ void __launch_bounds__(kSplitModelWarps * 32)
split_model_kernel(const SplitJob job)
{
    __shared__ uint32_t pre[kSplitModelWarps][256];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t gchunk = (uint64_t)blockIdx.x * kSplitModelWarps + w;
    const uint64_t blk = gchunk / job.chunks_per_stream;
    if (blk >= job.n_blocks) return;                              // warp-uniform
    const uint32_t chunk = (uint32_t)(gchunk % job.chunks_per_stream);
    const uint64_t off = job.in_off[blk];
    const uint32_t len = (uint32_t)(job.in_off[blk + 1] - off);
    const uint32_t beg = chunk * kSplitChunk;
    if (beg >= len) return;
    const uint32_t end = beg + kSplitChunk < len ? beg + kSplitChunk : len;

    // inc[i] = number of earlier symbols < i: exclusive prefix sum over the symbol axis of the chunk's
    // prefix histogram.  Lane l sums entries 8l..8l+7, a warp scan joins the lanes.
    const uint32_t *hp = job.hist + gchunk * 256;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = hp[lane * 8 + j]; sum += v[j]; }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (uint32_t)d) incl += y;
    }
    uint32_t run = incl - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) { pre[w][lane * 8 + j] = run; run += v[j]; }
    __syncwarp();
    WarpTable tab;
    tab.init(lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) tab.r[j] = pre[w][lane + 32 * j];

    WarpByteSource src;
    src.init(job.in + off + beg, end - beg, lane);
    uint2 *out = job.pairs + blk * job.pair_stride;
    uint2 mine = make_uint2(0, 0);
    for (uint32_t t = beg; t < end; ++t) {
        const uint32_t sym = src.next();
        const bool adapt = t < job.tcap;
        uint32_t cl, ch;
        tab.query(sym, adapt ? t : job.tcap, cl, ch);             // node 256 = number of updates so far
        tab.update(sym, adapt);
        if (((t - beg) & 31) == lane) mine = make_uint2(cl, ch);
        if (((t - beg) & 31) == 31) out[t - 31 + lane] = mine;    // 32 positions, one coalesced store
    }
    const uint32_t tail = (end - beg) & 31;
    if (lane < tail) out[end - tail + lane] = mine;
}

// ---- B: the serial coder chain, one warp per stream (state warp-uniform, lane 0 stores)
struct WarpBitSink2 : BitSink2 {
    bool en;
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        acc = (acc << n) | v;
        nb += n;
        const bool full = nb >= 32;                                // predicated, no branch on the chain
        const uint32_t word = __byte_perm((uint32_t)(acc >> (nb & 31)), 0, 0x0123);
        if (full && en) w0[wi] = word;
        wi += full ? 1u : 0u;
        nb &= 31u;
    }
    __device__ __forceinline__ uint32_t put_code(uint32_t bits, uint32_t n1, uint32_t pend, uint32_t k) {
        const bool emit = n1 != 0;
        const uint32_t n = emit ? n1 + pend : 0u;
        if (n > 32) {
            const uint32_t b = (bits >> (n1 - 1)) & 1u;
            put(b, 1);
            while (pend > 0) {
                const uint32_t m = pend < 32 ? pend : 32;
                put(b ? 0u : (0xFFFFFFFFu >> (32 - m)), m);
                pend -= m;
            }
            if (n1 > 1) put(bits & (0xFFFFFFFFu >> (33 - n1)), n1 - 1);
        } else {
            put(bits + __funnelshift_lc(0u, 1u, n - 1u) - __funnelshift_lc(0u, 1u, n1 - 1u), n);
        }
        return (emit ? 0u : pend) + k;
    }
    __device__ __forceinline__ uint32_t finish() {
        const uint32_t bytes = wi * 4 + (nb + 7) / 8;
        if (nb && en) w0[wi] = __byte_perm((uint32_t)(acc << (32 - nb)), 0, 0x0123);
        return bytes;
    }
};

// One step of the serial chain.  Unlike the lane kernels the chain carries the RANGE (minus one) itself:
// shifts double low and high alike, so range' = (quotient_hi - quotient_lo) << shifts is known two
// operations after the shift count, without waiting for the new low/high registers.
template <int CLS, bool C32>
__device__ __forceinline__ uint32_t split_step(uint32_t &L, uint32_t &rm1, uint32_t &pend, WarpBitSink2 &sink,
                                               uint32_t cl, uint32_t ch, uint32_t count,
                                               const typename Cls<CLS>::M &g, uint32_t one)
{
    using C = Cls<CLS>;
    using P = typename C::P;
    const P nh = C::mulr(ch, rm1), nl = C::mulr(cl, rm1);
    const uint32_t qh = (uint32_t)C::divc(nh, g, count), ql = (uint32_t)C::divc(nl, g, count);
    const uint32_t nh2 = ~(qh * one + (L - 1u));
    const uint32_t l2 = ql * one + L;
    const uint32_t n1 = common_prefix<C32>(~(l2 ^ nh2));
    const uint32_t k = clz_nz(~shl_c((l2 & nh2) << 1, n1));
    const uint32_t n = n1 + k;
    rm1 = shl_c(qh - ql, n) - 1u;                                  // (high' - low' + 1) << n, minus one
    pend = sink.put_code(top_bits(l2, n1), n1, pend, k);
    L = shl_c(l2, n) & 0x7FFFFFFFu;
    return n;
}

template <int CLS, bool C32>
__global__ void __launch_bounds__(32)
split_coder_kernel(const LaneEncJob job, const SplitJob sj)
{
    using C = Cls<CLS>;
    using M = typename C::M;
    const uint32_t lane = threadIdx.x;
    const uint64_t blk = blockIdx.x;
    const uint64_t off = job.in_off[blk];
    const uint32_t len = (uint32_t)(job.in_off[blk + 1] - off);
    const uint32_t c = job.c, one = job.one, tcap = job.tcap;
    const M *magic = reinterpret_cast<const M *>(job.magic);
    const uint2 *pairs = sj.pairs + blk * sj.pair_stride;
    WarpBitSink2 sink;
    sink.init(job.slots + blk * job.slot_stride);
    sink.en = lane == 0;
    const uint32_t maxv = c == 32 ? 0xFFFFFFFFu : ((1u << c) - 1u);
    uint32_t L = 0, rm1 = maxv, pend = 0;                          // low = 0, range - 1 = code_max (src/codec.rs:30-31)

    // 32 positions per round: lane l fetches the range and the reciprocal of position base + l (one round
    // ahead); the serial chain takes them from the lanes by shuffle, one STEP ahead, so that neither the
    // loads nor the shuffles sit on the chain
    auto fetch_pair = [&](uint32_t base) {
        const uint32_t t = base + lane;
        return t < len ? pairs[t] : make_uint2(0, 1);
    };
    auto fetch_magic = [&](uint32_t base) {
        const uint32_t t = base + lane;
        return C::ldm(magic + (t < tcap ? t : tcap));              // count_t = 257 + min(t, T)
    };
    uint2 pn = fetch_pair(0);
    M gn = fetch_magic(0);
    for (uint32_t base = 0; base < len; base += 32) {
        const uint2 p = pn;
        const M g = gn;
        pn = fetch_pair(base + 32);
        gn = fetch_magic(base + 32);
        const uint32_t m = len - base < 32 ? len - base : 32;
        uint32_t cl_n = __shfl_sync(kFullMask, p.x, 0), ch_n = __shfl_sync(kFullMask, p.y, 0);
        M g_n = g;
        g_n.m = __shfl_sync(kFullMask, g.m, 0);
        g_n.sh = __shfl_sync(kFullMask, g.sh, 0);
#pragma unroll 2
        for (uint32_t i = 0; i < m; ++i) {
            const uint32_t cl = cl_n, ch = ch_n;
            const M gi = g_n;
            const int nx = (int)((i + 1) & 31);
            cl_n = __shfl_sync(kFullMask, p.x, nx); ch_n = __shfl_sync(kFullMask, p.y, nx);
            g_n.m = __shfl_sync(kFullMask, g.m, nx);
            g_n.sh = __shfl_sync(kFullMask, g.sh, nx);
            const uint32_t t = base + i;
            split_step<CLS, C32>(L, rm1, pend, sink, cl, ch, kNsym + (t < tcap ? t : tcap), gi, one);
        }
    }
    const uint32_t tt = len < tcap ? len : tcap;
    const uint32_t countf = kNsym + tt;
    const M gf = C::ldm(magic + tt);
    const uint32_t shifts = split_step<CLS, C32>(L, rm1, pend, sink, countf - 1, countf, countf, gf, one);
    const uint32_t extra = c - shifts;                             // src/codec.rs:91-99
    sink.put_code(top_bits(L, extra), extra, pend, 0);
    const uint32_t bytes = sink.finish();
    if (lane == 0) { job.sizes[blk] = bytes; job.status[blk] = 0; }
}

}  // namespace rdx
