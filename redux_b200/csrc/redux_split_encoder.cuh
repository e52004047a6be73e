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

// ---- B: the serial coder chain, one warp per stream.
// The chain itself is only the range update and the closed-form renormalisation (every lane computes it;
// the state is warp-uniform).  What a step emits -- the n1 settled bits and its E3 count k -- is LATCHED
// by the lane whose index equals the step's position in the round, and the 32 steps of a round are packed
// by the whole warp at once (RoundPacker): pending runs by a segmented sum, bit offsets by a prefix sum,
// the bits ORed into a shared staging window and stored as whole big-endian words, coalesced.  No
// branch, store or packer state sits between two steps of the chain.
struct RoundPacker {
    uint32_t *stage;        // shared: [kStageWords], word 0 starts with the `cb` carried bits
    uint32_t *w0;           // output slot (16-byte aligned)
    uint32_t wi, cb, pend;  // words stored so far, carried bits (< 32), pending E3 run -- all warp-uniform
    uint32_t lane;
    static constexpr int kStageWords = 40;      // 31 carried bits + 32 steps x 32 bits + slack

    __device__ __forceinline__ void init(uint32_t *smem, uint8_t *slot, uint32_t l) {
        stage = smem; w0 = (uint32_t *)slot; wi = 0; cb = 0; pend = 0; lane = l;
        for (int i = (int)l; i < kStageWords; i += 32) stage[i] = 0;
        __syncwarp();
    }
    // stores the complete words of the window, keeps the partial one as the new carry
    __device__ __forceinline__ void drain(uint32_t total_bits) {
        __syncwarp();
        const uint32_t full = total_bits >> 5;
        const uint32_t carry = stage[full];
        for (uint32_t i = lane; i < full; i += 32) w0[wi + i] = __byte_perm(stage[i], 0, 0x0123);
        __syncwarp();
        for (int i = (int)lane; i < kStageWords; i += 32) stage[i] = 0;
        __syncwarp();
        if (lane == 0) stage[0] = carry;
        __syncwarp();
        wi += full; cb = total_bits & 31;
    }
    // warp-uniform append of the low `len` (<= 32) bits of v at the end of the window
    __device__ __forceinline__ void append_uniform(uint32_t v, uint32_t len) {
        if (len == 0) return;
        const uint64_t x = (uint64_t)v << (64 - len - cb);
        if (lane == 0) { stage[0] |= (uint32_t)(x >> 32); stage[1] |= (uint32_t)x; }
        drain(cb + len);
    }
    // put_bit semantics (src/codec.rs:39-46) for one step, warp-uniform, any run length
    __device__ __forceinline__ void step_uniform(uint32_t bits, uint32_t n1, uint32_t k) {
        if (n1) {
            const uint32_t b = (bits >> (n1 - 1)) & 1u;
            append_uniform(b, 1);
            for (uint32_t left = pend; left > 0;) {
                const uint32_t m = left < 32 ? left : 32;
                append_uniform(b ? 0u : (0xFFFFFFFFu >> (32 - m)), m);
                left -= m;
            }
            if (n1 > 1) append_uniform(bits & (0xFFFFFFFFu >> (33 - n1)), n1 - 1);
            pend = k;
        } else {
            pend += k;
        }
    }
    // One round: lane i < m holds step i's (bits, n1, k); lanes >= m must hold n1 = k = 0.
    __device__ __forceinline__ void round(uint32_t bits, uint32_t n1, uint32_t k) {
        // pending run in front of every step: k summed since the last emitting step (segmented sum)
        uint32_t S = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFullMask, S, d);
            if (lane >= (uint32_t)d) S += y;
        }
        const uint32_t Sprev = S - k;                              // sum of k over the steps before mine
        const uint32_t emask = __ballot_sync(kFullMask, n1 != 0);
        const uint32_t below = emask & ((1u << lane) - 1u);
        const uint32_t le = below ? 31u - (uint32_t)__clz((int)below) : 0u;
        const uint32_t Sle = __shfl_sync(kFullMask, Sprev, (int)le);
        const uint32_t pb = below ? Sprev - Sle : pend + Sprev;
        const uint32_t n = n1 ? n1 + pb : 0u;
        // pending run left behind by the round
        const uint32_t S31 = __shfl_sync(kFullMask, S, 31);
        const uint32_t last = emask ? 31u - (uint32_t)__clz((int)emask) : 0u;
        const uint32_t Slast = __shfl_sync(kFullMask, Sprev, (int)last);
        const uint32_t pend_out = emask ? S31 - Slast : pend + S31;
        if (__any_sync(kFullMask, n > 32)) {
            // a long E3 run (rare): step by step, warp-uniform
            for (int i = 0; i < 32; ++i) {
                const uint32_t bi = __shfl_sync(kFullMask, bits, i), ni = __shfl_sync(kFullMask, n1, i);
                const uint32_t ki = __shfl_sync(kFullMask, k, i);
                step_uniform(bi, ni, ki);
            }
            return;
        }
        // bit offsets: exclusive prefix sum of n
        uint32_t O = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFullMask, O, d);
            if (lane >= (uint32_t)d) O += y;
        }
        const uint32_t total = __shfl_sync(kFullMask, O, 31);
        O -= n;
        if (n) {
            // [b][pb x !b][rest] as one number: bits + 2^(n-1) - 2^(n1-1) (see BitSink2::put_code)
            const uint32_t v = bits + (1u << (n - 1)) - (1u << (n1 - 1));
            const uint32_t at = cb + O;
            const uint64_t x = (uint64_t)v << (64 - n - (at & 31));
            atomicOr(&stage[at >> 5], (uint32_t)(x >> 32));
            if ((uint32_t)x) atomicOr(&stage[(at >> 5) + 1], (uint32_t)x);
        }
        pend = pend_out;
        drain(cb + total);
    }
    // flush_bits (src/bitio/mod.rs:183-198): the last partial word, zero padded; returns the byte count
    __device__ __forceinline__ uint32_t finish() {
        __syncwarp();
        if (cb && lane == 0) w0[wi] = __byte_perm(stage[0], 0, 0x0123);
        return wi * 4 + (cb + 7) / 8;
    }
};

// One step of the serial chain.  Unlike the lane kernels the chain carries the RANGE (minus one) itself:
// shifts double low and high alike, so range' = (quotient_hi - quotient_lo) << shifts is known two
// operations after the shift count, without waiting for the new low/high registers.
template <int CLS, bool C32>
__device__ __forceinline__ void split_step(uint32_t &L, uint32_t &rm1, uint32_t cl, uint32_t ch, uint32_t count,
                                           const typename Cls<CLS>::M &g, uint32_t one,
                                           uint32_t &bits, uint32_t &n1, uint32_t &k)
{
    using C = Cls<CLS>;
    using P = typename C::P;
    const P nh = C::mulr(ch, rm1), nl = C::mulr(cl, rm1);
    const uint32_t qh = (uint32_t)C::divc(nh, g, count), ql = (uint32_t)C::divc(nl, g, count);
    const uint32_t nh2 = ~(qh * one + (L - 1u));
    const uint32_t l2 = ql * one + L;
    uint32_t n;
    renorm_counts<C32>(l2, nh2, n1, n);
    k = n - n1;
    rm1 = shl_c(qh - ql, n) - 1u;                                  // (high' - low' + 1) << n, minus one
    bits = top_bits(l2, n1);
    L = shl_c(l2, n) & 0x7FFFFFFFu;
}

template <int CLS, bool C32>
__global__ void __launch_bounds__(32)
split_coder_kernel(const LaneEncJob job, const SplitJob sj)
{
    using C = Cls<CLS>;
    using M = typename C::M;
    __shared__ uint32_t stage[RoundPacker::kStageWords];
    const uint32_t lane = threadIdx.x;
    const uint64_t blk = blockIdx.x;
    const uint64_t off = job.in_off[blk];
    const uint32_t len = (uint32_t)(job.in_off[blk + 1] - off);
    const uint32_t c = job.c, one = job.one, tcap = job.tcap;
    const M *magic = reinterpret_cast<const M *>(job.magic);
    const uint2 *pairs = sj.pairs + blk * sj.pair_stride;
    RoundPacker pk;
    pk.init(stage, job.slots + blk * job.slot_stride, lane);
    const uint32_t maxv = c == 32 ? 0xFFFFFFFFu : ((1u << c) - 1u);
    uint32_t L = 0, rm1 = maxv;                                    // low = 0, range - 1 = code_max (src/codec.rs:30-31)

    // 32 positions per round: lane l fetches the range and the reciprocal of position base + l (one round
    // ahead); the chain takes them from the lanes by shuffle, one STEP ahead
    auto fetch_pair = [&](uint32_t base) {
        const uint32_t t = base + lane;
        return t < len ? pairs[t] : make_uint2(0, 1);
    };
    // count_t = 257 + min(t, T); positions past the stream's own end (the last round is fetched whole and a
    // full round ahead) re-read the entry of the EOF step, so no index leaves the table the plan sized
    const uint32_t tlast = len < tcap ? len : tcap;
    auto fetch_magic = [&](uint32_t base) {
        const uint32_t t = base + lane;
        return C::ldm(magic + (t < tlast ? t : tlast));
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
        uint32_t my_bits = 0, my_n1 = 0, my_k = 0;
#pragma unroll 4
        for (uint32_t i = 0; i < m; ++i) {
            const uint32_t cl = cl_n, ch = ch_n;
            const M gi = g_n;
            const int nx = (int)((i + 1) & 31);
            cl_n = __shfl_sync(kFullMask, p.x, nx); ch_n = __shfl_sync(kFullMask, p.y, nx);
            g_n.m = __shfl_sync(kFullMask, g.m, nx);
            g_n.sh = __shfl_sync(kFullMask, g.sh, nx);
            const uint32_t t = base + i;
            uint32_t bits, n1, k;
            split_step<CLS, C32>(L, rm1, cl, ch, kNsym + (t < tcap ? t : tcap), gi, one, bits, n1, k);
            const bool mine = lane == i;
            my_bits = mine ? bits : my_bits; my_n1 = mine ? n1 : my_n1; my_k = mine ? k : my_k;
        }
        pk.round(my_bits, my_n1, my_k);
    }
    // EOF symbol: cum(256) = total - 1; then the tail of src/codec.rs:91-99: the remaining `extra` MSBs of
    // low, the first of them carrying the pending run
    const uint32_t tt = len < tcap ? len : tcap;
    const uint32_t countf = kNsym + tt;
    const M gf = C::ldm(magic + tt);
    uint32_t bits, n1, k;
    split_step<CLS, C32>(L, rm1, countf - 1, countf, countf, gf, one, bits, n1, k);
    pk.step_uniform(bits, n1, k);
    const uint32_t extra = c - (n1 + k);
    pk.step_uniform(top_bits(L, extra), extra, 0);
    const uint32_t bytes = pk.finish();
    if (lane == 0) { job.sizes[blk] = bytes; job.status[blk] = 0; }
}

}  // namespace rdx
