// redux_lane_codec.cuh -- throughput mapping: ONE STREAM PER LANE, 32 independent streams per warp.
//
// What the reference does per stream on one CPU thread (src/codec.rs:55-176 over
// src/model/adaptive_tree.rs) runs here in every lane of every warp:
//   * the stream's Fenwick table lives in shared memory, LANE-INTERLEAVED: node i of lane l is word
//     tab[i*32 + l], so any 32 lanes touching any 32 nodes hit 32 different banks -- the
//     data-dependent walks are bank-conflict free by construction;
//   * the table stores INCREMENTS only.  The reference initialises tree[i] = lowbit(i)
//     (adaptive_tree.rs:43-45) and a Fenwick prefix path decomposes i into its set bits, so
//     cum(i) = i + sum(increments on the path).  Node 256 (every data symbol's update ends there,
//     adaptive_tree.rs:86-89) equals the number of updates so far and node 257 (EOF) is constant 1:
//     neither is stored.  256 x u16 per stream = 512 B -> 14 warps (448 streams) per SM;
//   * count_t = min(257 + t, FMAX) depends on the position only, so the two divisions
//     by count (src/codec.rs:59-60) are multiplications by a per-position magic shared by all streams;
//   * renormalisation (src/codec.rs:62-89 / :140-158) is the closed form of redux_common.cuh;
//   * bits are packed into a 64-bit register and leave as whole big-endian 32-bit words
//     (MSB-first bytes, src/bitio/mod.rs:148-198) into the stream's private output slot.
// No __syncthreads, no warp collectives: lanes are fully independent and may be ragged.
#pragma once
#include "redux_common.cuh"

namespace rdx {

constexpr int kLaneWarpsPerCta = 7;                 // 7 warps x 16 KiB tables; 2 CTAs per SM
constexpr int kLaneThreads = kLaneWarpsPerCta * 32;
constexpr int kTabNodes = 256;                      // nodes 0..255 (node 0 is never touched)

struct LaneEncJob {
    const uint8_t *in;          // raw bytes
    const uint64_t *in_off;     // [n_blocks+1]
    uint64_t n_blocks;
    uint8_t *slots;             // n_blocks * slot_stride bytes, 16-byte aligned
    uint64_t slot_stride;
    uint32_t *sizes;            // [n_blocks] compressed bytes
    int32_t *status;            // [n_blocks]
    const void *magic;          // Magic32/Magic64 [magic_len], entry tt <-> count 257+tt
    uint32_t f, c;
    uint32_t tcap;              // FMAX - NSYM: number of model updates before the freeze
};

struct LaneDecJob {
    const uint8_t *comp;
    const uint64_t *comp_off;   // [n_blocks+1]
    uint64_t n_blocks;
    uint8_t *raw;
    const uint64_t *raw_off;    // [n_blocks+1] slot offsets (capacities)
    uint64_t *raw_len;          // [n_blocks]
    uint64_t *consumed;         // [n_blocks]
    int32_t *status;
    const void *magic;
    uint32_t f, c;
    uint32_t tcap;
};

// ------------------------------------------------------------------ arithmetic class traits
template <int CLS> struct Cls;
template <> struct Cls<kNarrow> {
    using S = uint32_t; using P = uint32_t; using M = Magic32;
    static __device__ __forceinline__ P mulr(uint32_t v, S rm1) { return v * rm1 + v; }   // v * range
    static __device__ __forceinline__ P divc(P n, const M &g, uint32_t) { return div_magic32(n, g); }
    static __device__ __forceinline__ M ldm(const M *p) {
        uint2 v = __ldg(reinterpret_cast<const uint2 *>(p)); M g; g.m = v.x; g.sh = v.y; return g;
    }
};
template <> struct Cls<kWide> {
    using S = uint32_t; using P = uint64_t; using M = Magic64;
    static __device__ __forceinline__ P mulr(uint32_t v, S rm1) { return (uint64_t)v * rm1 + v; }
    static __device__ __forceinline__ P divc(P n, const M &g, uint32_t) { return div_magic64(n, g); }
    static __device__ __forceinline__ M ldm(const M *p) {
        uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        M g; g.m = ((uint64_t)v.y << 32) | v.x; g.sh = v.z; g.pad = 0; return g;
    }
};
template <> struct Cls<kHuge> {
    using S = uint64_t; using P = uint64_t; using M = Magic64;
    static __device__ __forceinline__ P mulr(uint32_t v, S rm1) { return (uint64_t)v * rm1 + v; }
    static __device__ __forceinline__ P divc(P n, const M &, uint32_t count) { return n / count; }
    static __device__ __forceinline__ M ldm(const M *) { M g; g.m = 0; g.sh = 0; g.pad = 0; return g; }
};

// ------------------------------------------------------------------ Fenwick increments in smem
template <typename TW>
struct LaneTable {
    TW *t;   // already offset by lane: node i at t[i*32]
    __device__ __forceinline__ uint32_t ld(uint32_t node) const { return t[node * 32]; }

    __device__ __forceinline__ void clear() {
#pragma unroll 8
        for (int i = 0; i < kTabNodes; ++i) t[i * 32] = 0;
    }

    // (cum(s), cum(s+1)) for a data symbol s in 0..255: the paired walk of adaptive_tree.rs:63-80
    // flattened.  Bits of s below its lowest zero bit belong to cum(s) only, bits above to both,
    // node s+1 to cum(s+1) only.  `updates` = increments of the unstored node 256.
    __device__ __forceinline__ void query(uint32_t s, uint32_t updates, uint32_t &cl, uint32_t &ch) const {
        const uint32_t lowmask = s & ~(s + 1);
        uint32_t only_l = 0, both = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t bit = 1u << b;
            if (s & bit) {
                uint32_t v = ld(s & (0xFFu << b));
                if (lowmask & bit) only_l += v; else both += v;
            }
        }
        const uint32_t top = (s == 255u) ? updates : ld(s + 1);
        cl = s + only_l + both;
        ch = s + 1 + top + both;
    }

    // update(s+1) of adaptive_tree.rs:83-92 minus the two unstored nodes.
    __device__ __forceinline__ void update(uint32_t s) {
        uint32_t i = s + 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i < 256u) {
                t[i * 32] = (TW)(t[i * 32] + 1);
                i += i & (0u - i);
            }
        }
    }
};

// ------------------------------------------------------------------ bit packer (encoder output)
struct BitSink {
    uint64_t acc;      // newest bit at bit 0
    uint32_t nb;       // valid bits in acc, < 32 between calls
    uint32_t *w;       // next output word (slot is 16-byte aligned)
    uint32_t nwords;

    __device__ __forceinline__ void init(uint8_t *slot) { acc = 0; nb = 0; w = (uint32_t *)slot; nwords = 0; }
    // append the low n (<= 32) bits of v (v < 2^n)
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        acc = (acc << n) | v;
        nb += n;
        if (nb >= 32) {
            uint32_t word = (uint32_t)(acc >> (nb - 32));
            w[nwords++] = __byte_perm(word, 0, 0x0123);   // big-endian: first bit -> MSB of first byte
            nb -= 32;
        }
    }
    // `bits` = n >= 1 code bits (MSB first); the first is followed by `pend` copies of its inverse
    // (put_bit, src/codec.rs:39-46).
    __device__ __forceinline__ void put_with_pending(uint32_t bits, uint32_t n, uint32_t pend) {
        const uint32_t b = (bits >> (n - 1)) & 1u;
        const uint32_t rest = bits & ((1u << (n - 1)) - 1u);
        if (n + pend <= 32) {
            // [b][pend x !b][rest]
            uint32_t run = b ? 0u : ((1u << pend) - 1u);               // pend <= 31 here
            uint32_t v = ((((b << pend) | run)) << (n - 1)) | rest;
            put(v, n + pend);
        } else {
            put(b, 1);
            while (pend > 0) {
                uint32_t m = pend < 32 ? pend : 32;
                put(b ? 0u : (m == 32 ? 0xFFFFFFFFu : ((1u << m) - 1u)), m);
                pend -= m;
            }
            if (n > 1) put(rest, n - 1);
        }
    }
    // flush_bits (src/bitio/mod.rs:183-198): left-align, zero-pad. Returns the byte count.
    __device__ __forceinline__ uint32_t finish() {
        uint32_t bytes = nwords * 4 + (nb + 7) / 8;
        if (nb) {
            uint32_t word = (uint32_t)(acc << (32 - nb));
            w[nwords] = __byte_perm(word, 0, 0x0123);
        }
        return bytes;
    }
};

// ------------------------------------------------------------------ encoder
template <typename TW, int CLS>
__global__ void __launch_bounds__(kLaneThreads, 2)
encode_lane_kernel(const LaneEncJob job)
{
    using C = Cls<CLS>;
    using S = typename C::S;
    using P = typename C::P;
    using M = typename C::M;
    extern __shared__ uint4 smem_u4[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t blk = (uint64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    if (blk >= job.n_blocks) return;

    LaneTable<TW> tab;
    tab.t = reinterpret_cast<TW *>(smem_u4) + (size_t)warp * kTabNodes * 32 + lane;
    tab.clear();

    const uint64_t off = job.in_off[blk];
    const uint32_t len = (uint32_t)(job.in_off[blk + 1] - off);
    const uintptr_t addr = (uintptr_t)(job.in + off);
    const uint4 *p16 = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    const uint32_t skip = (uint32_t)(addr & 15);

    const uint32_t c = job.c;
    const S maxv = (S)((c == sizeof(S) * 8) ? ~(S)0 : ((((S)1) << c) - 1));
    const M *magic = reinterpret_cast<const M *>(job.magic);
    const uint32_t tcap = job.tcap;

    S low = 0, high = maxv;
    uint32_t pend = 0;
    BitSink sink;
    sink.init(job.slots + blk * job.slot_stride);

    uint4 cur = make_uint4(0, 0, 0, 0);
    if (len) cur = __ldg(p16);

    for (uint32_t t = 0; t <= len; ++t) {
        const bool is_eof = (t == len);
        const uint32_t tt = t < tcap ? t : tcap;
        const uint32_t count = kNsym + tt;
        const M g = C::ldm(magic + tt);
        uint32_t cl, ch;
        if (!is_eof) {
            const uint32_t gpos = skip + t;
            if ((gpos & 15) == 0 && t) cur = __ldg(p16 + (gpos >> 4));
            const uint32_t ws = (gpos >> 2) & 3;
            const uint32_t wv = ws == 0 ? cur.x : ws == 1 ? cur.y : ws == 2 ? cur.z : cur.w;
            const uint32_t sym = (wv >> ((gpos & 3) * 8)) & 0xFFu;
            tab.query(sym, tt, cl, ch);
            if (t < tcap) tab.update(sym);
        } else {
            cl = count - 1;      // cum(256) = total - freq(EOF), EOF's frequency never grows
            ch = count;
        }
        // src/codec.rs:58-60
        const S rm1 = high - low;
        const P nh = C::mulr(ch, rm1), nl = C::mulr(cl, rm1);
        high = low + (S)C::divc(nh, g, count) - 1;
        low = low + (S)C::divc(nl, g, count);
        // src/codec.rs:62-89 in closed form
        const Renorm<S> r = renorm<S>(low, high, c);
        if (r.n1) {
            if (sizeof(S) == 4 || r.n1 <= 32) {
                sink.put_with_pending((uint32_t)(low >> (c - r.n1)), r.n1, pend);
            } else {   // c > 32 only: more than 32 common bits
                const uint32_t hi_n = r.n1 - 32;
                sink.put_with_pending((uint32_t)((uint64_t)low >> (c - hi_n)), hi_n, pend);
                sink.put((uint32_t)((uint64_t)low >> (c - r.n1)), 32);
            }
            pend = r.k;
        } else {
            pend += r.k;
        }
        low = r.low; high = r.high;
        if (is_eof) {
            // src/codec.rs:91-99: the remaining `extra` MSBs of low, then flush
            const uint32_t extra = c - (r.n1 + r.k);
            if (extra) {
                if (sizeof(S) == 4 || extra <= 32) {
                    sink.put_with_pending((uint32_t)(low >> (c - extra)), extra, pend);
                } else {
                    const uint32_t hi_n = extra - 32;
                    sink.put_with_pending((uint32_t)((uint64_t)low >> (c - hi_n)), hi_n, pend);
                    sink.put((uint32_t)((uint64_t)low >> (c - extra)), 32);
                }
            }
        }
    }
    job.sizes[blk] = sink.finish();
    job.status[blk] = 0;
}

// ------------------------------------------------------------------ bit source (decoder input)
struct BitSource {
    uint64_t bb;            // bit buffer, valid bits are the low `bn`
    uint32_t bn;
    const uint32_t *w;      // aligned words
    uint32_t widx, nwords;
    uint64_t used, total;   // bits

    __device__ __forceinline__ void init(const uint8_t *src, uint64_t len) {
        const uintptr_t a = (uintptr_t)src;
        const uint32_t mis = (uint32_t)(a & 3);
        w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        nwords = (uint32_t)((mis + len + 3) >> 2);
        used = 0; total = len * 8;
        bb = 0; bn = 0; widx = 0;
        if (nwords) { bb = __byte_perm(__ldg(w), 0, 0x0123); bn = 32 - 8 * mis; widx = 1; }
    }
    __device__ __forceinline__ bool has(uint32_t n) const { return used + n <= total; }
    // next n (0..32) bits, MSB first (src/bitio/mod.rs:78-120); caller checked has(n)
    __device__ __forceinline__ uint32_t take(uint32_t n) {
        if (bn < n) {
            uint32_t x = (widx < nwords) ? __byte_perm(__ldg(w + widx), 0, 0x0123) : 0u;
            ++widx;
            bb = (bb << 32) | x;
            bn += 32;
        }
        bn -= n;
        used += n;
        const uint32_t mask = n ? (0xFFFFFFFFu >> (32 - n)) : 0u;
        return (uint32_t)(bb >> bn) & mask;
    }
    __device__ __forceinline__ uint64_t take64(uint32_t n) {   // n <= 64
        if (n <= 32) return take(n);
        uint64_t hi = take(n - 32);
        return (hi << 32) | take(32);
    }
};

// ------------------------------------------------------------------ byte sink (decoder output)
// Decoded symbols leave as aligned 32-bit words; the (possibly unaligned) head and tail of a
// block's slot are written byte-wise so neighbouring blocks are never touched.
struct ByteSink {
    uintptr_t dst;
    uint32_t wacc;
    __device__ __forceinline__ void init(uint8_t *d) { dst = (uintptr_t)d; wacc = 0; }
    __device__ __forceinline__ void store_bytes(uintptr_t from, uintptr_t to) const {
        for (uintptr_t p = from; p < to; ++p)
            *reinterpret_cast<uint8_t *>(p) = (uint8_t)(wacc >> (8 * (p & 3)));
    }
    __device__ __forceinline__ void put(uint64_t t, uint32_t sym) {
        const uintptr_t a = dst + t;
        const uint32_t pos = (uint32_t)(a & 3);
        wacc |= sym << (8 * pos);
        if (pos == 3) {
            if (a - 3 >= dst) *reinterpret_cast<uint32_t *>(a - 3) = wacc;
            else store_bytes(dst, a + 1);
            wacc = 0;
        }
    }
    __device__ __forceinline__ void finish(uint64_t n) const {
        const uintptr_t end = dst + n;
        if (end & 3) {
            const uintptr_t ws = end & ~(uintptr_t)3;
            store_bytes(ws > dst ? ws : dst, end);
        }
    }
};

// ------------------------------------------------------------------ decoder
template <typename TW, int CLS>
__global__ void __launch_bounds__(kLaneThreads, 2)
decode_lane_kernel(const LaneDecJob job)
{
    using C = Cls<CLS>;
    using S = typename C::S;
    using P = typename C::P;
    using M = typename C::M;
    extern __shared__ uint4 smem_u4[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t blk = (uint64_t)blockIdx.x * kLaneThreads + threadIdx.x;
    if (blk >= job.n_blocks) return;

    LaneTable<TW> tab;
    tab.t = reinterpret_cast<TW *>(smem_u4) + (size_t)warp * kTabNodes * 32 + lane;
    tab.clear();

    const uint64_t coff = job.comp_off[blk];
    const uint64_t clen = job.comp_off[blk + 1] - coff;
    const uint64_t roff = job.raw_off[blk];
    const uint64_t cap = job.raw_off[blk + 1] - roff;

    const uint32_t c = job.c;
    const S maxv = (S)((c == sizeof(S) * 8) ? ~(S)0 : ((((S)1) << c) - 1));
    const S body = maxv >> 1, half = body + 1;
    const M *magic = reinterpret_cast<const M *>(job.magic);
    const uint32_t tcap = job.tcap;

    BitSource src;
    src.init(job.comp + coff, clen);
    ByteSink out;
    out.init(job.raw + roff);

    int32_t st = 0;
    uint64_t t = 0;
    S low = 0, high = maxv, value = 0;
    // src/codec.rs:124-127: prime code_bits bits
    if (!src.has(c)) { st = 1; src.used = src.total; }
    else value = (S)src.take64(c);

    while (st == 0) {
        const uint32_t tt = t < tcap ? (uint32_t)t : tcap;
        const uint32_t count = kNsym + tt;
        const M g = C::ldm(magic + tt);
        // src/codec.rs:129-131 without the division: find i with cum(i)*range <= X < cum(i+1)*range,
        // X = (value-low+1)*count - 1.
        const S rm1 = high - low;
        const P X = C::mulr(count, (S)(value - low)) - 1;      // (value-low+1)*count - 1
        // step m = 256: node 256 = count-1 (adaptive_tree.rs:119-127 with the unstored node)
        P plo = 0, phi = C::mulr(count - 1, rm1);
        uint32_t sym;
        if (X >= phi) {
            sym = kEof; plo = phi; phi = C::mulr(count, rm1);
        } else {
            uint32_t i = 0;
#pragma unroll
            for (int m = 128; m >= 1; m >>= 1) {
                const uint32_t ti = i + m;
                const uint32_t tv = (uint32_t)m + tab.ld(ti);
                const P p = plo + C::mulr(tv, rm1);
                if (X >= p) { i = ti; plo = p; } else { phi = p; }
            }
            sym = i;
        }
        // src/codec.rs:133-134
        high = low + (S)C::divc(phi, g, count) - 1;
        low = low + (S)C::divc(plo, g, count);
        if (sym == kEof) break;                                  // src/codec.rs:136-138
        if (t < tcap) tab.update(sym);
        // src/codec.rs:140-158 in closed form
        const Renorm<S> r = renorm<S>(low, high, c);
        const uint32_t n = r.n1 + r.k;
        if (!src.has(n)) { st = 1; src.used = src.total; break; }   // Err(Eof) inside get_bit
        if (t >= cap) { st = 6; break; }                            // sink full
        const S chunk = (S)src.take64(n);
        S v1 = (r.n1 >= sizeof(S) * 8) ? (S)0 : (S)((value << r.n1) & maxv);
        v1 |= (r.k >= sizeof(S) * 8) ? (S)0 : (S)(chunk >> r.k);
        value = (v1 & half) | ((S)(v1 << r.k) & body) | (chunk & (S)((((S)1) << r.k) - 1));
        low = r.low; high = r.high;
        out.put(t, sym);
        ++t;
    }
    out.finish(t);
    job.raw_len[blk] = t;
    job.consumed[blk] = (src.used + 7) >> 3;
    job.status[blk] = st;
}

}  // namespace rdx
