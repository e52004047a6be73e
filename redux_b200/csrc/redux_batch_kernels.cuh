// redux_batch_kernels.cuh -- the block-batching front end's device side (new; the reference has
// no batching, src/lib.rs:102-120 codes one stream per call):
//   build_magic_kernel   per-position count reciprocals (set-up, cached per parameter set)
//   scan_sizes_kernel    exclusive prefix sum of the per-block compressed sizes -> out_offsets
//   compact_kernel       gathers the variable-length streams from their worst-case slots into one
//                        back-to-back buffer with 16-byte stores
//   generate_kernel      synthetic mixed-entropy blocks (bench / tests)
#pragma once
#include "redux_common.cuh"

namespace rdx {

template <typename M> struct MagicMaker;
template <> struct MagicMaker<Magic32> {
    static __device__ Magic32 make(uint32_t d, uint32_t nbits) { return make_magic32(d, nbits); }
};
template <> struct MagicMaker<Magic64> {
    // nbits == 0: the 65-bit scheme of the HUGE class (any 64-bit numerator)
    static __device__ Magic64 make(uint32_t d, uint32_t nbits) { return nbits ? make_magic64(d, nbits) : make_magic65(d); }
};

template <> struct MagicMaker<MagicD> {
    static __device__ MagicD make(uint32_t d, uint32_t) { return make_magicd(d); }
};

// out[tt] = magic of count first + tt for numerators < 2^nbits (first = 257, the fresh byte model's total, for the
// lane kernels; the start total of the call for the generic kernels)
template <typename M>
__global__ void build_magic_kernel(M *out, uint32_t n, uint32_t nbits, uint32_t first = kNsym)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = MagicMaker<M>::make(first + i, nbits);
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan of sizes[n] (u32) into offs[n+1] (u64), one CTA of 1024 threads.  n is the block
// count of one launch (65,536 in the headline config): 64 elements per thread, one pass.
// Also raises *overflow when the total exceeds `capacity`.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
scan_sizes_kernel(const uint32_t *__restrict__ sizes, uint64_t n, uint64_t *__restrict__ offs,
                  uint64_t capacity, int32_t *__restrict__ overflow)
{
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t per = (n + kScanThreads - 1) / kScanThreads;
    const uint64_t beg = (uint64_t)tid * per, end = beg + per < n ? beg + per : n;
    uint64_t local = 0;
    for (uint64_t i = beg; i < end; ++i) local += sizes[i];
    // inclusive scan of `local` across the CTA
    uint64_t x = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= (uint32_t)d) w += y;
        }
        warp_sums[lane] = w;
        if (lane == 31) carry_s = w;
    }
    __syncthreads();
    uint64_t run = x - local + (warp ? warp_sums[warp - 1] : 0);   // exclusive prefix of this thread
    for (uint64_t i = beg; i < end; ++i) { offs[i] = run; run += sizes[i]; }
    if (tid == 0) {
        offs[n] = carry_s;
        *overflow = carry_s > capacity ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Compaction.  Slot i (16-byte aligned, at slots + i*stride) holds sizes[i] bytes; they go to
// out + offs[i], which has arbitrary alignment.  One CTA per block: byte-wise head up to the first
// 16-byte boundary of the destination, then one 16-byte store per thread per step assembled from
// aligned 32-bit loads with a funnel shift, byte-wise tail.  Streams that would end beyond
// `capacity` are not copied and get status REDUX_OUT_CAPACITY (6).
// ---------------------------------------------------------------------------------------------
constexpr int kCompactThreads = 256;
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const uint8_t *__restrict__ slots, uint64_t stride, const uint32_t *__restrict__ sizes,
               const uint64_t *__restrict__ offs, uint64_t n, uint8_t *__restrict__ out,
               uint64_t capacity, int32_t *__restrict__ status)
{
    for (uint64_t blk = blockIdx.x; blk < n; blk += gridDim.x) {
        const uint32_t size = sizes[blk];
        const uint64_t o = offs[blk];
        if (o + size > capacity) {
            if (threadIdx.x == 0) status[blk] = 6;
            continue;
        }
        const uint8_t *src = slots + blk * stride;
        uint8_t *dst = out + o;
        uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
        if (head > size) head = size;
        if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
        const uint32_t nvec = (size - head) >> 4;
        const uint32_t sh = (head & 3) * 8;                        // source bit misalignment
        const uint32_t *srcw = reinterpret_cast<const uint32_t *>(src) + (head >> 2);
        uint4 *dstv = reinterpret_cast<uint4 *>(dst + head);
        for (uint32_t v = threadIdx.x; v < nvec; v += kCompactThreads) {
            const uint32_t *p = srcw + 4 * v;
            uint32_t a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
            uint4 q;
            if (sh) {
                uint32_t a4 = p[4];
                q.x = __funnelshift_r(a0, a1, sh); q.y = __funnelshift_r(a1, a2, sh);
                q.z = __funnelshift_r(a2, a3, sh); q.w = __funnelshift_r(a3, a4, sh);
            } else {
                q = make_uint4(a0, a1, a2, a3);
            }
            dstv[v] = q;
        }
        const uint32_t done = head + (nvec << 4);
        if (threadIdx.x < size - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic blocks: thread <-> one 8-byte group (gen_group of redux_common.cuh).
// ---------------------------------------------------------------------------------------------
__global__ void generate_kernel(uint8_t *out, uint64_t first_block, uint64_t n_blocks,
                                uint64_t block_len, uint64_t seed, const uint8_t *text_lut,
                                const uint8_t *corpus, uint64_t corpus_len)
{
    const uint64_t groups_per_block = (block_len + 7) >> 3;
    const uint64_t total = n_blocks * groups_per_block;
    for (uint64_t gidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total;
         gidx += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = gidx / groups_per_block, w = gidx - b * groups_per_block;
        const uint64_t v = gen_group(seed, first_block + b, w, text_lut, corpus, corpus_len, block_len);
        uint8_t *p = out + b * block_len + w * 8;
        const uint64_t left = block_len - w * 8;
        if (left >= 8 && (((uintptr_t)p & 7) == 0)) *reinterpret_cast<uint64_t *>(p) = v;
        else for (uint32_t j = 0; j < 8 && j < left; ++j) p[j] = (uint8_t)(v >> (8 * j));
    }
}

}  // namespace rdx
