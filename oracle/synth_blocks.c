/*
 * synth_blocks.c -- TEST INFRASTRUCTURE: the synthetic mixed-entropy blocks of BASELINE.json configs 3-4, written
 * out independently of the product (redux_b200/csrc/redux_common.cuh holds the device/host generator of the
 * product; tests/test_capi_cpu.py proves the two produce the same bytes).  bench.py's reference arm and
 * cpu_baseline generate their input here, so that the CPU arm never maps the product library.
 *
 * 8 output bytes per 64-bit draw: draw(block, w) = mix(seed + block * K + (w + 1) * GAMMA), mix = the
 * splitmix64 finaliser.  Class = block & 3: uniform bytes / text-like (256-slot Zipf table over 55 symbols)
 * / geometric (ctz of a 16-bit field) / sparse (0x00 with p = 63/64).
 */
#include <stdint.h>
#include <stddef.h>

#include "redux_oracle.h"

static uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void text_table(uint8_t lut[256])
{
    static const char alphabet[] = " etaoinshrdlucmfwypvbgk,.\n-'\";ETAOINSHRDLUCMFWYPVBGKjxqz0123456";
    uint32_t slot = 0;
    for (uint32_t r = 0; r < 64 && slot < 256; ++r) {      /* rank r owns max(1, 60/(r+1)) slots */
        uint32_t n = 60 / (r + 1);
        if (n < 1) n = 1;
        for (uint32_t k = 0; k < n && slot < 256; ++k) lut[slot++] = (uint8_t)alphabet[r];
    }
    while (slot < 256) lut[slot++] = (uint8_t)alphabet[63];
}

void oracle_generate_blocks(uint8_t *out, uint64_t first_block, uint64_t n_blocks, uint64_t block_len, uint64_t seed)
{
    oracle_generate_blocks_ex(out, first_block, n_blocks, block_len, seed, NULL, 0);
}

/* With a corpus of at least block_len bytes the text class (block & 3 == 1) is a block_len-byte window of it at
 * offset draw(block, 2^64 - 2) % (corpus_len - block_len + 1). */
void oracle_generate_blocks_ex(uint8_t *out, uint64_t first_block, uint64_t n_blocks, uint64_t block_len, uint64_t seed,
                               const uint8_t *corpus, uint64_t corpus_len)
{
    uint8_t lut[256];
    text_table(lut);
    for (uint64_t b = 0; b < n_blocks; ++b) {
        const uint64_t block = first_block + b;
        const uint32_t cls = (uint32_t)(block & 3);
        uint8_t *dst = out + b * block_len;
        if (cls == 1 && corpus && block_len && corpus_len >= block_len) {
            const uint64_t at = mix64(seed + block * 0xD1342543DE82EF95ull + (~(uint64_t)0) * 0x9E3779B97F4A7C15ull)
                                % (corpus_len - block_len + 1);
            for (uint64_t i = 0; i < block_len; ++i) dst[i] = corpus[at + i];
            continue;
        }
        for (uint64_t w = 0; w * 8 < block_len; ++w) {
            const uint64_t r = mix64(seed + block * 0xD1342543DE82EF95ull + (w + 1) * 0x9E3779B97F4A7C15ull);
            const uint64_t r2 = mix64(r ^ 0xA5A5A5A5A5A5A5A5ull);
            for (uint32_t j = 0; j < 8 && w * 8 + j < block_len; ++j) {
                uint8_t v;
                if (cls == 0) {
                    v = (uint8_t)(r >> (8 * j));
                } else if (cls == 1) {
                    v = lut[(r >> (8 * j)) & 255];
                } else {
                    const uint32_t u = (uint32_t)(((j < 4 ? r : r2) >> (16 * (j & 3))) & 0xFFFF);
                    if (cls == 2) v = (uint8_t)__builtin_ctz(u | 0x8000u);
                    else          v = (u < 0xFC00u) ? 0 : (uint8_t)(u & 0xFF);
                }
                dst[w * 8 + j] = v;
            }
        }
    }
}
