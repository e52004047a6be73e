/*
 * redux_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the peterbudai/redux adaptive arithmetic coder hot path
 * (src/codec.rs, src/model/{mod,adaptive_linear,adaptive_tree}.rs, src/bitio/mod.rs,
 * src/lib.rs:100-120).  It is the parity authority for the CUDA path in
 * redux_b200/csrc and the CPU baseline timed by bench.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (redux_b200/) never links or calls it.
 *
 * PINNING STATUS (be honest about what anchors this file):
 *   - bit packing / byte counting: pinned by the reference's own golden vectors
 *     (src/bitio/tests.rs:8-218), replayed in tests/test_oracle_bitio.py.
 *   - Linear == Tree observational equivalence, invalid symbol / value rejection:
 *     the reference's property tests (src/model/tests.rs:50-93), replayed seeded.
 *   - round trip + exact byte accounting on the corpora: tests/corpora.rs:40-41,59,61.
 *   - ENCODER BYTES: the reference holds NO golden compressed vector and no Rust
 *     toolchain exists in the build container (rustc/cargo absent), so the
 *     compressed bytes are "parity unpinned" by any reference-produced output.
 *     They are cross-checked against (a) the hand-derived known answers of
 *     SURVEY.md Appendix B.1 and (b) an independent pure-Python second reading
 *     (tests/golden/make_golden.py) and the SURVEY.md Appendix B.2 size/SHA table.
 */
#ifndef REDUX_ORACLE_H
#define REDUX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes mirror redux::Error (src/lib.rs:57-64). */
enum {
    ORACLE_OK = 0,
    ORACLE_EOF = 1,           /* Error::Eof */
    ORACLE_INVALID_INPUT = 2, /* Error::InvalidInput */
    ORACLE_IO_ERROR = 3       /* Error::IoError: here = output buffer full */
};

enum { ORACLE_MODEL_LINEAR = 0, ORACLE_MODEL_TREE = 1 };

/* Parameters (src/model/mod.rs:33-59). */
typedef struct {
    uint64_t symbol_bits, symbol_eof, symbol_count;
    uint64_t freq_bits, freq_max;
    uint64_t code_bits, code_min, code_one_fourth, code_half, code_three_fourths, code_max;
} oracle_params;

/* Parameters::new (src/model/mod.rs:63-81). Returns ORACLE_OK or ORACLE_INVALID_INPUT. */
int oracle_params_new(uint64_t symbol, uint64_t frequency, uint64_t code, oracle_params *out);

/* Opaque model: AdaptiveLinearModel / AdaptiveTreeModel behind the Model trait
 * (src/model/mod.rs:17-29). */
typedef struct oracle_model oracle_model;
oracle_model *oracle_model_new(int kind, const oracle_params *p);
void oracle_model_free(oracle_model *m);
uint64_t oracle_model_total_frequency(const oracle_model *m);
int oracle_model_get_frequency(oracle_model *m, uint64_t symbol, uint64_t *lo, uint64_t *hi);
int oracle_model_get_symbol(oracle_model *m, uint64_t value, uint64_t *symbol, uint64_t *lo, uint64_t *hi);
/* debug-only get_freq_table (adaptive_tree.rs:138-146 / adaptive_linear.rs:72-80):
 * writes symbol_count (lo,hi) pairs into out[2*symbol_count]. */
void oracle_model_get_freq_table(const oracle_model *m, uint64_t *out);

/* BitWriter / BitReader over memory buffers (src/bitio/mod.rs:54-199). */
typedef struct oracle_bitwriter oracle_bitwriter;
oracle_bitwriter *oracle_bitwriter_new(uint8_t *buf, size_t cap);
void oracle_bitwriter_free(oracle_bitwriter *w);
int oracle_bitwriter_write_bits(oracle_bitwriter *w, uint64_t symbol, uint64_t bits);
int oracle_bitwriter_flush_bits(oracle_bitwriter *w);
uint64_t oracle_bitwriter_get_count(const oracle_bitwriter *w);

typedef struct oracle_bitreader oracle_bitreader;
oracle_bitreader *oracle_bitreader_new(const uint8_t *buf, size_t len);
void oracle_bitreader_free(oracle_bitreader *r);
int oracle_bitreader_read_bits(oracle_bitreader *r, uint64_t bits, uint64_t *out);
uint64_t oracle_bitreader_get_count(const oracle_bitreader *r);

/* redux::compress (src/lib.rs:102-109): returns error code; counts = (bytes read, bytes written).
 * On error the bytes already written stay in out and the counts are still reported. */
int oracle_compress(int kind, uint64_t symbol_bits, uint64_t freq_bits, uint64_t code_bits,
                    const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                    uint64_t *in_count, uint64_t *out_count);

/* redux::decompress (src/lib.rs:113-120). */
int oracle_decompress(int kind, uint64_t symbol_bits, uint64_t freq_bits, uint64_t code_bits,
                      const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                      uint64_t *in_count, uint64_t *out_count);

/* The same with a model the caller trained first: get_frequency(train[i]) for i < n_train on the fresh
 * model (Model::get_frequency updates, src/model/mod.rs:23-25), then compress / decompress with it. */
int oracle_compress_trained(int kind, uint64_t symbol_bits, uint64_t freq_bits, uint64_t code_bits,
                            const uint64_t *train, size_t n_train,
                            const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                            uint64_t *in_count, uint64_t *out_count);
int oracle_decompress_trained(int kind, uint64_t symbol_bits, uint64_t freq_bits, uint64_t code_bits,
                              const uint64_t *train, size_t n_train,
                              const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                              uint64_t *in_count, uint64_t *out_count);

/* Upper bound of compress() output: every coded symbol (incl. EOF) emits at most code_bits bits
 * (SURVEY.md A.4). */
size_t oracle_compress_bound(size_t in_len, uint64_t symbol_bits, uint64_t code_bits);

/* CPU baseline driver: one independent stream per thread, n_threads workers pulling block
 * indices from a shared counter.  Block i input = in + in_off[i] .. in + in_off[i+1]; output slot
 * = out + out_off[i] with capacity out_off[i+1]-out_off[i]; out_len[i] receives the stream size.
 * Returns the first non-OK status (0 if all OK); status[i] per block.  decode: mirror. */
int oracle_compress_batch(int kind, uint64_t s, uint64_t f, uint64_t c,
                          const uint8_t *in, const uint64_t *in_off, uint64_t n_blocks,
                          uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                          int32_t *status, int n_threads);
int oracle_decompress_batch(int kind, uint64_t s, uint64_t f, uint64_t c,
                            const uint8_t *in, const uint64_t *in_off, uint64_t n_blocks,
                            uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                            uint64_t *consumed, int32_t *status, int n_threads);

/* synth_blocks.c: the synthetic mixed-entropy blocks of BASELINE.json configs 3-4 (block index & 3: uniform /
 * text-like / geometric / sparse), n_blocks blocks of block_len bytes starting at block index first_block. */
void oracle_generate_blocks(uint8_t *out, uint64_t first_block, uint64_t n_blocks, uint64_t block_len, uint64_t seed);
void oracle_generate_blocks_ex(uint8_t *out, uint64_t first_block, uint64_t n_blocks, uint64_t block_len, uint64_t seed,
                               const uint8_t *corpus, uint64_t corpus_len);

#ifdef __cplusplus
}
#endif
#endif /* REDUX_ORACLE_H */
