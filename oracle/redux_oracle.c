/*
 * redux_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See redux_oracle.h.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * peterbudai/redux tree).  All state is uint64_t with wrapping arithmetic, which is
 * what `cargo test --release` (the reference's CI mode, .travis.yml:5-7) executes.
 * The structure deliberately follows the reference statement by statement (bit-at-a-time
 * output, renormalisation as a loop, 1-byte bit buffer): this file is the checker, the
 * closed forms live in the CUDA path and are tested AGAINST this file.
 */
#include "redux_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- Parameters */

/* src/model/mod.rs:63-81 */
int oracle_params_new(uint64_t symbol, uint64_t frequency, uint64_t code, oracle_params *p)
{
    /* mod.rs:64 -- the four rejection clauses, in the reference's order */
    if (symbol < 1 || frequency < symbol + 2 || code < frequency + 2 || 64 < code + frequency)
        return ORACLE_INVALID_INPUT;
    /* the oracle additionally needs 1<<symbol to be an allocatable table size */
    if (symbol > 24)
        return ORACLE_INVALID_INPUT;
    p->symbol_bits = symbol;
    p->symbol_eof = (uint64_t)1 << symbol;          /* mod.rs:69 */
    p->symbol_count = ((uint64_t)1 << symbol) + 1;  /* mod.rs:70 */
    p->freq_bits = frequency;
    p->freq_max = ((uint64_t)1 << frequency) - 1;   /* mod.rs:72 */
    p->code_bits = code;
    p->code_min = 0;                                /* mod.rs:74 */
    p->code_one_fourth = (uint64_t)1 << (code - 2); /* mod.rs:75 */
    p->code_half = (uint64_t)2 << (code - 2);       /* mod.rs:76 */
    p->code_three_fourths = (uint64_t)3 << (code - 2); /* mod.rs:77 */
    p->code_max = ((uint64_t)1 << code) - 1;        /* mod.rs:78 (code <= 61 here since freq >= 3) */
    return ORACLE_OK;
}

/* -------------------------------------------------------------------- Models */

struct oracle_model {
    int kind;
    oracle_params params;
    uint64_t *table; /* linear: freq[symbol_count+1] cumulative; tree: Fenwick tree[symbol_count+1] */
    uint64_t len;    /* symbol_count + 1 */
    uint64_t count;  /* tree only: cached total (adaptive_tree.rs:13-15) */
};

static inline uint64_t last_one(uint64_t x) { return x & (0 - x); } /* adaptive_tree.rs:28-32 */

oracle_model *oracle_model_new(int kind, const oracle_params *p)
{
    oracle_model *m = (oracle_model *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    m->kind = kind;
    m->params = *p;
    m->len = p->symbol_count + 1;
    m->table = (uint64_t *)calloc(m->len, sizeof(uint64_t));
    if (!m->table) { free(m); return NULL; }
    if (kind == ORACLE_MODEL_LINEAR) {
        /* adaptive_linear.rs:21-30: freq[0]=0, freq[i]=i */
        for (uint64_t i = 1; i < m->len; i++) m->table[i] = i;
    } else {
        /* adaptive_tree.rs:36-48: tree[i] = last_one(i), count = symbol_count */
        for (uint64_t i = 0; i < m->len; i++) m->table[i] = last_one(i);
        m->count = p->symbol_count;
    }
    return m;
}

void oracle_model_free(oracle_model *m)
{
    if (m) { free(m->table); free(m); }
}

/* adaptive_linear.rs:47-49 / adaptive_tree.rs:100-103 */
uint64_t oracle_model_total_frequency(const oracle_model *m)
{
    return m->kind == ORACLE_MODEL_LINEAR ? m->table[m->params.symbol_count] : m->count;
}

/* adaptive_linear.rs:33-39 */
static void linear_update(oracle_model *m, uint64_t symbol)
{
    if (oracle_model_total_frequency(m) < m->params.freq_max)
        for (uint64_t i = symbol + 1; i < m->len; i++) m->table[i] += 1;
}

/* adaptive_tree.rs:51-59 */
static uint64_t tree_get_frequency_single(const oracle_model *m, uint64_t symbol)
{
    uint64_t i = symbol, sum = m->table[0];
    while (i > 0) { sum += m->table[i]; i -= last_one(i); }
    return sum;
}

/* adaptive_tree.rs:63-80 */
static void tree_get_frequency_range(const oracle_model *m, uint64_t symbol, uint64_t *lo, uint64_t *hi)
{
    uint64_t sumh = 0, suml = 0, h = symbol + 1, l = symbol;
    while (h != l) {
        if (h > l) { sumh += m->table[h]; h -= last_one(h); }
        else       { suml += m->table[l]; l -= last_one(l); }
    }
    uint64_t sumr = tree_get_frequency_single(m, h);
    *lo = suml + sumr;
    *hi = sumh + sumr;
}

/* adaptive_tree.rs:83-92 */
static void tree_update(oracle_model *m, uint64_t symbol)
{
    if (oracle_model_total_frequency(m) < m->params.freq_max) {
        uint64_t i = symbol;
        while (i <= m->params.symbol_count) { m->table[i] += 1; i += last_one(i); }
        m->count += 1;
    }
}

/* adaptive_linear.rs:51-59 / adaptive_tree.rs:105-113: lookup THEN update */
int oracle_model_get_frequency(oracle_model *m, uint64_t symbol, uint64_t *lo, uint64_t *hi)
{
    if (symbol > m->params.symbol_eof) return ORACLE_INVALID_INPUT;
    if (m->kind == ORACLE_MODEL_LINEAR) {
        *lo = m->table[symbol];
        *hi = m->table[symbol + 1];
        linear_update(m, symbol);
    } else {
        tree_get_frequency_range(m, symbol, lo, hi);
        tree_update(m, symbol + 1);
    }
    return ORACLE_OK;
}

/* adaptive_linear.rs:61-70 / adaptive_tree.rs:115-136 */
int oracle_model_get_symbol(oracle_model *m, uint64_t value, uint64_t *symbol, uint64_t *lo, uint64_t *hi)
{
    if (m->kind == ORACLE_MODEL_LINEAR) {
        for (uint64_t i = 0; i < m->len - 1; i++) {
            if (value < m->table[i + 1]) {
                *symbol = i; *lo = m->table[i]; *hi = m->table[i + 1];
                linear_update(m, i);
                return ORACLE_OK;
            }
        }
        return ORACLE_INVALID_INPUT;
    }
    uint64_t mm = m->params.symbol_eof, i = 0, v = value;
    while (mm > 0 && i < m->params.symbol_eof) { /* adaptive_tree.rs:119-127 */
        uint64_t ti = i + mm, tv = m->table[ti];
        if (v >= tv) { i = ti; v -= tv; }
        mm >>= 1;
    }
    uint64_t l, h;
    tree_get_frequency_range(m, i, &l, &h);
    if (value >= h) return ORACLE_INVALID_INPUT; /* adaptive_tree.rs:130-131 */
    tree_update(m, i + 1);
    *symbol = i; *lo = l; *hi = h;
    return ORACLE_OK;
}

void oracle_model_get_freq_table(const oracle_model *m, uint64_t *out)
{
    for (uint64_t i = 0; i < m->params.symbol_count; i++) {
        if (m->kind == ORACLE_MODEL_LINEAR) { out[2 * i] = m->table[i]; out[2 * i + 1] = m->table[i + 1]; }
        else { out[2 * i] = tree_get_frequency_single(m, i); out[2 * i + 1] = tree_get_frequency_single(m, i + 1); }
    }
}

/* -------------------------------------------------------------------- Bit I/O */

/* BitBuffer + BitWriter (src/bitio/mod.rs:33-51, 124-199) over a bounded memory sink. */
struct oracle_bitwriter {
    uint8_t byte;   /* buffer.bytes[0] */
    uint64_t bits;  /* buffer.bits */
    uint64_t count; /* buffer.count */
    uint8_t *out; size_t cap;
};

oracle_bitwriter *oracle_bitwriter_new(uint8_t *buf, size_t cap)
{
    oracle_bitwriter *w = (oracle_bitwriter *)calloc(1, sizeof(*w));
    if (w) { w->out = buf; w->cap = cap; }
    return w;
}
void oracle_bitwriter_free(oracle_bitwriter *w) { free(w); }
uint64_t oracle_bitwriter_get_count(const oracle_bitwriter *w) { return w->count; } /* mod.rs:141-145 */

/* mod.rs:183-198 */
static inline int bw_flush(oracle_bitwriter *w)
{
    if (w->bits > 0) {
        w->byte = (uint8_t)(w->byte << (8 - w->bits));
        if (w->count >= w->cap) return ORACLE_IO_ERROR; /* write_all failed */
        w->out[w->count] = w->byte;
        w->count += 1; w->byte = 0; w->bits = 0;
    }
    return ORACLE_OK;
}

/* mod.rs:148-181 */
static inline int bw_write(oracle_bitwriter *w, uint64_t symbol, uint64_t bits)
{
    if (bits > 64 || (bits < 64 && (symbol >> bits) > 0)) return ORACLE_INVALID_INPUT;
    while (bits > 0) {
        if (w->bits + bits <= 8) {
            if (w->bits > 0) w->byte = (uint8_t)(w->byte << bits);
            w->byte |= (uint8_t)symbol;
            w->bits += bits; bits = 0; symbol = 0;
        } else if (w->bits < 8) {
            uint64_t num = 8 - w->bits;
            if (w->bits > 0) w->byte = (uint8_t)(w->byte << num);
            w->byte |= (uint8_t)(symbol >> (bits - num));
            w->bits += num; bits -= num;
            symbol &= ((uint64_t)1 << bits) - 1;
        }
        if (w->bits == 8) { int e = bw_flush(w); if (e) return e; }
    }
    return ORACLE_OK;
}
int oracle_bitwriter_write_bits(oracle_bitwriter *w, uint64_t s, uint64_t b) { return bw_write(w, s, b); }
int oracle_bitwriter_flush_bits(oracle_bitwriter *w) { return bw_flush(w); }

/* BitReader (src/bitio/mod.rs:54-121) over a memory source. */
struct oracle_bitreader {
    uint8_t byte; uint64_t bits; uint64_t count;
    const uint8_t *in; size_t len;
};

oracle_bitreader *oracle_bitreader_new(const uint8_t *buf, size_t len)
{
    oracle_bitreader *r = (oracle_bitreader *)calloc(1, sizeof(*r));
    if (r) { r->in = buf; r->len = len; }
    return r;
}
void oracle_bitreader_free(oracle_bitreader *r) { free(r); }
uint64_t oracle_bitreader_get_count(const oracle_bitreader *r) { return r->count; } /* mod.rs:71-75 */

/* mod.rs:78-120 */
static inline int br_read(oracle_bitreader *r, uint64_t bits, uint64_t *out)
{
    if (bits > 64) return ORACLE_INVALID_INPUT;
    uint64_t result = 0;
    while (bits > 0) {
        if (r->bits >= bits) {
            result = (bits < 64) ? (result << bits) : 0;
            result |= (uint64_t)r->byte >> (r->bits - bits);
            r->bits -= bits;
            r->byte &= (uint8_t)(((unsigned)1 << r->bits) - 1);
            bits = 0;
        } else if (r->bits > 0) {
            result <<= r->bits;
            result |= r->byte;
            bits -= r->bits;
            r->byte = 0; r->bits = 0;
        } else {
            if (r->count >= r->len) return ORACLE_EOF; /* read() returned 0, mod.rs:106-108 */
            r->byte = r->in[r->count];
            r->count += 1; r->bits = 8;
        }
    }
    *out = result;
    return ORACLE_OK;
}
int oracle_bitreader_read_bits(oracle_bitreader *r, uint64_t b, uint64_t *o) { return br_read(r, b, o); }

/* ---------------------------------------------------------------------- Codec */

/* struct Codec (src/codec.rs:11-24) */
typedef struct {
    uint64_t low, high, pending, extra;
    oracle_model *model;
} codec;

/* codec.rs:28-36 */
static void codec_new(codec *c, oracle_model *m)
{
    c->low = m->params.code_min;
    c->high = m->params.code_max;
    c->pending = 0;
    c->extra = m->params.code_bits;
    c->model = m;
}

/* codec.rs:39-46 */
static inline int put_bit(codec *c, int bit, oracle_bitwriter *out)
{
    int e = bw_write(out, bit ? 1 : 0, 1);
    if (e) return e;
    while (c->pending > 0) {
        e = bw_write(out, bit ? 0 : 1, 1);
        if (e) return e;
        c->pending -= 1;
    }
    return ORACLE_OK;
}

/* codec.rs:49-52 */
static inline int get_bit(codec *c, oracle_bitreader *in)
{
    uint64_t b;
    int e = br_read(in, 1, &b);
    if (e) return e;
    c->pending = (c->pending << 1) | b;
    return ORACLE_OK;
}

/* codec.rs:55-101 */
static int compress_symbol(codec *c, uint64_t symbol, oracle_bitwriter *out)
{
    const oracle_params *p = &c->model->params;
    uint64_t count = oracle_model_total_frequency(c->model);      /* :56 (before the lookup mutates) */
    uint64_t lo, hi;
    int e = oracle_model_get_frequency(c->model, symbol, &lo, &hi); /* :57 */
    if (e) return e;
    uint64_t range = c->high - c->low + 1;                        /* :58 */
    c->high = c->low + (range * hi / count) - 1;                  /* :59 */
    c->low = c->low + (range * lo / count);                       /* :60 */

    for (;;) {                                                    /* :62-89 */
        if (c->high < p->code_half) {
            if ((e = put_bit(c, 0, out))) return e;
            if (symbol == p->symbol_eof) c->extra -= 1;
        } else if (c->low >= p->code_half) {
            if ((e = put_bit(c, 1, out))) return e;
            if (symbol == p->symbol_eof) c->extra -= 1;
        } else if (c->low >= p->code_one_fourth && c->high < p->code_three_fourths) {
            c->pending += 1;
            c->low -= p->code_one_fourth;
            c->high -= p->code_one_fourth;
            if (symbol == p->symbol_eof) c->extra -= 1;
        } else {
            break;
        }
        c->high = ((c->high << 1) + 1) & p->code_max;
        c->low = (c->low << 1) & p->code_max;
    }

    if (symbol == p->symbol_eof) {                                /* :91-99 */
        while (c->extra > 0) {
            uint64_t mask = c->low & p->code_half;
            if ((e = put_bit(c, mask != 0, out))) return e;
            c->low = (c->low << 1) & p->code_max;
            c->extra -= 1;
        }
        if ((e = bw_flush(out))) return e;
    }
    return ORACLE_OK;
}

/* codec.rs:104-120 */
static int compress_stream(codec *c, oracle_bitreader *in, oracle_bitwriter *out)
{
    const oracle_params *p = &c->model->params;
    for (;;) {
        uint64_t symbol;
        int e = br_read(in, p->symbol_bits, &symbol);
        if (e == ORACLE_EOF) symbol = p->symbol_eof;
        else if (e) return e;
        if ((e = compress_symbol(c, symbol, out))) return e;
        if (symbol == p->symbol_eof) break;
    }
    return ORACLE_OK;
}

/* codec.rs:123-161 */
static int decompress_symbol(codec *c, oracle_bitreader *in, uint64_t *symbol_out)
{
    const oracle_params *p = &c->model->params;
    int e;
    while (c->extra > 0) {                                        /* :124-127 */
        if ((e = get_bit(c, in))) return e;
        c->extra -= 1;
    }
    uint64_t range = c->high - c->low + 1;                        /* :129 */
    uint64_t count = oracle_model_total_frequency(c->model);      /* :130 */
    uint64_t value = ((c->pending - c->low + 1) * count - 1) / range; /* :131 */
    uint64_t symbol, lo, hi;
    if ((e = oracle_model_get_symbol(c->model, value, &symbol, &lo, &hi))) return e; /* :132 */
    c->high = c->low + (range * hi / count) - 1;                  /* :133 */
    c->low = c->low + (range * lo / count);                       /* :134 */
    *symbol_out = symbol;
    if (symbol == p->symbol_eof) return ORACLE_OK;                /* :136-138 */

    for (;;) {                                                    /* :140-158 */
        if (c->high < p->code_half) {
            /* do nothing */
        } else if (c->low >= p->code_half) {
            c->pending -= p->code_half; c->low -= p->code_half; c->high -= p->code_half;
        } else if (c->low >= p->code_one_fourth && c->high < p->code_three_fourths) {
            c->pending -= p->code_one_fourth; c->low -= p->code_one_fourth; c->high -= p->code_one_fourth;
        } else {
            break;
        }
        c->low = c->low << 1;
        c->high = (c->high << 1) + 1;
        if ((e = get_bit(c, in))) return e;
    }
    return ORACLE_OK;
}

/* codec.rs:164-176 -- note: never calls flush_bits */
static int decompress_stream(codec *c, oracle_bitreader *in, oracle_bitwriter *out)
{
    const oracle_params *p = &c->model->params;
    for (;;) {
        uint64_t symbol;
        int e = decompress_symbol(c, in, &symbol);
        if (e) return e;
        if (symbol == p->symbol_eof) break;
        if ((e = bw_write(out, symbol, p->symbol_bits))) return e;
    }
    return ORACLE_OK;
}

/* --------------------------------------------------------------------- Facade */

size_t oracle_compress_bound(size_t in_len, uint64_t symbol_bits, uint64_t code_bits)
{
    size_t nsym = (in_len * 8) / (size_t)symbol_bits + 1; /* data symbols + EOF */
    return (nsym * (size_t)code_bits + 7) / 8;
}

/* n_train > 0: the caller trains the Box<Model> before handing it over, exactly as a user of the
 * reference can: Model::get_frequency(symbol) looks up AND updates (src/model/mod.rs:23-25), and
 * compress()/decompress() accept whatever model they are given (src/lib.rs:102,113). */
static int run(int decode, int kind, uint64_t s, uint64_t f, uint64_t cb,
               const uint64_t *train, size_t n_train,
               const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
               uint64_t *in_count, uint64_t *out_count)
{
    oracle_params p;
    int e = oracle_params_new(s, f, cb, &p);
    if (e) { if (in_count) *in_count = 0; if (out_count) *out_count = 0; return e; }
    oracle_model *m = oracle_model_new(kind, &p);
    if (!m) return ORACLE_IO_ERROR;
    for (size_t i = 0; i < n_train; i++) {
        uint64_t lo, hi;
        if ((e = oracle_model_get_frequency(m, train[i], &lo, &hi))) { oracle_model_free(m); return e; }
    }
    codec c;
    codec_new(&c, m);                                             /* lib.rs:103 / :114 */
    struct oracle_bitreader r = {0, 0, 0, in, in_len};            /* lib.rs:104 */
    struct oracle_bitwriter w = {0, 0, 0, out, out_cap};          /* lib.rs:105 */
    e = decode ? decompress_stream(&c, &r, &w) : compress_stream(&c, &r, &w); /* lib.rs:107 / :118 */
    if (in_count) *in_count = r.count;                            /* lib.rs:108 */
    if (out_count) *out_count = w.count;
    oracle_model_free(m);
    return e;
}

int oracle_compress(int kind, uint64_t s, uint64_t f, uint64_t c, const uint8_t *in, size_t in_len,
                    uint8_t *out, size_t out_cap, uint64_t *in_count, uint64_t *out_count)
{
    return run(0, kind, s, f, c, NULL, 0, in, in_len, out, out_cap, in_count, out_count);
}

int oracle_decompress(int kind, uint64_t s, uint64_t f, uint64_t c, const uint8_t *in, size_t in_len,
                      uint8_t *out, size_t out_cap, uint64_t *in_count, uint64_t *out_count)
{
    return run(1, kind, s, f, c, NULL, 0, in, in_len, out, out_cap, in_count, out_count);
}

int oracle_compress_trained(int kind, uint64_t s, uint64_t f, uint64_t c, const uint64_t *train, size_t n_train,
                            const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                            uint64_t *in_count, uint64_t *out_count)
{
    return run(0, kind, s, f, c, train, n_train, in, in_len, out, out_cap, in_count, out_count);
}

int oracle_decompress_trained(int kind, uint64_t s, uint64_t f, uint64_t c, const uint64_t *train, size_t n_train,
                              const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                              uint64_t *in_count, uint64_t *out_count)
{
    return run(1, kind, s, f, c, train, n_train, in, in_len, out, out_cap, in_count, out_count);
}

/* ------------------------------------------------- CPU baseline batch driver */

typedef struct {
    int decode, kind; uint64_t s, f, c;
    const uint8_t *in; const uint64_t *in_off; uint64_t n;
    uint8_t *out; const uint64_t *out_off; uint64_t *out_len; uint64_t *consumed; int32_t *status;
    uint64_t next; /* shared work counter */
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    for (;;) {
        uint64_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        uint64_t ic = 0, oc = 0;
        int e = run(j->decode, j->kind, j->s, j->f, j->c, NULL, 0,
                    j->in + j->in_off[i], (size_t)(j->in_off[i + 1] - j->in_off[i]),
                    j->out + j->out_off[i], (size_t)(j->out_off[i + 1] - j->out_off[i]), &ic, &oc);
        j->out_len[i] = oc;
        if (j->consumed) j->consumed[i] = ic;
        j->status[i] = e;
    }
    return NULL;
}

static int run_batch(batch_job *j, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint64_t)n_threads > j->n && j->n > 0) n_threads = (int)j->n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int t = 1; t < n_threads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, j) == 0) started++;
    batch_worker(j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    free(th);
    for (uint64_t i = 0; i < j->n; i++) if (j->status[i]) return j->status[i];
    return ORACLE_OK;
}

int oracle_compress_batch(int kind, uint64_t s, uint64_t f, uint64_t c,
                          const uint8_t *in, const uint64_t *in_off, uint64_t n_blocks,
                          uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                          int32_t *status, int n_threads)
{
    batch_job j = {0, kind, s, f, c, in, in_off, n_blocks, out, out_off, out_len, NULL, status, 0};
    return run_batch(&j, n_threads);
}

int oracle_decompress_batch(int kind, uint64_t s, uint64_t f, uint64_t c,
                            const uint8_t *in, const uint64_t *in_off, uint64_t n_blocks,
                            uint8_t *out, const uint64_t *out_off, uint64_t *out_len,
                            uint64_t *consumed, int32_t *status, int n_threads)
{
    batch_job j = {1, kind, s, f, c, in, in_off, n_blocks, out, out_off, out_len, consumed, status, 0};
    return run_batch(&j, n_threads);
}
