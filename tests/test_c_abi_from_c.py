"""The boundary is a C ABI: include/redux_b200.h must compile as plain C99 and link against the shared
library from a C program (what a cgo / Rust-bindgen / JNI consumer does). No GPU needed: the program only
calls the host-arithmetic entry points and checks that a computing call fails loudly without a device."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include <stdio.h>
#include <string.h>
#include "redux_b200.h"
int main(void) {
    redux_parameters_t p;
    if (redux_parameters_new(8, 30, 32, &p) != REDUX_OK) return 1;
    if (p.symbol_eof != 256 || p.symbol_count != 257 || p.freq_max != ((1ull << 30) - 1) ||
        p.code_half != (2ull << 30) || p.code_max != 0xFFFFFFFFull) return 2;
    if (redux_parameters_new(8, 9, 16, NULL) != REDUX_INVALID_INPUT) return 3;     /* src/model/mod.rs:64 */
    if (redux_parameters_new(8, 40, 42, NULL) != REDUX_INVALID_INPUT) return 4;    /* code + freq > 64 */
    redux_params_t q = {4, 10, 16}, q17 = {17, 20, 24};
    if (redux_params_supported(&q) != REDUX_OK || redux_params_supported(&q17) != REDUX_UNSUPPORTED) return 5;
    if (redux_compress_bound(0, 16) != 2 || redux_compress_bound_ex(7, 12, 16) != 10) return 6;
    if (strcmp(redux_error_string(REDUX_EOF), "Unexpected end of file") != 0) return 7;
    uint64_t first, count;
    redux_debug_shard(65536, 8, 3, &first, &count);
    if (first != 24576 || count != 8192) return 8;
    redux_ctx_t *ctx = NULL;
    int rc = redux_ctx_create(NULL, 0, &ctx);
    printf("ctx_create=%d\n", rc);
    if (rc == REDUX_OK) {
        redux_params_t params = {8, 14, 16};
        const unsigned char in[5] = {0x72, 0x65, 0x64, 0x75, 0x78};
        unsigned char out[32];
        uint64_t ic = 0, oc = 0;
        rc = redux_compress(ctx, REDUX_MODEL_TREE, &params, in, 5, out, sizeof out, &ic, &oc);
        if (rc != REDUX_OK || ic != 5 || oc != 7 || out[0] != 0x71 || out[6] != 0x10) return 9;
        redux_ctx_destroy(ctx);
        printf("compress ok\n");
    } else if (rc != REDUX_CUDA_ERROR) return 10;
    return 0;
}
'''


def test_header_is_c99_and_links(tmp_path):
    src = tmp_path / "consumer.c"
    src.write_text(PROGRAM)
    exe = tmp_path / "consumer"
    libdir = os.path.join(ROOT, "redux_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lredux_b200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    if torch.cuda.is_available():
        assert "compress ok" in r.stdout
    else:
        assert "ctx_create=4" in r.stdout          # REDUX_CUDA_ERROR: no device, no fallback
