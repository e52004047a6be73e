"""Host-side multi-rank logic on CPU: world_size-2 `gloo` process group, no GPU.
Each rank takes its contiguous shard (the C front end's rule), "codes" it with the oracle (tests may),
the ranks exchange only sizes + timings, and the assembled global result must equal the single-process one."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as o
import redux_b200 as rb
from redux_b200 import sharding

SEED = 0x5EED202610180000
N, L, PARAMS = 37, 700, (8, 14, 16)          # 37 blocks: an uneven split over 2 ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_range(N, world, rank)
    raw = rb.generate_blocks_host(first, count, L, SEED)
    off = np.arange(count + 1, dtype=np.uint64) * np.uint64(L)
    rc, slots, slot_off, out_len, status = o.compress_batch(raw, off, o.TREE, PARAMS, threads=1)
    assert rc == 0
    shards = sharding.gather_sizes(out_len, dist)
    offsets, bases = sharding.global_offsets(shards)
    mine = b"".join(slots[int(slot_off[i]):int(slot_off[i]) + int(out_len[i])].tobytes() for i in range(count))
    t = sharding.max_over_ranks([0.25 + rank, 1.0 - rank], dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((offsets, bases, b"".join(gathered), t))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rule_matches_c_front_end():
    L_ = rb.lib()
    for n in (0, 1, 2, 7, 37, 65536, 65537):
        for g_total in (1, 2, 3, 4, 8):
            covered = 0
            for g in range(g_total):
                a, c = C.c_uint64(), C.c_uint64()
                L_.redux_debug_shard(n, g_total, g, C.byref(a), C.byref(c))
                assert (a.value, c.value) == sharding.shard_range(n, g_total, g)
                assert a.value == covered
                covered += c.value
            assert covered == n
    assert sharding.weak_first_block(65536, 3) == 196608


def test_two_rank_gloo_assembles_the_single_process_result():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    offsets, bases, blob, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    raw = rb.generate_blocks_host(0, N, L, SEED)
    want = []
    for i in range(N):
        rc, out, _, _ = o.compress(raw[i * L:(i + 1) * L], o.TREE, PARAMS)
        want.append(out)
    assert blob == b"".join(want)
    assert [int(x) for x in offsets] == [0] + list(np.cumsum([len(w) for w in want]))
    assert int(bases[0]) == 0 and int(bases[1]) == sum(len(w) for w in want[:sharding.shard_range(N, 2, 1)[0]])
    assert t == [1.25, 1.0]          # max over ranks, element-wise
