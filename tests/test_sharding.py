"""Multi-GPU host logic on CPU: record sharding keeps the output byte-identical (world_size 2 over
gloo; the per-shard translation is done by the oracle here, which is what the CUDA path is checked
against in test_gpu_parity.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fasta_gen import nasty_fasta, regular_fasta
from gotranseq_b200 import shard


def test_shard_ranges_cut_at_record_starts():
    data = regular_fasta(3, 50, 10, 500)
    for n in (1, 2, 3, 8, 64):
        r = shard.shard_ranges(data, n)
        assert r[0][0] == 0 and r[-1][1] == len(data)
        for (a, b), (c, d) in zip(r, r[1:]):
            assert b == c and a <= b
        for a, b in r[1:]:
            if a < len(data) and b > a:
                assert data[a:a + 1] == b">" and data[a - 1:a] == b"\n"


@pytest.mark.parametrize("seed", range(25))
def test_sharded_output_is_identical(seed):
    data = nasty_fasta(seed, max_records=20)
    kw = dict(frame="6", trim=bool(seed & 1), alternative=bool(seed & 2))
    want = oracle.translate(data, **kw)
    for world in (2, 3, 5):
        got = b"".join(shard.translate_shard(data, r, world, lambda b: oracle.translate(b, **kw)) for r in range(world))
        assert got == want, (seed, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, kw, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = shard.translate_shard(data, rank, world, lambda b: oracle.translate(b, **kw))
    # order restored by the host: gather sizes, then the bytes, in rank order (no data-path collective
    # in the product; this only checks the plumbing bench.py uses for its max-over-ranks timing)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(out)], dtype=torch.int64))
    mx = int(max(int(s) for s in sizes))
    buf = torch.zeros(max(mx, 1), dtype=torch.uint8)
    if out:
        buf[: len(out)] = torch.frombuffer(bytearray(out), dtype=torch.uint8)
    bufs = [torch.zeros(max(mx, 1), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank == 0:
        q.put(b"".join(bytes(b[: int(s)].numpy()) for b, s in zip(bufs, sizes)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    data = regular_fasta(9, 400, 50, 3000)
    kw = dict(frame="6", table=11, clean=True, trim=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, kw, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == oracle.translate(data, **kw)
