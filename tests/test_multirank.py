"""The N>1 host logic on CPU: two gloo ranks shard one BGZF file with hgpu_shard_range, and the
shards tile the block list / byte stream exactly (no data-path collective; only the per-rank
lengths are gathered, as bench.py / DESIGN.md §5 describe)."""
import os, random, socket, sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, img_bytes, q):
    import htslib_b200 as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = np.frombuffer(img_bytes, dtype=np.uint8)
    off, ln, isz = H.bgzf_scan(img)
    first, count, base = H.shard_range(isz, world, rank)
    mine = int(isz[first:first + count].astype(np.int64).sum())
    # the only exchange: 8 bytes per rank (SURVEY.md §8e "optionally all-gather of output lengths")
    t = torch.tensor([mine], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    lens = [int(g.item()) for g in gathered]
    assert base == sum(lens[:rank])
    q.put((rank, first, count, base, mine, len(off)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_tiles_the_file():
    from _libs import bgzf_file
    rng = random.Random(1)
    data = bytes(rng.randrange(256) for _ in range(700001))
    img = bgzf_file(data, 1, block=9973)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, img, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60); assert p.exitcode == 0
    nblk = res[0][5]
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == nblk
    assert res[0][3] == 0 and res[1][3] == res[0][4] and res[0][4] + res[1][4] == len(data)
    assert abs(res[0][2] - res[1][2]) <= 1


def test_shard_range_properties():
    import htslib_b200 as H
    lens = np.arange(1, 1001, dtype=np.uint32)
    for world in (1, 2, 3, 4, 8, 7):
        cover, base = 0, 0
        for r in range(world):
            f, c, b = H.shard_range(lens, world, r)
            assert f == cover and b == base
            cover += c; base += int(lens[f:f + c].astype(np.int64).sum())
        assert cover == 1000 and base == int(lens.astype(np.int64).sum())
    f, c, b = H.shard_range(np.zeros(0, dtype=np.uint32), 4, 2)
    assert (f, c, b) == (0, 0, 0)


def _cram_worker(rank, world, port, path, q):
    import htslib_b200 as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = np.fromfile(path, dtype=np.uint8)
    cont, _ = H.cram_scan_containers(img)
    blocks, _ = H.cram_scan_blocks(img)
    data = np.where(cont["n_records"] > 0)[0]                        # data containers are the shardable units
    size = np.array([int(blocks[blocks["container"] == i]["uncomp_size"].astype(np.int64).sum()) for i in data], dtype=np.uint32)
    first, count, base = H.shard_range(size, world, rank)
    mine = data[first:first + count]
    recs = int(cont["n_records"][mine].sum())
    nblk = int(cont["n_blocks"][mine].sum())
    t = torch.tensor([recs, nblk], dtype=torch.int64)
    g = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(g, t)
    q.put((rank, [int(i) for i in mine], recs, nblk, [x.tolist() for x in g], int(cont["n_records"].sum()),
           int(cont["n_blocks"][data].sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_cram_containers():
    """CRAM: containers (one slice each here) are the independent units; two ranks take contiguous container ranges
    that cover every record and every block once."""
    from _libs import GOLD
    path = os.path.join(GOLD, "htslib", "ce#1000.v31arith.cram")       # 4 data containers of 300/300/300/100 reads
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cram_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60); assert p.exitcode == 0
    a, b = res
    assert a[1] and b[1] and a[1][-1] + 1 == b[1][0]                    # contiguous, in file order
    assert a[2] + b[2] == a[5] == 1000 and a[3] + b[3] == a[6]
    assert a[4] == b[4] == [[a[2], a[3]], [b[2], b[3]]]                 # what the all-gather delivered
