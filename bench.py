#!/usr/bin/env python
"""bench.py — BGZF inflate (+ CRAM rANS-Nx16 decode) throughput on B200, beside the reference's
CPU path on the same box.  Contract: see the task statement / DESIGN.md §Measurement.

One "step" = one pass of the BGZF inflate hot path over the whole synthetic 150 bp BAM
(BASELINE.json configs[1]: 10 GB uncompressed, zlib level 6, one warp per 64 KiB block).
  value      uncompressed GB/s, kernel with inputs resident in HBM (CUDA events)
  e2e        same metric through the C-ABI host entry point (hgpu_bgzf_inflate_file_host) with
             pinned HOST buffers: H2D of the compressed file + kernel + D2H of the output, timed
  roofline   (C+U) algorithmic bytes / kernel time against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref/libhts_ref.so: bgzf_read + bgzf_mt thread
             pool) on the host cores, on a bounded sample of the same file
  extra.rans CRAM 3.1 rANS-Nx16 decode (BASELINE.json configs[2]) measured after the timed steps
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gb", type=float, default=10.0, help="uncompressed BAM gigabytes per GPU (10 = BASELINE configs[1])")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--quals", default="novaseq")
    ap.add_argument("--rans-slices", type=int, default=-2, help="CRAM slices for the rANS leg (0 = skip)")
    ap.add_argument("--tok3-blocks", type=int, default=9472, help="read-name blocks for the tok3 leg (0 = skip)")
    ap.add_argument("--cpu-sample-gb", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def ncu_traffic(kernel_tag):
    """dram__bytes_read+write of one launch from the committed ncu --set full capture (profiles/),
    with the algorithmic bytes of THAT captured launch, so the ratio can be applied honestly."""
    p = os.path.join(ROOT, "profiles", "r1_%s_ncu_summary.txt" % kernel_tag)
    if not os.path.exists(p):
        return None
    rd = wr = None
    for line in open(p):
        f = line.split()
        if len(f) >= 3 and f[0] == "dram__bytes_read.sum" and rd is None: rd = (float(f[1]), f[2])
        if len(f) >= 3 and f[0] == "dram__bytes_write.sum" and wr is None: wr = (float(f[1]), f[2])
    if not rd or not wr:
        return None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    return {"dram_bytes": rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1), "source": os.path.relpath(p, ROOT)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm: the unmodified reference built into oracle/_ref (bgzf_read over bgzf_mt)
# --------------------------------------------------------------------------------------------
def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libhts_ref.so")
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    r.hopen.restype = C.c_void_p
    r.hopen.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_size_t]
    r.bgzf_hopen.restype = C.c_void_p
    r.bgzf_hopen.argtypes = [C.c_void_p, C.c_char_p]
    r.bgzf_read.restype = C.c_ssize_t
    r.bgzf_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    r.bgzf_close.argtypes = [C.c_void_p]
    r.bgzf_mt.argtypes = [C.c_void_p, C.c_int, C.c_int]
    r.hfile_mem_steal_buffer.restype = C.c_void_p
    r.hfile_mem_steal_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    return r


def ref_decompress(r, img, threads):
    """bgzip -d equivalent in-process: mem: hFILE -> bgzf_hopen -> bgzf_mt(threads) -> bgzf_read loop.
    img: np.uint8 BGZF file image.  Returns (uncompressed bytes, seconds)."""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    mem = libc.malloc(img.size)                       # the mem: backend frees it at close
    C.memmove(mem, img.ctypes.data, img.size)
    t0 = time.perf_counter()
    hf = r.hopen(b"mem:", b"r:", mem, img.size)
    fp = r.bgzf_hopen(hf, b"r")
    if threads > 1:
        r.bgzf_mt(fp, threads, 256)
    buf = (C.c_uint8 * (16 << 20))()
    total = 0
    while True:
        n = r.bgzf_read(fp, buf, len(buf))
        if n <= 0:
            break
        total += n
    r.bgzf_close(fp)
    return total, time.perf_counter() - t0


def cpu_sample(corpus, sample_bytes):
    from tools import synth
    ulen, clen = corpus["ulen"], corpus["clen"]
    cu = np.cumsum(ulen.astype(np.int64))
    nb = int(np.searchsorted(cu, sample_bytes)) + 1
    nb = min(nb, len(ulen))
    cbytes = int(clen[:nb].astype(np.int64).sum())
    img = np.concatenate([corpus["comp"][:cbytes], np.frombuffer(synth.BGZF_EOF, dtype=np.uint8)])
    return img, int(cu[nb - 1]), cbytes, nb


def run_cpu_baseline(corpus, sample_gb, reps=2):
    r = ref_lib()
    cores = len(os.sched_getaffinity(0))
    if r is None:
        return None
    img, ubytes, cbytes, nb = cpu_sample(corpus, sample_gb * 1e9)
    best = None
    for _ in range(reps):
        total, sec = ref_decompress(r, img, cores)
        assert total == ubytes, (total, ubytes)
        best = sec if best is None else min(best, sec)
    return {"value": ubytes / best / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
            "sample": "first %d BGZF blocks (%.2f GB uncompressed) of the same file, bgzf_read over bgzf_mt(%d threads), zlib arm, best of %d"
                      % (nb, ubytes / 1e9, cores, reps)}


# --------------------------------------------------------------------------------------------
def make_corpus(args, rank, world=1):
    """Per-GPU corpus of args.gb uncompressed GB.  At N=1 every block is unique.  At N>1 the host
    cores are shared by N generators, so each rank makes gb/N unique GB (own seed) and repeats it N
    times at distinct addresses: per-GPU work stays fixed (weak scaling) and far larger than L2."""
    from tools import synth
    t0 = time.time()
    procs = max(1, len(os.sched_getaffinity(0)) // world)
    corpus = synth.bam_bgzf_corpus(args.gb * 1e9 / world, level=args.level, quals=args.quals, seed=42 + rank, procs=procs)
    if world > 1:
        corpus["comp"] = np.tile(corpus["comp"], world)
        corpus["clen"] = np.tile(corpus["clen"], world)
        corpus["ulen"] = np.tile(corpus["ulen"], world)
        corpus["n_reads"] *= world
    corpus["tile"] = world
    corpus["gen_s"] = time.time() - t0
    return corpus


def rans_leg(args, ctx, torch, dev):
    import htslib_b200 as H
    """CRAM 3.1 'normal'-profile shaped rANS blocks per slice of 10 000 x 150 bp reads: QS 1.5 MB
    32-way order-1, BF 15 kB 4-way order-1, CF/AP/NF/FN/BS 10 kB 4-way order-0 (SURVEY.md §8a').
    Streams are produced by the product's own GPU encoder."""
    from tools import synth
    if args.rans_slices == 0:
        return None
    rng = np.random.default_rng(4242)
    uniq = 16                                   # unique slices, tiled to rans_slices (distinct addresses)
    raws, orders = [], []
    for s_ in range(uniq):
        raws.append((synth.novaseq_quals(rng, 1_500_000) + 33).astype(np.uint8).tobytes()); orders.append(5)
        raws.append(rng.choice(np.array([99, 147, 83, 163], dtype=np.uint16), size=7500).astype("<u2").tobytes()); orders.append(1)
        for k in range(5):
            raws.append(np.clip(rng.normal(60, 25, size=10000), 0, 255).astype(np.uint8).tobytes()); orders.append(0)
    # input manufacture with the product's own GPU encoder (hgpu_rans_nx16_encode_batch_dev); its
    # streams are checked against the reference decoder in tests/test_gpu_rans_enc.py
    comps = ctx.rans_nx16_encode(raws, orders, torch.cuda.current_stream().cuda_stream)
    assert all(c is not None for c in comps)
    ulens = [len(r) for r in raws]
    if os.environ.get("RANS_ONLY_QS"):
        keep = [i for i in range(len(comps)) if ulens[i] > 100000]
        comps = [comps[i] for i in keep]; raws = [raws[i] for i in keep]; ulens = [ulens[i] for i in keep]
    wave = H.lib().hgpu_rans_nx16_wave_size(ctx.h)
    if args.rans_slices < 0:                      # -k: k full waves of quality blocks
        args.rans_slices = -args.rans_slices * wave
    per = len(comps) // uniq                    # streams per slice
    nsl = args.rans_slices                      # exactly this many slices (unique ones tiled round-robin)
    sel = np.concatenate([np.arange(per) + per * (s_ % uniq) for s_ in range(nsl)])
    in_len = np.array([len(comps[i]) for i in sel], dtype=np.uint32)
    out_len = np.array([ulens[i] for i in sel], dtype=np.uint32)
    # largest streams first so the persistent grid's tail is short
    order = np.argsort(-out_len.astype(np.int64), kind="stable")
    in_len, out_len = in_len[order], out_len[order]
    src_idx = sel[order]
    in_off = np.concatenate([[0], np.cumsum((in_len.astype(np.int64) + 15) // 16 * 16)[:-1]]).astype(np.uint64)
    out_off = np.concatenate([[0], np.cumsum((out_len.astype(np.int64) + 15) // 16 * 16)[:-1]]).astype(np.uint64)
    blob = np.zeros(int(in_off[-1]) + int(in_len[-1]) + 64, dtype=np.uint8)
    for o, si in zip(in_off, src_idx):
        c = comps[si]
        blob[int(o):int(o) + len(c)] = np.frombuffer(c, dtype=np.uint8)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
    d_in = torch.from_numpy(blob).to(dev)
    d_out = torch.empty(int(out_off[-1]) + int(out_len[-1]) + 64, dtype=torch.uint8, device=dev)
    d_io, d_il, d_oo, d_ol = t(in_off), t(in_len), t(out_off), t(out_len)
    n = len(in_len)
    d_got = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream   # run_ours made a real (non-default) stream current
    assert st != 0
    mx = int(out_len.max())
    times = []
    for it in range(2 + 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.rans_nx16_decode_dev(d_in, d_io, d_il, d_out, d_oo, d_ol, d_got, d_st, mx, st)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    assert int(d_st.abs().sum().item()) == 0, "rANS decode reported errors"
    # spot-check one QS block against the generator's input
    k = int(np.where(src_idx == 0)[0][0])
    got = d_out[int(out_off[k]):int(out_off[k]) + int(out_len[k])].cpu().numpy().tobytes()
    assert got == raws[0], "rANS spot check against the generator's input"
    ms = float(np.mean(times))
    U, Cb = int(out_len.astype(np.int64).sum()), int(in_len.astype(np.int64).sum())
    hbm, how = peaks()
    res = {"workload": "CRAM3.1 rANS-Nx16 decode, %d slices x (QS 1.5MB X32-O1 + BF 15kB O1 + 5x10kB O0), NovaSeq 4-bin quals, %d unique slices tiled"
                       % (nsl, uniq),
           "streams": n, "resident_streams_per_wave": int(wave), "uncompressed_GB": U / 1e9, "compressed_GB": Cb / 1e9, "ms": ms,
           "value": U / ms / 1e6, "unit": "GB/s (uncompressed)",
           "roofline": {"bound": "hbm", "achieved": (U + Cb) / ms / 1e6, "peak": hbm, "unit": "GB/s",
                        "frac": (U + Cb) / ms / 1e6 / hbm, "traffic": None, "peak_source": how}}
    try:                                                          # the reference's own decoder on the same streams, beside it
        res.update(ref_rans_rate(comps, ulens))
    except Exception as ex:
        res["cpu_reference_error"] = repr(ex)
    return res


def ref_rans_rate(comps, ulens, seconds=2.0):
    """rans_uncompress_to_4x16 of the unmodified reference (oracle/_ref) over the given streams: one core, then
    one thread per host core (ctypes releases the GIL), each for about `seconds`."""
    from concurrent.futures import ThreadPoolExecutor
    r = ref_lib()
    if r is None:
        return {}
    r.rans_uncompress_to_4x16.restype = C.c_void_p
    r.rans_uncompress_to_4x16.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]
    # the quality blocks carry ~97 % of the bytes; the 10-15 kB blocks would only measure Python's call overhead
    big = [i for i in range(len(comps)) if ulens[i] >= 100000] or list(range(len(comps)))
    comps = [comps[i] for i in big]; ulens = [ulens[i] for i in big]

    def work(k, until):
        out = (C.c_uint8 * max(1, max(ulens)))()
        done = 0
        i = k
        while time.perf_counter() < until:
            m = C.c_uint(ulens[i % len(comps)])
            if not r.rans_uncompress_to_4x16(comps[i % len(comps)], len(comps[i % len(comps)]), out, C.byref(m)):
                raise RuntimeError("reference decoder failed on stream %d" % (i % len(comps)))
            done += m.value
            i += 1
        return done
    t0 = time.perf_counter()
    one = work(0, t0 + seconds) / (time.perf_counter() - t0)
    cores = len(os.sched_getaffinity(0))
    allc = 0.0
    for _ in range(3):                                            # best of three: the first pass also warms the cores up
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            tot = sum(ex.map(lambda k: work(k, t0 + seconds), range(cores)))
        allc = max(allc, tot / (time.perf_counter() - t0))
    return {"cpu_reference_1core_GBps": one / 1e9, "cpu_reference_allcores_GBps": allc / 1e9, "cpu_reference_cores": cores,
            "cpu_reference_sample": "%d quality blocks, %.0f s per arm (all-core arm: best of 3), rans_uncompress_to_4x16 of oracle/_ref" % (len(comps), seconds)}


def run_ours(args):
    import torch
    import htslib_b200 as H
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ctx = H.Context(local)
    corpus = make_corpus(args, rank, world)
    comp, clen, ulen = corpus["comp"], corpus["clen"], corpus["ulen"]
    nb = len(clen)
    in_off = np.concatenate([[0], np.cumsum(clen.astype(np.int64))[:-1]]).astype(np.uint64)
    out_off = np.concatenate([[0], np.cumsum(ulen.astype(np.int64))[:-1]]).astype(np.uint64)
    U, Cb = int(ulen.astype(np.int64).sum()), int(clen.astype(np.int64).sum())
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
    d_in = torch.empty(Cb + 64, dtype=torch.uint8, device=dev)
    d_in[:Cb].copy_(torch.from_numpy(comp.copy()))
    d_out = torch.empty(U + 64, dtype=torch.uint8, device=dev)
    d_io, d_il, d_oo, d_cap = t(in_off), t(clen), t(out_off), t(ulen)
    d_len = torch.zeros(nb, dtype=torch.int32, device=dev)
    d_st = torch.zeros(nb, dtype=torch.int32, device=dev)
    # a real stream: handle 0 (legacy default) would make the library fall back to its own stream,
    # which torch.cuda.Event on the current stream cannot see
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    L = H.lib()

    def step():
        ctx.bgzf_inflate_dev(d_in, d_io, d_il, d_out, d_oo, d_cap, d_len, d_st, stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # correctness gate before timing: every block OK, lengths right, stream CRCs match the generator
    assert int(d_st.abs().sum().item()) == 0, "inflate reported errors"
    assert bool((d_len.cpu().numpy().astype(np.uint32) == ulen).all())
    import zlib
    # shards were compressed independently; check the first shard's CRC32 end to end
    k_shard = corpus["shard_blocks"][0]
    crc = zlib.crc32(d_out[:int(ulen[:k_shard].astype(np.int64).sum())].cpu().numpy().tobytes())
    assert crc == corpus["shard_crcs"][0], "first shard CRC mismatch"

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    launches0 = L.hgpu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    launches = L.hgpu_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps

    # ---- BAM record unpack over the inflated stream, still on the device (configs[3] front half) ----
    bam_extra = None
    if world == 1:
        try:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = ctx.bam_unpack_dev(d_out, U, d_oo, True, stream)
            torch.cuda.synchronize()
            first = time.perf_counter() - t0
            n_rec = r["n"]
            bad = int((r["status"] != 0).sum().item())
            written = int(r["data_off"][n_rec].item()) + 2 * int(r["seq_off"][n_rec].item()) + 48 * n_rec
            del r
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            r = ctx.bam_unpack_dev(d_out, U, d_oo, True, stream)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            hbm, how = peaks()
            bam_extra = {"workload": "record index + layout + unpack (bam1_core_t, data, ASCII SEQ, QUAL+33) of the inflated stream",
                         "records": n_rec, "bad_records": bad, "ms": ms, "includes": "two small D2H syncs for counts/sizes and output allocation",
                         "records_per_s": n_rec / ms * 1e3, "value": U / ms / 1e6, "unit": "GB/s (input stream)",
                         "roofline": {"bound": "hbm", "achieved": (U + written) / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                      "frac": (U + written) / ms / 1e6 / hbm, "traffic": None, "peak_source": how}}
            assert n_rec == corpus["n_reads"] and bad == 0
            # ---- write half of the BAM round trip (configs[3]): pack (bam_write1) + BGZF compress ----
            try:
                packed, pk_off, pk_st = ctx.bam_pack_dev(r["core"], r["data"], r["data_off"], n_rec, stream)
                torch.cuda.synchronize()
                same = bool(torch.equal(packed, d_out[:U]))
                nblk = (U + 0xff00 - 1) // 0xff00
                pin_off = torch.arange(nblk, dtype=torch.int64, device=dev) * 0xff00
                pin_len = torch.full((nblk,), 0xff00, dtype=torch.int32, device=dev)
                pin_len[-1] = U - (nblk - 1) * 0xff00
                c_out = torch.empty(nblk * 65536 + 64, dtype=torch.uint8, device=dev)
                c_off = torch.arange(nblk, dtype=torch.int64, device=dev) * 65536
                c_len = torch.zeros(nblk, dtype=torch.int32, device=dev); c_st = torch.zeros(nblk, dtype=torch.int32, device=dev)
                def squeeze():
                    H.check(L.hgpu_bgzf_compress_batch_dev(ctx.h, packed.data_ptr(), pin_off.data_ptr(), pin_len.data_ptr(), nblk, 6,
                                                           c_out.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), stream), "compress")
                squeeze(); torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(); squeeze(); ev1.record(); torch.cuda.synchronize()
                cms = ev0.elapsed_time(ev1)
                csz = int(c_len.sum().item())
                ev0.record()
                packed2, _, _ = ctx.bam_pack_dev(r["core"], r["data"], r["data_off"], n_rec, stream)
                ev1.record(); torch.cuda.synchronize()
                pms = ev0.elapsed_time(ev1)
                # re-inflate the first 2000 of our own blocks on the device and compare with the stream
                k = min(nblk, 2000)
                chk = torch.zeros(k * 65536, dtype=torch.uint8, device=dev)
                cap = torch.full((k,), 65536, dtype=torch.int32, device=dev)
                ol = torch.zeros(k, dtype=torch.int32, device=dev); st2 = torch.zeros(k, dtype=torch.int32, device=dev)
                ctx.bgzf_inflate_dev(c_out, c_off[:k].contiguous(), c_len[:k].contiguous(), chk, c_off[:k].contiguous(), cap, ol, st2, stream)
                torch.cuda.synchronize()
                ok = int(st2.abs().sum().item()) == 0 and bool((ol == pin_len[:k]).all().item())
                ok = ok and bool(torch.equal(chk.view(k, 65536)[0, :0xff00], packed[:0xff00]))
                bam_extra["write_half"] = {"pack_ms": pms, "pack_GBps": U / pms / 1e6, "pack_reproduces_stream": same,
                                           "deflate_ms": cms, "deflate_GBps": U / cms / 1e6, "deflate_level": ">=1 (LZ77 + fixed Huffman)",
                                           "compressed_bytes": csz, "ratio": csz / U, "zlib6_ratio": Cb / U,
                                           "size_vs_zlib6": csz / Cb, "reinflate_check": ok, "errors": int(c_st.abs().sum().item())}
                del packed, packed2, c_out, chk
            except Exception as ex:
                bam_extra["write_half"] = {"error": repr(ex)}
            del r
        except Exception as ex:
            bam_extra = {"error": repr(ex)}

    # ---- e2e: host buffers through the C-ABI file entry point ----
    e2e = None
    if not args.no_e2e:
        from tools import synth
        h_file = torch.empty(Cb + len(synth.BGZF_EOF), dtype=torch.uint8).pin_memory()
        h_file[:Cb].copy_(torch.from_numpy(comp.copy()))
        h_file[Cb:].copy_(torch.frombuffer(bytearray(synth.BGZF_EOF), dtype=torch.uint8))
        h_out = torch.empty(U, dtype=torch.uint8).pin_memory()
        fn, on = h_file.numpy(), h_out.numpy()
        del d_out
        torch.cuda.empty_cache()
        rc, n, bad = ctx.bgzf_inflate_file_host(fn, on)           # warm-up (allocates staging)
        assert rc == 0 and n == U, (rc, n, bad, H.last_error())
        assert zlib.crc32(on[:int(ulen[:k_shard].astype(np.int64).sum())].tobytes()) == corpus["shard_crcs"][0]
        reps = max(2, min(args.steps, 5))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            rc, n, bad = ctx.bgzf_inflate_file_host(fn, on)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / reps
        if world > 1:
            tt = torch.tensor([sec], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
        e2e = {"value": U * world / sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": Cb + len(synth.BGZF_EOF) + nb * 24,
               "d2h_bytes_per_step": U + nb * 8, "api": "hgpu_bgzf_inflate_file_host (pinned host buffers, 3-stream chunk pipeline)"}
        del h_file, h_out

    if rank != 0:
        return
    hbm, how = peaks()
    achieved = (U + Cb) / ms_step / 1e6
    out = {
        "metric": "BGZF inflate + CRAM rANS decode GB/s at 1/2/4/8 B200 vs reference CPU",
        "value": U * world / ms_step / 1e6, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BGZF inflate of %.1f GB synthetic 150bp BAM per GPU (30x-shaped, %s 4-bin quals), zlib level %d blocks, one warp per 64 KiB block"
                               % (U / 1e9, args.quals, args.level),
                   "blocks_per_gpu": nb, "uncompressed_bytes_per_gpu": U, "compressed_bytes_per_gpu": Cb,
                   "value_is": "uncompressed bytes / kernel time (inputs resident in HBM)",
                   "l2_policy": "inputs larger than L2 (%.1f GB in + %.1f GB out per step vs 126 MB)" % (Cb / 1e9, U / 1e9),
                   "data_gen_s": round(corpus["gen_s"], 1), "unique_data": "all blocks unique" if corpus["tile"] == 1 else
                                   "%.2f GB unique per GPU repeated %dx at distinct addresses (host cores shared by %d generators)"
                                   % (U / 1e9 / corpus["tile"], corpus["tile"], world)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": None, "kernel": "bgzf_inflate_kernel", "algorithmic_bytes_per_launch": U + Cb,
                     "peak_source": how,
                     "traffic_note": "ncu --set full was captured on the same kernel at --gb 1 (profiles/r1_inflate_ncu_summary.txt): "
                                     "dram read+write per launch there is in traffic_capture; a 10 GB launch was not captured (null above)",
                     "traffic_capture": ncu_traffic("inflate")},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e:
        out["e2e"] = e2e
    if bam_extra:
        out.setdefault("extra", {})["bam_unpack"] = bam_extra
    if world == 1 and not args.no_cpu_baseline:
        cb = run_cpu_baseline(corpus, args.cpu_sample_gb)
        if cb:
            out["cpu_baseline"] = cb
    if world == 1:
        try:
            rl = rans_leg(args, ctx, torch, dev)
            if rl:
                out.setdefault("extra", {})["rans_nx16_decode"] = rl
        except Exception as ex:                                   # the headline line must still print
            out.setdefault("extra", {})["rans_nx16_decode"] = {"error": repr(ex)}
        try:
            tl = tok3_leg(args, ctx)
            if tl:
                out.setdefault("extra", {})["tok3_decode"] = tl
        except Exception as ex:
            out.setdefault("extra", {})["tok3_decode"] = {"error": repr(ex)}
    print(json.dumps(out), flush=True)


def tok3_leg(args, ctx):
    """CRAM 3.1 read-name (tok3) blocks, 10 000 names per slice, through the host-buffer API
    (hgpu_tok3_decode_batch_host: framing walk on the host, rANS token streams + name rebuild on the
    device).  Inputs are fixtures written once by the reference encoder (tests/golden/make_tok3_slices.py)."""
    import gzip
    import htslib_b200 as H
    if args.tok3_blocks <= 0:
        return None
    d = os.path.join(ROOT, "tests", "golden", "tok3_slices")
    uniq = [open(os.path.join(d, "slice%d.tok3" % k), "rb").read() for k in range(8)]
    want = [gzip.open(os.path.join(d, "slice%d.names.gz" % k), "rb").read() for k in range(8)]
    Lh = H.lib()
    n = args.tok3_blocks
    comps = [uniq[k % 8] for k in range(n)]
    in_len = np.array([len(c) for c in comps], dtype=np.uint32)
    in_off = np.concatenate([[0], np.cumsum(in_len.astype(np.uint64))[:-1]]).astype(np.uint64)
    blob = np.frombuffer(b"".join(comps) + b"\0" * 8, dtype=np.uint8)
    caps = np.array([Lh.hgpu_tok3_out_bound(c, len(c)) for c in comps], dtype=np.uint32)
    out_off = np.concatenate([[0], np.cumsum(caps.astype(np.uint64))[:-1]]).astype(np.uint64)
    out = np.zeros(int(caps.astype(np.uint64).sum()) + 8, dtype=np.uint8)
    got = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Lh.hgpu_tok3_last_ms.argtypes = [C.c_void_p]
    wall, dev_ms = [], []
    for it in range(1 + 3):
        t0 = time.perf_counter()
        rc = Lh.hgpu_tok3_decode_batch_host(ctx.h, p(blob), p(in_off), p(in_len), n, p(out), p(out_off), p(caps), p(got), p(st))
        t1 = time.perf_counter()
        assert rc == 0, H.last_error()
        ms2 = (C.c_float * 2)()
        Lh.hgpu_tok3_last_ms(ms2)
        if it >= 1:
            wall.append(t1 - t0); dev_ms.append((ms2[0], ms2[1]))
    assert int(np.abs(st).sum()) == 0, "tok3 decode reported errors"
    for k in range(8):
        assert out[int(out_off[k]):int(out_off[k]) + int(got[k])].tobytes() == want[k], "tok3 output differs from the names the fixture was made from"
    U = int(got.astype(np.int64).sum()); Cb = int(in_len.astype(np.int64).sum())
    ent = float(np.mean([a for a, b in dev_ms])); nam = float(np.mean([b for a, b in dev_ms]))
    res = {"workload": "%d tok3 blocks x 10 000 Illumina names (8 unique reference-encoded fixtures tiled), host buffers" % n,
           "blocks": n, "names": n * 10000, "uncompressed_GB": U / 1e9, "compressed_GB": Cb / 1e9,
           "entropy_ms": ent, "rebuild_ms": nam, "device_GBps": U / (ent + nam) / 1e6,
           "names_per_s_device": n * 10000 / ((ent + nam) / 1e3),
           "e2e_ms": float(np.mean(wall)) * 1e3, "e2e_GBps": U / float(np.mean(wall)) / 1e9,
           "unit": "GB/s (uncompressed names)"}
    r = ref_lib()
    if r is not None:                                             # the reference's own decoder, one core, the 8 unique blocks
        r.tok3_decode_names.restype = C.c_void_p
        libc = C.CDLL(None)
        t0 = time.perf_counter(); tot = 0
        for rep in range(3):
            for c in uniq:
                m = C.c_uint(0)
                q = r.tok3_decode_names(c, C.c_uint(len(c)), C.byref(m))
                tot += m.value
                libc.free(C.c_void_p(q))
        res["cpu_reference_1core_GBps"] = tot / (time.perf_counter() - t0) / 1e9
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = ref_lib()
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhts_ref.so was not built (needs /root/reference once)"}))
        return
    corpus = make_corpus(args, 0)
    cores = len(os.sched_getaffinity(0))
    img, ubytes, cbytes, nb = cpu_sample(corpus, args.cpu_sample_gb * 1e9)
    for _ in range(args.warmup):
        ref_decompress(r, img, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        total, _s = ref_decompress(r, img, cores)
        assert total == ubytes
    sec = (time.perf_counter() - t0) / args.steps
    v = ubytes / sec / 1e9
    sample = "first %d BGZF blocks (%.2f GB uncompressed) of the same synthetic BAM per step; unmodified htslib bgzf_read over bgzf_mt(%d), zlib arm (no libdeflate here)" % (nb, ubytes / 1e9, cores)
    print(json.dumps({
        "impl": "reference",
        "metric": "BGZF inflate + CRAM rANS decode GB/s at 1/2/4/8 B200 vs reference CPU",
        "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BGZF inflate of synthetic 150bp BAM (same generator as the GPU arm), zlib level %d blocks" % args.level,
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()
                dist.destroy_process_group()
