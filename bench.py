#!/usr/bin/env python
"""Benchmark of the .gst -> DXT1 decode path (BASELINE.json metric: decoded GTexel/s and
compressed GB/s per B200 and at 2/4/8 GPUs, against the HBM roofline).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU decode

A step is one LoadCompressedDXTs-style call over the workload (default: BASELINE.json configs[3],
a batch of 1024 2048x2048 textures; --config picks another).  With N GPUs the batch is sharded,
image i to rank i mod N (--scaling weak: every rank decodes the whole batch instead).  Inputs
are reference-encoded .gst streams of seeded synthetic images (`--distinct` different images,
tiled to the batch size); they are resident in HBM when the timed region starts.  The `e2e`
figure runs the same workload through gst_decompress_host_batch with pinned HOST buffers:
host packing, H2D, decode and D2H of every DXT1 block inside the timed region (the shape of
GenTC::DecompressDXT, and of what the CPU reference arm produces).  `e2e_resident` is the
photos_sf shape: the same host buffers through gst_load_host_batch, textures left in device
memory, one 16-byte probe per image read back.

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# BASELINE.json configs (index = position in "configs"): (width, height, images per GPU, name)
CONFIGS = {
    0: (512, 512, 1, "configs[0]: single 512x512 texture"),
    1: (512, 512, 128, "configs[1]: batch of 128 512x512 textures"),
    2: (4096, 4096, 1, "configs[2]: single 4096x4096 texture"),
    3: (2048, 2048, 1024, "configs[3]: batch of 1024 2048x2048 textures"),
    4: (1920, 1024, 600, "configs[4]: 600 frames 1920x1024"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", type=int, default=3, help="index into BASELINE.json configs")
    ap.add_argument("--images", type=int, default=0, help="override images per GPU")
    ap.add_argument("--distinct", type=int, default=32, help="distinct encoded images tiled to the batch")
    ap.add_argument("--page", type=int, default=32, help="images per page on the e2e path")
    ap.add_argument("--depth", type=int, default=4, help="groups of frames in flight of the configs[4] frame streamer")
    ap.add_argument("--group", type=int, default=0, help="frames per decode call of gst_streamer_play (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong",
                    help="strong (default): the configured batch is the job, image / frame i goes to rank i mod N; "
                         "weak: every GPU decodes the whole configured batch")
    ap.add_argument("--upload", choices=["auto", "staged", "direct"], default="auto",
                    help="e2e legs: pack pages into pinned staging (staged) or DMA every pinned .gst file from where it "
                         "lies (direct, gst_ctx_set_direct_upload).  auto: on one GPU one untimed step of each and the faster "
                         "one is kept (staged wins where the link is the limit, direct where host memory is), direct "
                         "from 2 GPUs on (the host memory system is the limit there)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU baseline sample")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def load_streams(cfg_id, width, height, distinct, rank, world):
    """`distinct` reference-encoded streams of seeded synthetic images (seed = config*10000+i,
    SURVEY.md section 8d).  Rank 0 encodes what the cache lacks; the others wait."""
    import gst_fixtures as fx
    seeds = [cfg_id * 10000 + i for i in range(distinct)]
    if cfg_id == 4:
        # the motion sequence: one base image translated by 2 px per frame; `distinct` consecutive frames, played
        # forwards and backwards to fill the 600 (the encoder is ~1.2 s per frame and core)
        make = lambda: fx.encode_motion(width, height, cfg_id * 10000, distinct)
    else:
        make = lambda: fx.encode_images(width, height, seeds)
    if world > 1:
        import torch.distributed as dist
        if rank == 0:
            make()
        dist.barrier()
    streams = make()
    return [g for g, _ in streams], [d for _, d in streams]


def cpu_decode_rate(files, sample, threads):
    """Reference CPU decode (oracle/_ref: ans/decode.cpp + codec/wavelet.cpp linked unmodified,
    scan/assembly restated) of `sample` images on `threads` host threads.  Returns
    (seconds, kind).  Falls back to the plain-C port when the reference library is absent."""
    import gst_fixtures as fx
    L = fx.ref()
    hdr = fx.header_of(files[0])
    per = hdr["width"] * hdr["height"] // 2
    n = sample
    outs = [np.empty(per, dtype=np.uint8) for _ in range(min(n, 4 * threads))]
    if L is not None:
        ptrs = (C.c_void_p * n)(*[files[i % len(files)].ctypes.data for i in range(n)])
        lens = (C.c_size_t * n)(*[files[i % len(files)].size for i in range(n)])
        optr = (C.c_void_p * n)(*[outs[i % len(outs)].ctypes.data for i in range(n)])
        t0 = time.perf_counter()
        failed = L.gstref_decode_batch(ptrs, lens, n, optr, threads)
        dt = time.perf_counter() - t0
        assert failed == 0, f"{failed} reference decodes failed"
        return dt, "reference"
    O = fx.oracle()
    t0 = time.perf_counter()
    work = list(range(n))
    lock = threading.Lock()

    def run():
        out = np.empty(per, dtype=np.uint8)
        while True:
            with lock:
                if not work:
                    return
                i = work.pop()
            f = files[i % len(files)]
            O.gsto_decode(f.ctypes.data, f.size, 0, out.ctypes.data, None, None, None)  # ctypes drops the GIL

    ts = [threading.Thread(target=run) for _ in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return time.perf_counter() - t0, "port"


def run_reference(args, rank, world, cfg):
    width, height, images, name = cfg
    if rank != 0:
        return
    files, _ = load_streams(args.config, width, height, min(args.distinct, max(images, 1)), 0, 1)
    threads = os.cpu_count() or 1
    # a bounded sample per step, large enough that every host thread gets several images (8 per thread: the rate of
    # the 16-thread pool is 0.91 GTexel/s on 2 images per thread, 1.16 on 8)
    sample = args.cpu_sample or max(1, min(images, 8 * threads))
    for _ in range(args.warmup):
        cpu_decode_rate(files, sample, threads)
    total, kind = 0.0, "reference"
    for _ in range(args.steps):
        dt, kind = cpu_decode_rate(files, sample, threads)
        total += dt
    texels = float(width) * height * sample * args.steps
    cmp_bytes = float(sum(files[i % len(files)].size - 28 for i in range(sample))) * args.steps
    val = texels / total / 1e9
    line = {
        "impl": "reference", "metric": "decoded GTexel/s (.gst -> DXT1)", "value": val, "unit": "GTexel/s",
        "compressed_gb_s": cmp_bytes / total / 1e9, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8/i32", "data": "synthetic",
        "config": {"workload": name, "width": width, "height": height,
                   "note": "CPU decode of a bounded sample per step, host memory only"},
        "cpu_baseline": {"value": val, "unit": "GTexel/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} images of {width}x{height} per step, one image per task"},
        "e2e": {"value": val, "unit": "GTexel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank, local_rank, world = dist_env()
    cfg = CONFIGS[args.config]
    if args.images:
        cfg = (cfg[0], cfg[1], args.images, cfg[3] + f" (images overridden to {args.images})")
    if args.impl == "reference":
        run_reference(args, rank, world, cfg)
        return
    width, height, images_total, name = cfg

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    import gst_b200
    import gst_fixtures as fx
    from gst_b200.capi import check, lib

    from gst_b200.shard import reduce_job, shard_indices

    dec = gst_b200.Decoder(local_rank)
    direct = args.upload == "direct" or (args.upload == "auto" and world >= 2)
    if direct:
        check(lib().gst_ctx_set_direct_upload(dec.ctx, 1))
    distinct = min(args.distinct, images_total)
    files, goldens = load_streams(args.config, width, height, distinct, rank, world)
    # The job is the configured batch (BASELINE.json: "sharded across 1/2/4/8 B200", frames "streamed across 8 B200"):
    # image / frame i goes to rank i mod N, no data-path collective.  --scaling weak makes every
    # rank decode the whole configured batch instead (N independent replicas).
    strong = args.scaling == "strong"
    if args.config == 4 and distinct > 1:
        period = 2 * distinct - 2  # frame f of the sequence: 0, 1, .., distinct-1, distinct-2, .., 1, 0, 1, ..
        pick = lambda i: (i % period) if (i % period) < distinct else period - (i % period)
    else:
        pick = lambda i: i % distinct
    order = [pick(i) for i in (shard_indices(images_total, rank, world) if strong else range(images_total))]
    if not strong:
        order = [(j + rank) % distinct for j in order]  # replicas start at different images
    images = len(order)
    assert images > 0, f"rank {rank} has no image: {images_total} images over {world} ranks"
    batch = [files[j] for j in order]
    N = fx.header_of(files[0])["width"] * fx.header_of(files[0])["height"] // 16
    stream = dec.GetDefaultCommandQueue()

    class Resident:
        """One batch packed in the LoadCompressedDXTs layout, resident in HBM, and the call that decodes it."""

        def __init__(self, blobs):
            self.n = len(blobs)
            self.packed, self.hdrs = gst_b200.pack_batch(blobs)
            self.d_cmp, self.d_out = dec.malloc(self.packed.size), dec.malloc(8 * N * self.n)
            dec.upload(self.d_cmp, self.packed)
            self.harr = (gst_b200.capi.gst_header * self.n)(*[h.to_c() for h in self.hdrs])
            self.launches = lib().gst_launches_for_batch(self.harr, self.n)

        def step(self):
            check(lib().gst_load_dxt_batch(dec.ctx, self.harr, self.n, stream, self.d_cmp.ptr, self.d_cmp.nbytes,
                                           self.d_out.ptr, None, 0, None))

        def free(self):
            self.d_cmp.free()
            self.d_out.free()

    res = Resident(batch)
    step = res.step

    # ---- parity before timing: every distinct image against the CPU oracle ----------------
    dec.memset(res.d_out, 0xEE)
    step()
    dec.sync(stream)
    checked = 0
    for pos in range(min(distinct, images)):
        got = dec.download(res.d_out, 8 * N, offset=pos * 8 * N)
        j = order[pos]
        if pos < 2:
            want = fx.oracle_decode(files[j], taps=False)["out"]
            assert np.array_equal(got, want), f"image {pos}: CUDA output differs from the CPU oracle"
        assert fx.matches_golden(got, goldens[j]), f"image {pos}: CUDA output differs from the encoder's PhysicalBlocks()"
        checked += 1
    last = dec.download(res.d_out, 8 * N, offset=(images - 1) * 8 * N)
    assert fx.matches_golden(last, goldens[order[-1]]), "last image of the batch differs"

    texels_rank = float(width) * height * images
    cmp_rank = float(sum(f.size - 28 for f in batch))
    alg_bytes_rank = cmp_rank + 8.0 * N * images  # (file - 28) + W*H/2 per image
    # L2: when one step's inputs + outputs do not exceed twice the 126 MB L2, a 256 MB buffer is overwritten between
    # the timed steps (outside the per-step events), so that no step finds its inputs in the cache
    flush = alg_bytes_rank <= 2 * 126e6
    d_flush = dec.malloc(256 << 20) if flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dec.sync()

    def timed(run_step, steps):
        """-> (total device ms of `steps` steps, sorted per-step ms); CUDA events on the launching stream."""
        barrier()
        if not flush:
            marks = [dec.record(stream)]
            for _ in range(steps):
                run_step()
                marks.append(dec.record(stream))  # one event per step: median and best step
            marks[-1].wait()
            per = [marks[i].elapsed_ms(marks[i + 1]) for i in range(steps)]
            total = marks[0].elapsed_ms(marks[-1])
        else:
            pairs = []
            for _ in range(steps):
                dec.memset(d_flush, 0x5A)
                a = dec.record(stream)
                run_step()
                pairs.append((a, dec.record(stream)))
            pairs[-1][1].wait()
            per = [a.elapsed_ms(b) for a, b in pairs]
            total = float(sum(per))
        barrier()
        return total, sorted(per)

    sampler = ClockSampler(local_rank)
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    dec.sync(stream)

    # ---- timed region: K steps, the production path (no per-kernel events) ----------------------
    ms_total, per_step = timed(step, args.steps)
    # whole job: max over ranks of the device time, sum over ranks of the units processed
    ms_total, (texels_job, cmp_job, h2d_job, d2h_job, alg_job) = reduce_job(
        ms_total, [texels_rank, cmp_rank, float(res.packed.size), 8.0 * N * images, alg_bytes_rank])
    ms_step = ms_total / args.steps

    # ---- per-kernel times: a separate short pass with events around every kernel -----------------
    dec.profile(True)
    prof_steps = min(args.steps, 5)
    timed(step, prof_steps)
    dec.profile(False)
    kernel_ms, calls = dec.profile_read()
    per_call = {k: v / max(calls, 1) for k, v in kernel_ms.items()}

    # ---- the other scaling mode, for the record (N > 1 only) --------------------------------------
    other = None
    if world > 1:
        o_order = ([(pick(i) + rank) % distinct for i in range(images_total)] if strong
                   else [pick(i) for i in shard_indices(images_total, rank, world)])
        if o_order:
            o_res = Resident([files[j] for j in o_order])
            for _ in range(3):
                o_res.step()
            o_ms, _ = timed(o_res.step, max(3, args.steps // 2))
            o_ms, (o_tex,) = reduce_job(o_ms, [float(width) * height * len(o_order)])
            o_step = o_ms / max(3, args.steps // 2)
            other = {"scaling": "weak" if strong else "strong", "images_per_gpu": len(o_order), "ms_per_step": o_step,
                     "value": o_tex / (o_step * 1e-3) / 1e9, "unit": "GTexel/s"}
            o_res.free()

    # ---- end to end: host .gst buffers -> host DXT1 blocks ---------------------------------
    e2e = None
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 10))
    if not args.no_e2e:
        pin_files = []
        for f in files:
            pb = dec.pinned(f.size)
            pb.array[:] = f
            pin_files.append(pb)
        pin_out = dec.pinned(8 * N * images)
        ptrs = (C.c_void_p * images)(*[pin_files[j].ptr for j in order])
        lens = (C.c_size_t * images)(*[files[j].size for j in order])

        streamer = gst_b200.FrameStreamer(dec, width, height, depth=args.depth) if args.config == 4 else None

        def e2e_step():
            if streamer is None:
                check(lib().gst_decompress_host_batch(dec.ctx, ptrs, lens, images, args.page, 0, pin_out.ptr, pin_out.nbytes))
                return
            # configs[4]: the demo player loop (demo/demo.cpp:145-243, 504-600) with `depth` groups of frames in flight:
            # every frame is copied into its slot's pinned staging, a group goes up in one DMA, is decoded in one call
            # and copied back to the host on the slot's stream
            streamer.play(ptrs, lens, images, host_out=pin_out.ptr, direct=False, group=args.group)

        pin_out.array[:] = 0xEE
        e2e_step()  # warm-up: grows the staging buffers
        upload_probe = None
        if streamer is None and args.upload == "auto" and world == 1:
            # Which upload policy wins depends on the box's host side (staged packing is 4 % ahead
            # where the link is the limit, direct DMA of the pinned files where host memory is): one untimed step of
            # each, after a warm-up step of the other policy, and the context keeps the faster one.
            def one_step_ms():
                t = time.perf_counter()
                e2e_step()
                return (time.perf_counter() - t) * 1e3
            staged_ms = one_step_ms()
            check(lib().gst_ctx_set_direct_upload(dec.ctx, 1))
            e2e_step()
            direct_ms = one_step_ms()
            direct = direct_ms < staged_ms
            check(lib().gst_ctx_set_direct_upload(dec.ctx, 1 if direct else 0))
            upload_probe = {"staged_ms": staged_ms, "direct_ms": direct_ms}
            pin_out.array[:] = 0xEE
            e2e_step()
        for pos in sorted({0, images // 3, images // 2, (2 * images) // 3, images - 1}):  # frames / images across the step
            assert fx.matches_golden(pin_out.array[pos * 8 * N:(pos + 1) * 8 * N], goldens[order[pos]]), f"e2e output {pos} differs"
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        dt, _ = reduce_job(dt, [])
        e2e = {"value": texels_job * e2e_steps / dt / 1e9, "unit": "GTexel/s",
               "h2d_bytes_per_step": int(h2d_job), "d2h_bytes_per_step": int(d2h_job),
               "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps, "page_images": args.page,
               "compressed_gb_s": cmp_job * e2e_steps / dt / 1e9,
               "api": "gst_streamer_play (staged upload, decode and read-back per group of frames on the slot's stream)"
                      if streamer is not None else "gst_decompress_host_batch, " + ("direct" if direct else "staged") + " upload",
               "timing": "host wall clock around blocking calls (copies + kernels inside), max over ranks"}
        if upload_probe:
            e2e["upload_probe"] = upload_probe
        if streamer is not None:
            e2e["frames_per_s"] = images_total * e2e_steps / dt
            streamer.close()
    # ---- photos_sf shape: host .gst buffers -> textures resident in device memory ---------------
    e2e_res = None
    if not args.no_e2e:
        probe = dec.pinned(16 * images)

        def res_step():
            check(lib().gst_load_host_batch(dec.ctx, ptrs, lens, images, args.page, 0, res.d_out.ptr, res.d_out.nbytes))
            # the step's result read: first and last block of every image (strided D2H)
            dec.download_2d(probe, res.d_out, 8, images, src_pitch=8 * N, dst_pitch=16)
            dec.download_2d(probe, res.d_out, 8, images, src_pitch=8 * N, dst_pitch=16, src_offset=8 * N - 8, dst_offset=8)

        dec.memset(res.d_out, 0xEE)
        res_step()
        dec.sync()
        got0 = dec.download(res.d_out, 8 * N, offset=0)
        assert fx.matches_golden(got0, goldens[order[0]]), "resident e2e output differs"
        gotl = dec.download(res.d_out, 8 * N, offset=(images - 1) * 8 * N)
        assert fx.matches_golden(gotl, goldens[order[-1]]), "resident e2e output differs"
        assert np.array_equal(probe.array[:8], got0[:8]) and np.array_equal(probe.array[-8:], gotl[-8:])
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res_step()
        barrier()
        dt = time.perf_counter() - t0
        dt, _ = reduce_job(dt, [])
        e2e_res = {"value": texels_job * e2e_steps / dt / 1e9, "unit": "GTexel/s",
                   "h2d_bytes_per_step": int(h2d_job), "d2h_bytes_per_step": int(16 * images_total if strong else 16 * images * world),
                   "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps, "page_images": args.page,
                   "note": "host .gst -> DXT1 resident in device memory (LoadCompressedDXTs into a device buffer); "
                           "16-byte probe per image read back"}
    clocks = sampler.stop()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        per_image_s = 0.11 * (width * height) / (2048.0 * 2048.0)  # survey probe, one core
        sample = args.cpu_sample or int(max(1, min(images, 15.0 / max(per_image_s, 1e-4))))
        dt, kind = cpu_decode_rate(files, sample, threads)
        # the same decode on ONE host thread as well (a few images)
        one = int(max(1, min(sample, 3.0 / max(per_image_s, 1e-4))))
        dt1, _ = cpu_decode_rate(files, one, 1)
        cpu = {"value": float(width) * height * sample / dt / 1e9, "unit": "GTexel/s", "cores": threads, "kind": kind,
               "sample": f"{sample} images of {width}x{height} ({distinct} distinct), one image per task, {dt:.2f} s wall",
               "single_thread": {"value": float(width) * height * one / dt1 / 1e9, "unit": "GTexel/s", "cores": 1,
                                 "sample": f"{one} images, {dt1:.2f} s wall"}}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        # the roofline figure is the STEP: algorithmic bytes (compressed bytes read once + DXT1 bytes
        # written once, intermediates not counted) of this rank over its step time, against the measured HBM copy
        # bandwidth.  Per kernel: rans_streams reads every compressed stream and frequency table, wavelet_assemble
        # writes every DXT1 block.
        alg = {"rans_streams": cmp_rank, "wavelet_assemble": 8.0 * N * images}
        kernels = {}
        for k, v in per_call.items():
            kernels[k] = {"ms": v}
            if k in alg and v > 0:
                kernels[k].update(alg_bytes=alg[k], achieved=alg[k] / (v * 1e-3) / 1e9, frac=alg[k] / (v * 1e-3) / 1e9 / peak)
        # DRAM bytes per step from the committed `ncu --set full` capture (profiles/traffic.json: dram__bytes_read +
        # dram__bytes_write per image of a given size, per kernel)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            per_img = tj.get(f"{width}x{height}")
            if per_img:
                traffic = float(sum(v for k, v in per_img.items() if isinstance(v, (int, float)))) * images
        except Exception:
            pass
        step_rank_ms = per_step[len(per_step) // 2]
        step_gbs = alg_job / (ms_step * 1e-3) / 1e9 / world  # per GPU
        line = {
            "metric": "decoded GTexel/s (.gst -> DXT1)", "value": texels_job / (ms_step * 1e-3) / 1e9,
            "unit": "GTexel/s", "compressed_gb_s": cmp_job / (ms_step * 1e-3) / 1e9,
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
            "ms_per_step_median_rank0": step_rank_ms, "ms_per_step_best_rank0": per_step[0],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8/i32 (integer only)",
            "data": "synthetic",
            "config": {"workload": name, "width": width, "height": height, "images": images_total if strong else images * world,
                       "images_per_gpu": images, "distinct_images": distinct,
                       "bits_per_texel": 8.0 * cmp_rank / texels_rank,
                       "sharding": ("image i -> rank i mod N" if strong else "every rank decodes the whole batch") +
                                   ", no data-path collective",
                       "l2": "a 256 MB buffer is overwritten between timed steps (working set fits the 126 MB L2)" if flush
                             else "inputs + outputs per step exceed twice the 126 MB L2, no flush",
                       "parity": f"{checked} distinct images + last checked bit-exact before timing"},
            "roofline": {"bound": "hbm", "kernel": "step (build_tables + rans_streams + wavelet_assemble)",
                         "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                         "peak_source": peak_src, "traffic": traffic, "alg_bytes_per_step_per_gpu": alg_job / world,
                         "kernels": kernels},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_resident": e2e_res, "other_scaling": other,
            "gpu_launches": res.launches * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    dec.close()


if __name__ == "__main__":
    main()
