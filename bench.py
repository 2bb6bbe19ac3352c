#!/usr/bin/env python3
"""bench.py -- metaseg images/s on synthetic 2048x2048 DAPI (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the CPU path on the box's host cores)

A step = one pass of the whole hot path (pre-process -> tile -> U-Net -> stitch/quantise/argmax ->
meta_inference -> ecDNA count) over `--images-per-step` distinct synthetic images per GPU.
  value : images/s with the raw images already resident in HBM (CUDA events, max over ranks)
  e2e   : images/s through ecseg_segment_image_host with pinned HOST buffers (H2D image copy and
          D2H label-map + count copy inside the timed region)
  roofline     : the U-Net stage (tcgen05 implicit-GEMM kernels), algorithmic FLOPs / CUDA-event time
  cpu_baseline : the CPU oracle port timed on this box's host cores on a bounded sample (rank 0, N=1)
Images shard across GPUs with no collective on the data path (weak scaling); torch.distributed is
used only for the start/stop barrier and the max-over-ranks of the device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 2048
TILES_PER_IMAGE = 100
METRIC = "metaseg images/s (2048x2048 DAPI -> seg + ecDNA count)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1420.3), d.get("hbm_gbs", 6465.2), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "power_w_max": None}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        out["power_w_max"] = max(float(r[3]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any("Active" == r[4 + i].strip() for r in rows)]
        out["samples"] = len(rows)
        return out


# --------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port; the reference's TF/skimage stack is not installable)
# --------------------------------------------------------------------------------------------
CPU_STAGES = ("preprocess_tile", "unet", "stitch_quantise_argmax", "meta_inference", "count_cc")


def cpu_step(oracle_net, img, n_tiles: int = TILES_PER_IMAGE):
    """The reference's per-image path (src/utils.py:109-120 + src/metaseg.py:46) on the host cores for ONE 2048x2048
    image, every stage at full size: all 100 tiles go through the U-Net, nothing is extrapolated.  (`n_tiles` < 100
    is only for the warm-up call.)  Returns (seconds, per-stage seconds)."""
    from oracle import metaseg_oracle as mo
    t = [time.perf_counter()]
    pre = mo.meta_preprocess(img)
    pos, tiles = mo.im2patches_overlap(pre[..., None])
    t.append(time.perf_counter())
    probs = oracle_net.predict_on_batch(tiles[:n_tiles])
    if n_tiles < len(tiles):
        probs = np.concatenate([probs, np.broadcast_to(probs[:1], (len(tiles) - n_tiles,) + probs.shape[1:])])
    t.append(time.perf_counter())
    lab = mo.quantise_argmax(mo.patches2im_overlap(probs, pos))
    t.append(time.perf_counter())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lab = mo.meta_inference(lab)
        t.append(time.perf_counter())
        mo.count_cc(lab == 3)
    t.append(time.perf_counter())
    return t[-1] - t[0], [b - a for a, b in zip(t[:-1], t[1:])]


def make_oracle_net():
    import torch
    from ecseg_b200 import weights as wmod
    from oracle.unet_oracle import UNetOracle
    torch.set_num_threads(os.cpu_count() or 1)
    return UNetOracle(wmod.make_weights(0), batch=4)


def workload_config(images_per_step: int, contexts: int, pool: int):
    """`config` of the bench line -- the same dict on both arms (the reference arm describes its bounded sample of
    this workload in `cpu_baseline.sample`)."""
    return {"workload": f"{images_per_step} x 2048x2048 synthetic DAPI images per GPU per step (100 tiles each), "
                        "whole path: preprocess+tile+U-Net+stitch+meta_inference+count",
            "weights": "random-init seed 0 of the metaseg.h5 architecture (BN folded)",
            "contexts_per_gpu": contexts,
            "l2": "each image moves ~6 GB of activations through HBM, far beyond the 126 MB L2; "
                  f"{pool} distinct images cycled",
            "collectives": "none on the data path; torch.distributed (NCCL) only for the start barrier and the "
                           "max-over-ranks of the timer"}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port: reference control flow,
    TensorFlow -> torch-CPU oneDNN fp32, skimage -> scipy; neither library is installable offline) on this box's host
    cores.  MEASURED, not extrapolated: every step is one whole 2048x2048 image, all 100 tiles through the U-Net.
    Under torchrun only rank 0 works; the other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ecseg_b200 import synth
    net = make_oracle_net()
    imgs = [synth.synth_dapi(1000 + s, H, W) for s in range(2)]
    for i in range(args.warmup):
        cpu_step(net, imgs[i % 2])
    t0 = time.perf_counter()
    res = [cpu_step(net, imgs[i % 2]) for i in range(args.steps)]
    wall = time.perf_counter() - t0
    sec = wall / args.steps
    stages = np.mean([r[1] for r in res], axis=0)
    val = 1.0 / sec
    n_ctx = max(1, args.contexts)
    sample = ("each step = ONE whole 2048x2048 image of the workload (1 of the GPU arm's "
              f"{args.images_per_step} images per step), all 100 tiles through the U-Net, every stage at full size; "
              "nothing extrapolated; rank 0 only (one CPU process whatever --gpus says)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.images_per_step, n_ctx, max(args.images_per_step, 8)),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_stage_s_per_image": {k: float(v) for k, v in zip(CPU_STAGES, stages)},
        "timed_region_s": wall,
        "note": "reference control flow restated (oracle/); TensorFlow -> torch-CPU oneDNN, skimage -> scipy: neither is "
                "installable offline.  One CPU process on rank 0 at every N: compare per-N GPU values with it as context only.",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# with artefacts: the whole `make metaseg` loop (SURVEY section 8d, config 3 "with artefacts")
# --------------------------------------------------------------------------------------------
def run_artifacts(args, weights, local, rank=0, barrier=None):
    """TIFF files in -> dapi/<name>.tif + labels/<stem>.png + labels/<stem>.npy out through ecseg_b200.pipeline
    (decode, GPU path with on-device PNG deflate / int64 widening, file writes overlapped), on a bounded sample,
    next to the reference's writer calls timed on the host for the same label maps."""
    import shutil

    import cv2

    from ecseg_b200 import spec, synth
    from ecseg_b200.pipeline import FilesPipeline, output_paths

    need = args.artifact_images * (H * W * 10 + (1 << 20)) + (1 << 30)
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(cand) and os.access(cand, os.W_OK) and shutil.disk_usage(cand).free > need:
            base = cand
            break
    d = tempfile.mkdtemp(prefix="ecseg_art_", dir=base)
    try:
        os.mkdir(os.path.join(d, "dapi"))
        os.mkdir(os.path.join(d, "labels"))
        n = args.artifact_images
        distinct = [synth.synth_dapi(5000 + 100 * rank + s, H, W) for s in range(8)]
        paths = []
        for i in range(n):
            p = os.path.join(d, f"img{i:04d}.tif")
            cv2.imwrite(p, distinct[i % 8], [cv2.IMWRITE_TIFF_COMPRESSION, 1])
            paths.append(p)
        pipe = FilesPipeline(weights, args.precision, H, W, device=local, n_ctx=max(1, args.contexts),
                             n_readers=args.readers, n_writers=args.writers, max_bytes_per_px=1)
        try:
            pipe.run(paths[:8])                       # warm-up: page in buffers, create output files once
            if barrier:
                barrier()                             # all ranks start their timed run together (shared host cores)
            rows = pipe.run(paths)
            st = dict(pipe.stats)
        finally:
            pipe.close()
        png_bytes = float(np.mean([os.path.getsize(output_paths(p)[1]) for p in paths]))
        # the reference's three writer calls + its read, timed on this host for the same content
        lab = np.load(output_paths(paths[0])[2])
        pal = np.array([[c[2], c[1], c[0], c[3]] for c in spec.PALETTE], np.uint8)
        t = []
        for _ in range(2):
            t0 = time.perf_counter()
            img = cv2.imread(paths[0], cv2.IMREAD_UNCHANGED)
            cv2.imwrite(os.path.join(d, "ref_dapi.tif"), 255 - img)
            cv2.imwrite(os.path.join(d, "ref.png"), pal[lab])
            np.save(os.path.join(d, "ref.npy"), lab)
            t.append(time.perf_counter() - t0)
        return {"value": st["images_per_s"], "unit": "images/s", "images": n, "wall_s": st["wall_s"],
                "what": "uncompressed 2048x2048 u8 TIFF files -> dapi tif + RGBA png + int64 npy files, decode / GPU / "
                        "writes overlapped (ecseg_b200.pipeline); PNG deflated on the GPU, .npy payload widened on the GPU",
                "where": d.rsplit("/", 1)[0], "readers": args.readers, "writers": args.writers,
                "reader_busy_s": st["reader_busy_s"], "writer_busy_s": st["writer_busy_s"],
                "png_bytes_per_image": png_bytes, "file_bytes_written_per_image": png_bytes + 8 * H * W + 128 + H * W + 128,
                "cpu_reference_io_ms_per_image": float(np.min(t)) * 1e3,
                "cpu_reference_io_note": "cv2.imread + cv2.imwrite(tif) + cv2.imwrite(RGBA png; stands in for plt.imsave) + "
                                         "np.save(int64), one thread, same files: the I/O the reference adds to every image",
                "n_ec_sum": int(sum(r[1] for r in rows))}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# --------------------------------------------------------------------------------------------
# --workload postproc: BASELINE.json config 4 (post-processing only, HBM-bound integer work)
# --------------------------------------------------------------------------------------------
PP_BYTES_PER_PX = 53      # algorithmic bytes with the two no-op merge_comp passes elided


def postproc_line(args, steps=None, cpu_baseline=True):
    """BASELINE.json config 4.  A step = `--maps-per-step` synthetic 4-class 2048x2048 label maps through
    meta_inference + count_cc on one GPU, spread over `--pp-contexts` post-processing-only contexts / streams (label
    maps are independent).  Returns the JSON line as a dict."""
    import torch
    from ctypes import c_void_p

    from ecseg_b200 import synth
    from ecseg_b200.engine import Engine

    local = int(os.environ.get("LOCAL_RANK", "0"))
    steps = steps or args.steps
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, B = max(1, args.pp_contexts), args.maps_per_step
    engs = [Engine(local, H, W, max_tiles=0) for _ in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    n_distinct = args.pp_distinct
    host_maps = [torch.from_numpy(synth.synth_label_map(s, H, W)).pin_memory() for s in range(n_distinct)]
    pristine = [m.to(dev) for m in host_maps]
    work = [torch.empty((H, W), dtype=torch.uint8, device=dev) for _ in range(B)]      # B x 4 MB cycled; workspace
    n_out = torch.zeros(B, dtype=torch.int32, device=dev)                              # per context is ~120 MB > L2
    px_out = torch.zeros(B, dtype=torch.int64, device=dev)
    host_out = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in range(S)]
    torch.cuda.synchronize()

    def fork():
        ev = torch.cuda.Event(); ev.record()
        for s_ in streams:
            s_.wait_event(ev)

    def join():
        for s_ in streams:
            ev = torch.cuda.Event(); ev.record(s_)
            torch.cuda.current_stream().wait_event(ev)

    def step(i0, e2e=False):
        for j in range(B):
            k = j % S
            e = engs[k]
            with torch.cuda.stream(streams[k]):
                if e2e:
                    work[j].copy_(host_maps[(i0 + j) % n_distinct], non_blocking=True)
                else:
                    work[j].copy_(pristine[(i0 + j) % n_distinct], non_blocking=True)
                e._chk(e.lib.ecseg_postprocess(e.ctx, work[j].data_ptr(), H, W, 0, n_out[j:].data_ptr(), px_out[j:].data_ptr(),
                                               c_void_p(streams[k].cuda_stream)))
                if e2e:
                    host_out[k].copy_(work[j], non_blocking=True)

    for i in range(args.warmup):
        fork(); step(i * B); join()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    l0 = sum(e.launch_count() for e in engs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); fork()
    for i in range(steps):
        step(i * B)
    join(); ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = sum(e.launch_count() for e in engs) - l0
    clocks = sampler.stop()
    counts = n_out.cpu().numpy().tolist()
    # one context alone, one map at a time: the per-launch latency view of the same kernels
    # (on a non-default stream, like every product path: the launch sequence is replayed as a CUDA graph there)
    n_single = min(8, B)
    with torch.cuda.stream(streams[0]):
        for rep in range(2):                 # first pass instantiates the graphs of these buffers, second is timed
            ev0.record()
            for j in range(n_single):
                work[j].copy_(pristine[j % n_distinct], non_blocking=True)
                engs[0]._chk(engs[0].lib.ecseg_postprocess(engs[0].ctx, work[j].data_ptr(), H, W, 0, n_out[j:].data_ptr(),
                                                           px_out[j:].data_ptr(), c_void_p(streams[0].cuda_stream)))
            ev1.record()
            streams[0].synchronize()
    ms_single = ev0.elapsed_time(ev1) / n_single
    # end to end: pinned host map in, label map + count back on the host
    fork(); step(0, e2e=True); join(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fork(); step(i * B, e2e=True); join()
        n_out.cpu()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    _, hbm_peak, peak_src = peaks()
    maps = B * steps
    value = maps / (ms / 1e3)
    achieved = PP_BYTES_PER_PX * H * W * value / 1e9
    line = {
        "metric": "meta_inference + count_cc label maps/s (2048x2048, BASELINE.json config 4)", "value": value, "unit": "maps/s",
        "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": f"{B} synthetic 4-class 2048x2048 label maps per step ({n_distinct} distinct, ellipses + discs + 2000 "
                               "salt pixels), fill_holes x2 + size_thresh + boundary erase + nucleus-in-metaphase + dilation + "
                               "count_cc(I==3); merge_comp x2 elided (proven no-op)",
                   "contexts": S, "l2": f"{S} contexts x ~120 MB of label / statistics workspace in flight, beyond the 126 MB L2"},
        "e2e": {"value": maps / e2e_s, "unit": "maps/s", "h2d_bytes_per_step": B * H * W, "d2h_bytes_per_step": B * (H * W + 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": None, "kernel": "k_ccl_tile / k_ccl_border / k_ccl_roots + rule kernels (%d launches per map, replayed as one CUDA graph)" % round(launches / max(1, maps)),
                     "peak_source": peak_src, "algorithmic_bytes_per_map": PP_BYTES_PER_PX * H * W,
                     "single_stream_ms_per_map": ms_single,
                     "single_stream_gbs": PP_BYTES_PER_PX * H * W / (ms_single / 1e3) / 1e9},
        "clocks": clocks, "n_ec_first_maps": counts[:4],
    }
    if cpu_baseline and not args.no_cpu_baseline:
        from oracle import metaseg_oracle as mo
        m = host_maps[0].numpy().astype(np.int64)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = mo.meta_inference(m.copy())
            ref_n = mo.count_cc(out == 3)[0]
        sec = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "maps/s", "cores": 1, "kind": "port",
                                "sample": "one 2048x2048 label map through oracle meta_inference (with merge_comp, as the "
                                          "reference runs it) + count_cc", "n_ec": int(ref_n), "gpu_n_ec_same_map": counts[0]}
    for e in engs:
        e.close()
    return line


def run_postproc(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    print(json.dumps(postproc_line(args)), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def unet_traffic():
    """DRAM bytes per image of the U-Net launches from the committed ncu --set full capture (tools/ncu_traffic.py);
    per image like `achieved`.  None when the capture is not in the tree."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "unet_dram_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return t["bytes_per_image"], f"{t['what']}; {t['launches']} launches; {t['source']}"
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture in profiles/"


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from ecseg_b200 import spec, synth, weights as wmod
    from ecseg_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # `--contexts` library contexts per GPU, each on its own CUDA stream: image j of a step goes to
    # context j mod C, so the small latency-bound post-processing kernels and the H2D / D2H copies of
    # one image overlap the U-Net of the next (the U-Net kernels are persistent, one CTA per SM).
    n_ctx = max(1, args.contexts)
    weights = wmod.make_weights(0)
    engs = [Engine(local, H, W) for _ in range(n_ctx)]
    for e in engs:
        e.load_weights(weights, args.precision)
    eng = engs[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_ctx)]
    B = args.images_per_step
    pool = max(B, 8)
    # distinct images per rank; every image's activations (~6 GB) dwarf the 126 MB L2
    host_imgs = [torch.from_numpy(synth.synth_dapi(1000 * rank + s, H, W)).pin_memory() for s in range(pool)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    host_labels = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in range(n_ctx)]
    host_dapi = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in range(n_ctx)]   # utils.py:112 writes it per image
    outs = [(torch.empty((H, W), dtype=torch.uint8, device=dev), torch.empty((H, W), dtype=torch.uint8, device=dev),
             torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev))
            for _ in range(n_ctx)]
    torch.cuda.synchronize()

    def fork():
        ev = torch.cuda.Event()
        ev.record()
        for s_ in streams:
            s_.wait_event(ev)

    def join():
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            torch.cuda.current_stream().wait_event(ev)

    def step_resident(i0):
        for j in range(B):
            k = j % n_ctx
            with torch.cuda.stream(streams[k]):
                engs[k].segment_device(dev_imgs[(i0 + j) % pool], H, W, 1, 1, out=outs[k])

    def step_e2e(i0):
        n = 0
        pending = [False] * n_ctx
        for j in range(B):
            k = j % n_ctx
            if pending[k]:
                n += engs[k].segment_host_wait()[0]
            with torch.cuda.stream(streams[k]):
                engs[k].segment_host_async(host_imgs[(i0 + j) % pool].numpy(), host_labels[k].numpy(), host_dapi[k].numpy())
            pending[k] = True
        for k in range(n_ctx):          # every step ends with its results (labels + counts) on the host
            if pending[k]:
                n += engs[k].segment_host_wait()[0]
        return n

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        fork(); step_resident(i * B); join()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = sum(e.launch_count() for e in engs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fork()
    for i in range(args.steps):
        step_resident(i * B)
    join()
    ev1.record()
    torch.cuda.synchronize()
    launches = sum(e.launch_count() for e in engs) - launches0
    ms = ev0.elapsed_time(ev1)
    barrier()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop() if sampler else None

    # stage shares over a few more images for the roofline of the dominant (U-Net) kernels
    # (one context alone, back to back, first half discarded: the run is long enough -- ~0.6 s -- to sit under the
    #  same power cap as the timed region, so the U-Net time is a SUSTAINED figure like the peak it is divided by)
    stage = np.zeros(4)
    n_stage = args.stage_images
    for i in range(n_stage):
        with torch.cuda.stream(streams[0]):      # a non-default stream, like the timed loops and the product pipeline
            eng.segment_device(dev_imgs[i % pool], H, W, 1, 1, out=outs[0])
        if i >= n_stage // 2:
            stage += np.array(eng.last_stage_ms())
    stage /= n_stage - n_stage // 2

    # ---- end to end with host buffers ----
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i * B)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i * B)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3     # host clock around synchronous steps: copies + device time
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(t.item())
    dev_err = eng.device_error()

    if rank == 0:
        tf_peak, hbm_peak, peak_src = peaks()
        total_images = world * B * args.steps
        value = total_images / (ms_max / 1e3)
        e2e_val = total_images / (e2e_ms_max / 1e3)
        flops_img = spec.unet_flops_per_tile() * TILES_PER_IMAGE
        achieved = flops_img / (stage[1] / 1e3) / 1e12
        from ecseg_b200.engine import unet_work
        flops_ref, flops_exec = unet_work(H, W, labels_only=True)
        assert abs(flops_ref - flops_img) < 1e-6 * flops_img
        achieved_exec = flops_exec / (stage[1] / 1e3) / 1e12
        traffic, traffic_note = unet_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(B, n_ctx, pool),
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * H * W,
                    "d2h_bytes_per_step": B * (2 * H * W + 24),
                    "what": "pinned host image in; label map + dapi plane (what src/utils.py:112 saves) + count and status words back"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "kernel": "k_conv_tc (tcgen05 implicit-GEMM conv, 21 launches per image) + k_head_tc = the U-Net stage",
                         "peak_source": peak_src, "flops_per_image": flops_img, "unet_ms_per_image": float(stage[1]),
                         "flops_executed_per_image": flops_exec, "achieved_executed": achieved_exec,
                         "frac_executed": achieved_exec / tf_peak,
                         "flops_note": "achieved / frac use the ALGORITHMIC FLOPs of SURVEY 8(d): what model.predict_on_batch computes, "
                                       "every tile in full (97.014 GFLOP x 100 tiles).  The whole-image path issues fewer: the last four "
                                       "layers skip the 16x16 blocks that lie entirely in the part of a tile the reference's stitcher never "
                                       "takes (tiles overlap by 25 px; bit-identical label map, tests/test_gpu_*).  *_executed are the "
                                       "same figures with the FLOPs actually issued -- the tensor pipe's own utilisation."},
            "stage_ms_note": "stages timed on one context running alone (no overlap), 64 images back to back, mean of the last 32",
            "stage_ms_per_image": {"preprocess": float(stage[0]), "unet": float(stage[1]), "stitch": float(stage[2]),
                                   "postprocess": float(stage[3])},
            "postprocess_hbm": {"algorithmic_bytes": 53 * H * W, "achieved_gbs": 53 * H * W / (stage[3] / 1e3) / 1e9,
                                "peak_gbs": hbm_peak, "frac": 53 * H * W / (stage[3] / 1e3) / 1e9 / hbm_peak},
            "clocks": clocks, "device_error": dev_err,
        }
        if world == 1 and not args.no_extras:
            line["config1_example"] = run_config1(eng, args.precision)
            line["config5_overlay_chain"] = run_config5(eng)
    for e in engs:
        e.close()
    if args.artifact_images > 0:
        # BASELINE config 3 "with artefacts": every rank runs the product pipeline over its own TIFF files
        art = run_artifacts(args, weights, local, rank, barrier)
        t = torch.tensor([art["wall_s"], float(art["images"])], device=dev, dtype=torch.float64)
        if world > 1:
            wall = t[:1].clone(); dist.all_reduce(wall, op=dist.ReduceOp.MAX)
            tot = t[1:].clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            art["images"] = int(tot.item()); art["wall_s"] = float(wall.item())
            art["value"] = art["images"] / art["wall_s"]
            art["what"] += f"; {world} ranks, one pipeline per GPU, whole-job images / slowest rank's wall time"
        if rank == 0:
            line["artifacts"] = art
    if rank == 0:
        if world == 1 and not args.no_extras:
            pl = postproc_line(args, steps=max(1, args.pp_maps // args.maps_per_step), cpu_baseline=False)
            line["config4_postproc"] = {k: pl[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "e2e", "gpu_launches",
                                                           "roofline", "n_ec_first_maps")}
        if world == 1 and not args.no_cpu_baseline:
            net = make_oracle_net()
            img = host_imgs[0].numpy()
            cpu_step(net, img, 2)
            sec, stages = cpu_step(net, img)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "ONE whole 2048x2048 image of the workload through the oracle port (oracle/, "
                                              "torch-CPU fp32 U-Net on all 100 tiles, scipy post-processing), after a 2-tile warm-up",
                                    "cpu_stage_s_per_image": {k: float(v) for k, v in zip(CPU_STAGES, stages)}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config1(eng, precision):
    """BASELINE config 1 as an extra of the bench line: input.tif := 255 - example_ecSeg/dapi.jpeg (1040x1392, 35 tiles;
    pixels frozen in tests/golden/example.npz together with what the reference's utils.meta_segment returned for it with
    the fp32 CPU U-Net) through the whole GPU path: throughput, label agreement, count."""
    import torch
    path = os.path.join(ROOT, "tests", "golden", "example.npz")
    if not os.path.isfile(path):
        return {"unavailable": "tests/golden/example.npz missing"}
    g = np.load(path)
    img = np.ascontiguousarray(g["input"])
    h, w = img.shape
    notie = np.unpackbits(g["notie"])[: h * w].reshape(h, w).astype(bool)
    pinned = torch.from_numpy(img).pin_memory()
    out = torch.empty((h, w), dtype=torch.uint8).pin_memory()
    n_ec = 0
    for _ in range(3):
        eng.segment_host_async(pinned.numpy(), out.numpy()); n_ec = eng.segment_host_wait()[0]
    reps = 24
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.segment_host_async(pinned.numpy(), out.numpy()); eng.segment_host_wait()
    dt = (time.perf_counter() - t0) / reps
    pre, _ = eng.preprocess(img)
    raw = eng.stitch_argmax(eng.unet_forward(eng.tile(pre)), h, w).cpu().numpy()
    return {"workload": "255 - example_ecSeg/dapi.jpeg, 1040x1392 u8, 35 tiles, seed-0 weights", "images_per_s_e2e": 1.0 / dt,
            "ms_per_image_e2e": dt * 1e3, "contexts": 1, "precision": precision,
            "label_agreement_vs_reference_run": float((raw == g["raw"])[notie].mean()),
            "label_agreement_note": "raw (pre-meta_inference) label map vs the reference's utils.meta_segment flow run with the "
                                    "fp32 CPU U-Net (tests/golden/example.npz), quantised top-2 ties excluded",
            "n_ec": int(n_ec), "n_ec_reference_run": int(g["count"][0]),
            "final_map_equal_px_frac": float((out.numpy() == g["final"]).mean())}


def run_config5(eng):
    """BASELINE config 5 as an extra: metaseg followed by meta_overlay (src/meta_overlay.py:59-83) on synthetic RGB
    DAPI + green / red FISH images, chained on the GPU (the label map never leaves it)."""
    import torch
    from ecseg_b200 import synth
    dev = eng.device
    imgs = [torch.from_numpy(synth.synth_fish(900 + s, H, W)).to(dev) for s in range(4)]
    res = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_seg = t_ov = 0.0
    reps = 8
    for i in range(reps + 2):
        im = imgs[i % 4]
        ev[0].record()
        labels, _dapi, _n, _px = eng.segment_device(im, H, W, 3, 1)
        ev[1].record()
        res = eng.overlay_counts(im, labels, 85)          # syncs (reads the 12 counts back)
        ev[2].record(); torch.cuda.synchronize()
        if i >= 2:
            t_seg += ev[0].elapsed_time(ev[1]); t_ov += ev[1].elapsed_time(ev[2])
    return {"workload": "2048x2048 RGB u8 (DAPI in B, FISH in G/R), color_sensitivity 85: ecseg_segment_image then "
                        "ecseg_overlay_counts on the same device buffers", "metaseg_ms_per_image": t_seg / reps,
            "meta_overlay_ms_per_image": t_ov / reps, "images_per_s_chain": 1e3 * reps / (t_seg + t_ov),
            "last_counts": res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ecseg_b200", choices=["ecseg_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("ECSEG_PRECISION", "fp16"), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--images-per-step", type=int, default=8)
    ap.add_argument("--contexts", type=int, default=2, help="library contexts (CUDA streams) per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 1 / 4 / 5 extras of the bench line")
    ap.add_argument("--pp-maps", type=int, default=4096, help="label maps of the config-4 extra (BASELINE: 4096)")
    ap.add_argument("--pp-distinct", type=int, default=64, help="distinct synthetic label maps cycled in the config-4 runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-images", type=int, default=64, help="images of the per-stage (roofline) measurement")
    ap.add_argument("--workload", default="metaseg", choices=["metaseg", "postproc"],
                    help="metaseg = the headline bench line; postproc = BASELINE.json config 4 (extra line, HBM roofline)")
    ap.add_argument("--maps-per-step", type=int, default=64)
    ap.add_argument("--pp-contexts", type=int, default=8)
    ap.add_argument("--artifact-images", type=int, default=64, help="files through the with-artefacts pipeline (0 = skip)")
    ap.add_argument("--readers", type=int, default=4)
    ap.add_argument("--writers", type=int, default=6)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    # timing rule: at least 3 warm-up steps on the GPU arms, whatever was asked for (the line reports what was done)
    args.warmup = max(args.warmup, 3)
    if args.workload == "postproc":
        run_postproc(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
