#!/usr/bin/env python
"""bench.py -- XFeat extract+match frames/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frames (VGA by default; --height 720 --width 1280
is BASELINE.json's 1280x720 stream): every frame is extracted (top-4096 keypoints + 64-D descriptors) and
brute-force matched (4096 x 4096, best / second-best / reverse-best) against the previous frame of the stream --
one extract + one match per frame.  The batch of a step is --chunks x --batch frames (default 64 x 32 = 2048),
processed as --chunks launches groups of --batch frames (the activations of 32 VGA frames are 2.2 GB), so that the
K timed steps last seconds, not milliseconds: clocks and power are at steady state when the number is taken.

  value : whole-job frames/s with the input frames already resident in HBM (xfb_extract_batch_device +
          xfb_match_frame_pairs_device), CUDA-event timed on the launching stream, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI (xfb_submit / xfb_wait, two batches in flight):
          pinned host frames in, every result (keypoints, scores, descriptors, five match arrays) back in
          host memory, H2D and D2H copies inside the timed region (host clock around device sync).
  roofline / cpu_baseline : see DESIGN.md "Measurement".

--impl reference times the reference's own CPU implementation of the path (oracle/_ref/ref_xfeat =
the reference's XFextractor compiled unchanged, all host threads; matcher = the C port in
oracle/matcher_oracle.c, single thread like the reference's matchers) on bounded samples.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "xfeat_extract_match_fps_vga"
UNIT = "frames/s"
H, W, TOPK = 480, 640, 4096          # defaults; --height / --width override H, W (see main)
INT_MAX = 2 ** 31 - 1

# conv MACs per frame at H x W (SURVEY.md 8a per-layer table): name -> (cin, cout, k, output downsample log2)
LAYER_GEOM = {
    "block1.0": (1, 4, 3, 0), "block1.1": (4, 8, 3, 1), "block1.2": (8, 8, 3, 1), "block1.3": (8, 24, 3, 2),
    "block2.0": (24, 24, 3, 2), "block2.1": (24, 24, 3, 2), "block3.0": (24, 64, 3, 3), "block3.1": (64, 64, 3, 3),
    "block3.2": (64, 64, 1, 3), "block4.0": (64, 64, 3, 4), "block4.1": (64, 64, 3, 4), "block4.2": (64, 64, 3, 4),
    "block5.0": (64, 128, 3, 5), "block5.1": (128, 128, 3, 5), "block5.2": (128, 128, 3, 5), "block5.3": (128, 64, 1, 5),
    "block_fusion.0": (64, 64, 3, 3), "block_fusion.1": (64, 64, 3, 3), "block_fusion.2": (64, 64, 1, 3),
    "heatmap_head.0": (64, 64, 1, 3), "heatmap_head.1": (64, 64, 1, 3), "keypoint_head.0": (64, 64, 1, 3),
    "keypoint_head.1": (64, 64, 1, 3), "keypoint_head.2": (64, 64, 1, 3),
}


def layer_flops(name, h, w):
    cin, cout, k, lvl = LAYER_GEOM[name]
    return 2.0 * cin * cout * k * k * (h >> lvl) * (w >> lvl)


def conv_flops_per_frame(h, w):
    f = sum(layer_flops(n, h, w) for n in LAYER_GEOM)
    f += 2.0 * 24 * (h >> 2) * (w >> 2)              # skip1.1
    f += 2.0 * 64 * 1 * (h >> 3) * (w >> 3)          # heatmap_head.2
    f += 2.0 * 64 * 65 * (h >> 3) * (w >> 3)         # keypoint_head.3
    return f


STRIDE2 = {"block1.1", "block1.3", "block3.0", "block4.0", "block5.0"}   # layers whose input map is one level finer than the output


def conv_layer_roofline_safe(*a):
    try:
        return conv_layer_roofline(*a)
    except Exception as ex:   # a reporting extra must never cost the bench line
        return {"error": repr(ex)}


def conv_layer_roofline(ms_per_launch, h, w, batch, peaks, split_cost=None, split_name=None):
    """T_roof of SURVEY.md 8d with measured peaks: per conv layer max(FLOP time, byte time), summed, against the measured time.
    FLOP time: block1 on the FP32 SIMT pipe (80 TFLOP/s nominal), every other layer on the tensor cores as an fp32-accurate
    split product costing `split_cost` plain bf16/fp16 MMA passes (measured bf16 peak / split_cost); bytes = input + output
    activations once (fp32) + weights."""
    split_cost = CONV_SPLIT_COST if split_cost is None else split_cost
    split_name = CONV_SPLIT_NAME if split_name is None else split_name
    p_tensor = peaks["bf16_tflops"] * 1e12 / split_cost
    p_simt = 80e12
    bw = peaks["hbm_gbs"] * 1e9
    meas = roof = 0.0
    for name, (cin, cout, k, lvl) in LAYER_GEOM.items():
        if name not in ms_per_launch:
            continue
        lin = lvl - 1 if name in STRIDE2 else lvl
        byts = 4.0 * batch * (cin * (h >> lin) * (w >> lin) + cout * (h >> lvl) * (w >> lvl)) + 4.0 * cin * cout * k * k
        tf = layer_flops(name, h, w) * batch / (p_simt if name.startswith("block1") else p_tensor)
        roof += max(tf, byts / bw) * 1e3
        meas += ms_per_launch[name]
    if meas <= 0.0:
        return None
    return {"layers": "the BasicLayer convolutions (block1 .. keypoint_head.2)", "measured_ms": meas, "roof_ms": roof, "frac": roof / meas,
            "pipes": "block1: fp32 SIMT 80 TFLOP/s nominal; others: %s = %s bf16 peak / %g; bytes at the %s copy bandwidth"
                     % (split_name, peaks["source"], split_cost, peaks["source"])}


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(n_extract, n_match, threads, topk=TOPK, h=H, w=W, warm=2):
    """Times the reference CPU path on bounded samples.  Returns dict with per-frame / per-pair seconds."""
    from oracle import matcher_oracle as mo
    from xfeatslam_b200.frames import synthetic_frames
    ref = REPO / "oracle" / "_ref" / "ref_xfeat"
    out = {"kind": "reference" if ref.exists() else "port"}
    frames = synthetic_frames(900, max(n_extract, 1), h, w)
    if ref.exists():
        with tempfile.TemporaryDirectory() as td:
            fp = Path(td) / "frames.u8"
            frames.tofile(fp)
            r = subprocess.run([str(ref), "bench", str(fp), str(h), str(w), str(topk), str(n_extract), str(warm), str(threads)],
                               check=True, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        j = json.loads(line)
        out["extract_s_per_frame"] = j["seconds"] / j["frames"]
        out["extract_threads"] = j["threads"]
    else:
        import torch
        from oracle import xfeat_oracle as xo
        torch.set_num_threads(threads)
        wts = xo.load_weights()
        for i in range(min(warm, len(frames))):
            xo.detect_and_compute(frames[i], wts, topk)
        t0 = time.perf_counter()
        for i in range(n_extract):
            xo.detect_and_compute(frames[i], wts, topk)
        out["extract_s_per_frame"] = (time.perf_counter() - t0) / n_extract
        out["extract_threads"] = threads
    rng = np.random.RandomState(0)
    A = rng.randn(topk, 64).astype(np.float32); A /= np.linalg.norm(A, axis=1, keepdims=True)
    B = (A[rng.permutation(topk)] + 0.05 * rng.randn(topk, 64)).astype(np.float32); B /= np.linalg.norm(B, axis=1, keepdims=True)
    mo.lib()
    t0 = time.perf_counter()
    for _ in range(n_match):
        mo.bruteforce(A, B)
    out["match_s_per_pair"] = (time.perf_counter() - t0) / n_match if n_match else 0.0
    out["fps"] = 1.0 / (out["extract_s_per_frame"] + out["match_s_per_pair"])
    return out


def best_reference_threads(ncpu, h=H, w=W):
    """libtorch's intra-op pool scales badly past ~16-32 threads on this small CNN: probe a few settings on a
    3-frame sample and give the reference its best one ("all the host threads it can use")."""
    best, best_t = ncpu, None
    for thr in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
        try:
            r = cpu_reference_run(n_extract=3, n_match=0, threads=thr, warm=1, h=h, w=w)
        except Exception:
            continue
        if best_t is None or r["extract_s_per_frame"] < best_t:
            best, best_t = thr, r["extract_s_per_frame"]
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    Hf, Wf = args.height, args.width
    ncpu = os.cpu_count() or 1
    n_frames = max(1, args.steps + args.warmup)
    n_match = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    thr = best_reference_threads(ncpu, Hf, Wf)
    r = cpu_reference_run(n_extract=n_frames, n_match=n_match, threads=thr, warm=max(args.warmup, 1), h=Hf, w=Wf)
    wall = time.perf_counter() - t0
    value = r["fps"]
    sample = ("each step = ONE frame of the workload: %d %dx%d frames through the reference XFextractor::operator() (libtorch CPU, %d threads: %.1f ms/frame) "
              "+ %d brute-force 4096x4096 DescriptorDistance matches (C port, 1 thread like the reference's matchers: %.2f s/pair); fps = 1/(extract+match)"
              % (n_frames, Wf, Hf, r["extract_threads"], r["extract_s_per_frame"] * 1e3, n_match, r["match_s_per_pair"]))
    line = {
        "impl": "reference", "metric": metric_name(Hf, Wf), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncpu, "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    """The workload both arms are quoted on (identical dict in the b200 and the reference line; run details live in `run`)."""
    return {"workload": workload_name(args.height, args.width), "frames_per_step_per_gpu": max(1, args.chunks) * args.batch,
            "launch_groups_per_step": max(1, args.chunks), "frames_per_launch_group": args.batch, "topk": TOPK, "matches_per_frame": 1,
            "pairing": PAIRS_DOC, "match_size": "4096x4096x64", "parallelism": "frame-sharded dp%d, no data-path collective" % world}


def workload_name(h, w):
    tag = "vga" if (h, w) == (480, 640) else ("hd" if (h, w) == (720, 1280) else "img")
    return "%s_%dx%d_top%d_extract+match_prev" % (tag, w, h, TOPK)


def metric_name(h, w):
    return METRIC if (h, w) == (480, 640) else "xfeat_extract_match_fps_%dx%d" % (w, h)


def kernel_roofline(name, ms_per_batch, launches_per_batch, h, w, batch, peaks, sustained):
    """Roofline of one kernel tag of the per-kernel profile: algorithmic work of ONE batch (SURVEY 8d per-unit figures x the
    units a batch holds) / the time all launches of that tag take per batch.  Tensor-core kernels against the measured cuBLAS
    bf16 rate (sustained figure: the kernel is timed inside a seconds-long loop), byte kernels against the measured copy rate."""
    Hi, Wi = (h // 32) * 32, (w // 32) * 32
    bf16 = (peaks["bf16_tflops_sustained"] if sustained and peaks.get("bf16_tflops_sustained") else peaks["bf16_tflops"])
    src = peaks["source"] + (" cuBLAS bf16 sustained" if sustained and peaks.get("bf16_tflops_sustained") else " cuBLAS bf16 burst")
    r = {"kernel": name, "avg_launch_ms": ms_per_batch / max(launches_per_batch, 1), "ms_per_batch": ms_per_batch, "launches_per_batch": launches_per_batch}
    if name == "match_tile":
        alg = 2.0 * TOPK * TOPK * 64 * batch     # ONE GEMM per frame pair (mutual-NN reuses the same matrix column-wise)
        r.update({"bound": "tensor", "achieved": alg / (ms_per_batch * 1e-3) / 1e12, "peak": bf16, "unit": "TFLOP/s", "peak_source": src,
                  "algorithmic_flops_per_batch": alg,
                  "pipe_used": "tcgen05 kind::f16 filter GEMM (fp16 operands, fp32 TMEM accumulators) + exact fp64 verification of the survivors; "
                               "algorithmic work = one 2*N1*N2*64 GEMM per pair, whatever the kernel executes"})
    elif name in LAYER_GEOM and not name.startswith("block1"):
        alg = layer_flops(name, Hi, Wi) * batch
        peak = bf16 / CONV_SPLIT_COST
        r.update({"bound": "tensor", "achieved": alg / (ms_per_batch * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                  "peak_source": src + " / %g (%s)" % (CONV_SPLIT_COST, CONV_SPLIT_NAME), "algorithmic_flops_per_batch": alg, "pipe_used": CONV_SPLIT_NAME})
    else:
        byts = kernel_bytes(name, Hi, Wi, batch)
        r.update({"bound": "hbm", "achieved": byts / (ms_per_batch * 1e-3) / 1e9 if byts else 0.0, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "peak_source": peaks["source"] + " copy bandwidth", "algorithmic_bytes_per_batch": byts, "pipe_used": "fp32 SIMT, HBM-bound"})
    r["frac"] = r["achieved"] / r["peak"] if r["peak"] else None
    return r


# the tensor-core convolutions compute an fp32-accurate product from split operands: cost in plain bf16/fp16 MMA passes
CONV_SPLIT_COST = 3.0
CONV_SPLIT_NAME = "tcgen05 kind::f16, fp16 two-piece split (A_hi x [W_hi ; W_lo] + A_lo x W_hi = three fp16 products per output, fp32 TMEM accumulators)"


def kernel_bytes(name, h, w, batch):
    """Compulsory HBM bytes of the byte-bound kernels per batch (fp32 activations in + out once)."""
    if name in LAYER_GEOM:
        cin, cout, k, lvl = LAYER_GEOM[name]
        lin = lvl - 1 if name in STRIDE2 else lvl
        return 4.0 * batch * (cin * (h >> lin) * (w >> lin) + cout * (h >> lvl) * (w >> lvl))
    px = h * w
    table = {"prep_stats": px * 1.0, "prep_norm": px * (1.0 + 4 + 4 + 0.25), "nms_score": px * 4.0 + px / 64 * 4, "pyramid_sum": px / 64 * 64 * 4 * 2.4,
             "keypoint_head.3": px / 64 * 64 * 4 + px * 4, "heatmap_out": px / 64 * 65 * 4, "describe": TOPK * (4 * 256 + 256 + 16),
             "topk_select_sort": 8192 * 8 * 2, "match_prep": TOPK * (256 + 160)}
    return batch * table.get(name, 0.0)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from xfeatslam_b200 import shard
    from xfeatslam_b200.capi import XFeatB200
    from xfeatslam_b200.frames import synthetic_frame

    Hf, Wf = args.height, args.width
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the b200 arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    Bsz, K, Wm, CH = args.batch, args.steps, args.warmup, max(1, args.chunks)
    POOL = 4

    NC = max(1, args.contexts)
    ctxs = [XFeatB200(max_h=Hf, max_w=Wf, max_batch=Bsz, max_topk=TOPK, device=local) for _ in range(NC)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
    for c_, s_ in zip(ctxs, streams):
        c_.set_stream(s_.cuda_stream)
    ctx, stream = ctxs[0], streams[0]

    # frame shard of this rank: frames are independent units, global frame i -> rank (i mod world)
    # (shard.frame_indices); the pool holds this rank's frames of POOL consecutive global batches
    host_pool = [torch.from_numpy(np.stack([synthetic_frame(7000 + g, Hf, Wf) for g in shard.frame_indices(rank, world, world * Bsz * POOL)[p * Bsz:(p + 1) * Bsz]])).pin_memory()
                 for p in range(POOL)]
    dev_pool = [t.to(dev) for t in host_pool]
    d_outs = [{"nv": torch.zeros(Bsz, dtype=torch.int32, device=dev), "xy": torch.zeros(Bsz, TOPK, 2, dtype=torch.float32, device=dev),
               "sc": torch.zeros(Bsz, TOPK, dtype=torch.float32, device=dev), "ds": torch.zeros(Bsz, TOPK, 64, dtype=torch.float32, device=dev),
               "m": [torch.zeros(Bsz, TOPK, dtype=torch.int32, device=dev) for _ in range(5)]} for _ in range(NC)]
    d_nv, d_m = d_outs[0]["nv"], d_outs[0]["m"]
    pairs = PAIRS_FN(Bsz)
    # two pinned host output sets per context: xfb_submit keeps two batches in flight (copies overlap compute)
    h_out = [{"nv": torch.zeros(Bsz, dtype=torch.int32).pin_memory(), "xy": torch.zeros(Bsz, TOPK, 2, dtype=torch.float32).pin_memory(),
              "sc": torch.zeros(Bsz, TOPK, dtype=torch.float32).pin_memory(), "ds": torch.zeros(Bsz, TOPK, 64, dtype=torch.float32).pin_memory(),
              "m": [torch.zeros(Bsz, TOPK, dtype=torch.int32).pin_memory() for _ in range(5)]} for _ in range(2 * NC)]

    def batch_device(i, which=None):
        """One launch group (extract + match of --batch frames) on context i mod NC; every pointer is device memory, nothing synchronises."""
        w = (i % NC) if which is None else which
        fr, o = dev_pool[i % POOL], d_outs[w]
        ctxs[w].extract_ptrs(fr.data_ptr(), Bsz, Hf * Wf, Hf, Wf, Wf, TOPK, 0.05, o["nv"].data_ptr(), o["xy"].data_ptr(), o["sc"].data_ptr(),
                             o["ds"].data_ptr(), device=True)
        ctxs[w].match_frame_pairs(pairs, INT_MAX, [t.data_ptr() for t in o["m"]], device=True)

    def batch_host(i):
        """End-to-end launch group through the host-buffer C-ABI: pinned frames in, every result back in host memory.
        Group i goes to context i mod NC, slot (i div NC) mod 2: 2 * NC groups in flight."""
        w, slot = i % NC, (i // NC) % 2
        fr, o = host_pool[i % POOL], h_out[w * 2 + slot]
        ctxs[w].submit(slot, fr.data_ptr(), Bsz, Hf * Wf, Hf, Wf, Wf, TOPK, 0.05, o["nv"].data_ptr(), o["xy"].data_ptr(), o["sc"].data_ptr(),
                       o["ds"].data_ptr(), pairs=pairs, init=INT_MAX, match_ptrs=[t.data_ptr() for t in o["m"]])

    def wait_all():
        for c_ in ctxs:
            c_.wait(0); c_.wait(1)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(batch_fn):
        """Wm warm-up steps, then EXACTLY K timed steps of CH launch groups each, over the NC contexts.  CUDA events on stream 0:
        the start event precedes all work (barrier), the end event is recorded on stream 0 after it has waited for the last
        work of every other stream."""
        for i in range(max(Wm * CH, NC)):
            batch_fn(i)
        barrier()
        l0 = sum(c_.launch_count() for c_ in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in streams[1:]:
            s_.wait_event(e0)
        for i in range(K * CH):
            batch_fn(Wm * CH + i)
        for s_ in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), sum(c_.launch_count() for c_ in ctxs) - l0

    def profiled(n_batches):
        """Per-kernel CUDA-event timing (xfb_profile_*): a separate pass on ONE context, so that no other stream overlaps the
        kernel being timed.  Not part of `value`."""
        barrier()
        ctx.profile(True)
        for i in range(n_batches):
            batch_device(i, which=0)
        barrier()
        prof = ctx.profile_read()
        ctx.profile(False)
        return prof

    def timed_e2e():
        """K steps of CH submissions, 2 * NC in flight; the timed region starts before the first H2D copy and ends when the
        last result is in host memory (xfb_wait), measured on the host clock around device synchronisation."""
        for i in range(max(Wm * CH, 2 * NC)):
            batch_host(i)
        wait_all()
        barrier()
        t0 = time.perf_counter()
        for i in range(K * CH):
            batch_host(Wm * CH + i)      # xfb_submit waits for this slot's previous results first
        wait_all()
        torch.cuda.synchronize(dev)
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        return ms

    if args.diagnose_e2e:
        return diagnose_e2e(args, dev, world, rank, Bsz, K, CH, Wm, Hf, Wf, host_pool, h_out, timed, batch_device, timed_e2e, batch_host, wait_all, barrier, ctxs)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(batch_device)
    ms_e2e = timed_e2e()
    PROF_BATCHES = 10
    prof = profiled(PROF_BATCHES)
    clocks = sampler.stop() if rank == 0 else None

    nv = d_nv.cpu().numpy()
    matched = int((d_m[0].cpu().numpy() >= 0).sum())
    frames_rank = K * CH * Bsz
    # NCCL carries counts / timings only: all-gather of int64[4] per rank + max-reduce of the time
    counters_t, ms_dev_max = shard.gather_counters(frames_rank, int(nv.sum()), matched, ms_dev, device=dev)
    _, ms_e2e_max = shard.gather_counters(frames_rank, int(nv.sum()), matched, ms_e2e, device=dev)
    counters_all = counters_t.numpy()
    total_frames = int(counters_all[:, 0].sum())
    value = total_frames / (ms_dev_max * 1e-3)
    e2e_value = total_frames / (ms_e2e_max * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        sustained = ms_dev_max >= 1500.0          # a seconds-long timed region: compare with the sustained cuBLAS figure
        # dominant kernel = the tag with the largest share of device time; all launches of a tag per batch count together
        tot = sum(v[0] for v in prof.values()) or 1.0
        name, (kms, kcnt) = max(prof.items(), key=lambda kv: kv[1][0])
        roofline = kernel_roofline(name, kms / PROF_BATCHES, kcnt / PROF_BATCHES, Hf, Wf, Bsz, peaks, sustained)
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/), per launch
        traffic = None
        for tp in sorted((REPO / "profiles").glob("r*_dram_traffic_per_launch.json"), reverse=True):
            tj = json.loads(tp.read_text())
            keys = {"match_tile": ("mm_kernel<1>", "ms_kernel")}.get(name, (name,))
            hits = [v for key in keys for k, v in tj.items() if key in k]
            if hits:
                traffic = hits[0]["dram_bytes_per_launch"] if isinstance(hits[0], dict) else hits[0]
                break
        shares = {k: round(v[0] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]}
        all_ms = {k: round(v[0] / PROF_BATCHES, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        Hi, Wi = (Hf // 32) * 32, (Wf // 32) * 32
        roofline.update({"traffic": traffic, "share_of_step": round(kms / tot, 4), "top_shares": shares, "kernel_ms_per_batch": all_ms,
                         "profiled_batches": PROF_BATCHES,
                         "whole_path_conv_tflops": conv_flops_per_frame(Hi, Wi) * value / max(world, 1) / 1e12,
                         "conv_layer_roofline": conv_layer_roofline_safe({k: v[0] / PROF_BATCHES for k, v in prof.items()}, Hi, Wi, Bsz, peaks),
                         "match_roofline": kernel_roofline("match_tile", prof["match_tile"][0] / PROF_BATCHES, prof["match_tile"][1] / PROF_BATCHES,
                                                           Hf, Wf, Bsz, peaks, sustained) if "match_tile" in prof else None})
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ncpu = os.cpu_count() or 1
            r = cpu_reference_run(n_extract=16 if Hf * Wf <= 480 * 640 else 6, n_match=3, threads=best_reference_threads(ncpu, Hf, Wf), h=Hf, w=Wf)
            cpu = {"value": r["fps"], "unit": UNIT, "cores": ncpu, "kind": r["kind"],
                   "sample": "%d %dx%d frames reference XFextractor (libtorch CPU, %d threads, %.1f ms/frame) + 3 brute-force 4096x4096 matches "
                             "(C port, 1 thread, %.2f s/pair)" % (16 if Hf * Wf <= 480 * 640 else 6, Wf, Hf, r["extract_threads"],
                                                                   r["extract_s_per_frame"] * 1e3, r["match_s_per_pair"])}
        frame_bytes = Hf * Wf
        out_bytes = TOPK * (8 + 4 + 256) + 4 + 5 * TOPK * 4
        line = {
            "metric": metric_name(Hf, Wf), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_dev_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "run": {"contexts": NC, "batches_in_flight_e2e": 2 * NC, "timed_region_s": ms_dev_max * 1e-3,
                    "l2": "working set of one launch group %.1f GB >> 126 MB L2; input pool of %d distinct groups rotates" % (Bsz * 0.07 * Hi * Wi / 307200.0, POOL)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": CH * Bsz * frame_bytes, "d2h_bytes_per_step": CH * Bsz * out_bytes,
                    "ms_per_step": ms_e2e_max / K, "timed_region_s": ms_e2e_max * 1e-3},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "counters": {"frames": total_frames, "keypoints_last_group": int(counters_all[:, 1].sum()), "matched_last_group": int(counters_all[:, 2].sum())},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for c_ in ctxs:
        c_.close()
    return 0


def diagnose_e2e(args, dev, world, rank, Bsz, K, CH, Wm, Hf, Wf, host_pool, h_out, timed, batch_device, timed_e2e, batch_host, wait_all, barrier, ctxs):
    """Where does the end-to-end time go when N ranks share one host?  Per rank: (a) the copies alone (the H2D / D2H traffic of
    xfb_submit replayed with torch on two streams, no kernels), (b) the kernels alone (device-resident), (c) the real xfb_submit
    path, (d) host time spent inside the xfb_submit calls.  Rank 0 prints every rank's numbers (one JSON line)."""
    import torch
    import torch.distributed as dist
    n_groups = K * CH
    # (a) copies only
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    d_in = torch.empty_like(host_pool[0], device=dev)
    o0 = h_out[0]
    d_o = {k: (torch.empty_like(v, device=dev) if not isinstance(v, list) else [torch.empty_like(x, device=dev) for x in v]) for k, v in o0.items()}
    def copies(i):
        o = h_out[i % len(h_out)]
        with torch.cuda.stream(s_in):
            d_in.copy_(host_pool[i % len(host_pool)], non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in ("nv", "xy", "sc", "ds"):
                o[k].copy_(d_o[k], non_blocking=True)
            for a_, b_ in zip(o["m"], d_o["m"]):
                a_.copy_(b_, non_blocking=True)
    for i in range(8):
        copies(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_groups):
        copies(i)
    torch.cuda.synchronize(dev)
    ms_copy = (time.perf_counter() - t0) * 1e3
    barrier()
    # (b) kernels only, (c) the real path, (d) host time inside xfb_submit
    ms_dev, _ = timed(batch_device)
    ms_e2e = timed_e2e()
    for i in range(4):
        batch_host(i)
    wait_all()
    barrier()
    cpu = 0.0
    t_all = time.perf_counter()
    for i in range(n_groups):
        t1 = time.perf_counter()
        batch_host(Wm * CH + i)
        cpu += time.perf_counter() - t1
    wait_all()
    torch.cuda.synchronize(dev)
    ms_e2e2 = (time.perf_counter() - t_all) * 1e3
    barrier()
    frame_bytes = Hf * Wf
    out_bytes = TOPK * (8 + 4 + 256) + 4 + 5 * TOPK * 4
    mine = torch.tensor([ms_copy / n_groups, ms_dev / n_groups, ms_e2e / n_groups, ms_e2e2 / n_groups, cpu * 1e3 / n_groups], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    if rank == 0:
        rows = [[round(float(x), 4) for x in v.cpu()] for v in allv]
        gb = Bsz * (frame_bytes + out_bytes) / 1e9
        print(json.dumps({"diagnose_e2e": True, "n_gpus": world, "workload": workload_name(Hf, Wf), "frames_per_launch_group": Bsz,
                          "bytes_per_group": {"h2d": Bsz * frame_bytes, "d2h": Bsz * out_bytes},
                          "per_rank_ms_per_group": {"columns": ["copies_only", "kernels_only", "e2e_submit", "e2e_submit_again", "host_time_in_submit"], "rows": rows},
                          "copy_GBps_per_rank": [round(gb / (r[0] * 1e-3), 2) for r in rows],
                          "copy_GBps_aggregate": round(sum(gb / (r[0] * 1e-3) for r in rows), 2), "host_cpus": os.cpu_count()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for c_ in ctxs:
        c_.close()
    return 0


def PAIRS_FN(bsz):
    """Frame i of a launch group is matched against frame i - 1 of the group; frame 0 against the group's LAST frame (the
    groups of the synthetic stream are independent draws, so 'the previous frame' of frame 0 is simply another frame:
    same work, one 4096 x 4096 match per frame).  A group of one frame has no partner."""
    if bsz < 2:
        raise SystemExit("--batch must be >= 2: a frame is matched against another frame of its launch group")
    return np.array([[i, (i - 1) % bsz] for i in range(bsz)], np.int32)


PAIRS_DOC = "frame i vs frame i-1 of its launch group (frame 0 vs the group's last frame); --batch >= 2 enforced, never a self-match"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per launch group (one xfb_extract_batch + one xfb_match_frame_pairs call)")
    ap.add_argument("--chunks", type=int, default=80, help="launch groups per step: a step processes chunks x batch frames per GPU")
    ap.add_argument("--diagnose-e2e", action="store_true", help="per-rank split of the end-to-end time: copies only / kernels only / xfb_submit / host time")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=2,
                    help="independent xfb contexts (one CUDA stream each) the batches alternate over: frames are independent units, and two "
                         "batches in flight fill the SMs that one batch's latency-bound small-layer kernels leave idle")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
