#!/usr/bin/env python
"""bench.py -- masked-NST image-steps/s @640x400 VGG-19 (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch: a closure evaluation (VGG-19 forward to relu4_2,
Gram style + content losses, backward to the image) plus one L-BFGS iteration for every image of the batch
(BASELINE config[1]: 64 synthetic OpenEDS2019-shaped 640x400 eyes, iris-masked, random-init VGG-19,
Gram style loss, each image its own problem).  N > 1: one process per GPU (torchrun), each rank owns its
own batch (weak scaling), no collective on the inner loop; timing = max over ranks.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed.  `e2e`: the same job
through the public API `iris_b200.nst()` from pinned HOST tensors to HOST results (H2D/D2H inside the
timed region).  `roofline`: the tcgen05 conv kernel family, CUDA-event timed inside the timed region.
`cpu_baseline`: the CPU oracle (a port of the reference path, oracle/nst_oracle.py) on this box's host cores.
`--impl reference` times that CPU path alone with the same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "masked-NST image-steps/sec @640x400 VGG-19"
UNIT = "image-steps/s"
H, W = 640, 400
FLOPS_PER_IMAGE_STEP = 301.66e9  # fwd->relu4_2 142.44 G + dgrad 142.44 G + Gram fwd/bwd 16.78 G


def make_inputs(batch, seed0):
    """Iris-masked synthetic eyes: frame * ((seg == 2) & (frame <= 0.8)), replicated to 3 channels."""
    import numpy as np
    import torch

    from iris_b200 import synthetic

    def masked(seeds):
        frames, segs = synthetic.synthetic_batch(seeds, H, W)
        m = (segs == 2) & (frames <= np.float32(0.8))
        return torch.from_numpy(frames * m).repeat(1, 3, 1, 1).contiguous()

    c = masked([seed0 + i for i in range(batch)])
    s = masked([seed0 + 100000 + i for i in range(batch)])
    return c, s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.samples:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(steps, warmup, threads=None):
    """The reference path's CPU implementation (oracle port): B = 1 closure evaluations + L-BFGS on one
    640x400 masked eye (BASELINE config[0] shape), all host threads.  Returns image-steps/s."""
    import torch

    from oracle import nst_oracle as O

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    weights = O.random_vgg19_weights(0)
    c, s = make_inputs(1, 1)
    with torch.no_grad():
        _, c_feats, _ = O.vgg19_forward(c, weights, full=False)
        _, _, s_feats = O.vgg19_forward(s, weights, full=False)
        t_gram = [O.gram_matrix(t) for t in s_feats]
    x = c.clone()
    opt = O.LBFGS(x, lr=1.0)
    count = [0]
    t0 = [None]

    def closure():
        if count[0] == warmup:
            t0[0] = time.perf_counter()
        with torch.no_grad():
            x.clamp_(0, 1)
        cl, sl, g = O.nst_eval(x, c_feats, t_gram, weights, False, 1.0, 1e6)
        count[0] += 1
        return cl + 1e6 * sl, g.reshape(-1)

    total = warmup + steps
    while count[0] < total:
        opt.step(closure)
    dt = time.perf_counter() - t0[0]
    done = count[0] - warmup
    return done / dt, done, dt, threads


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    rate, done, dt, threads = cpu_reference_rate(steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "iris-masked Gatys NST, 640x400 synthetic eyes, random-init VGG-19, Gram style loss "
                               "(s_loss_weight 1e6), L-BFGS; reference arm: one image per step on the host CPU"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d closure evaluations + L-BFGS iterations of one 640x400 image (oracle/nst_oracle.py, "
                                   "torch CPU fp32, %d threads)" % (done, threads)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="isx", choices=["isx", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-features", action="store_true")
    ap.add_argument("--streams", type=int, default=1, help="sub-batches on separate CUDA streams (overlap L-BFGS and convs)")
    ap.add_argument("--history-bf16", action="store_true", help="opt-in: store the L-BFGS (s, y) history in bf16")
    ap.add_argument("--feature-images", type=int, default=512, help="images per GPU of the feature-extraction leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 40
        run_reference(args, rank)
        return
    if args.steps is None:
        args.steps = 300  # BASELINE config[1]: 300 steps
    if args.warmup < 3:
        args.warmup = 3

    import ctypes

    import torch
    import torch.distributed as dist

    import iris_b200
    from iris_b200 import _lib, pipelines, sharding

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sharding.init_from_env("nccl")
    lib = _lib.load()
    lib.isx_launch_count.restype = ctypes.c_ulonglong
    _lib.call("isx_device_check", local)

    B = args.batch
    K, Wm = args.steps, args.warmup
    vgg = iris_b200.VGG19(weights="random", seed=0)
    c_host, s_host = make_inputs(B, 1000 * rank + 1)
    c_host, s_host = c_host.pin_memory(), s_host.pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident leg
    with torch.cuda.device(dev), torch.no_grad():
        jkw = dict(clone_content=True, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6, lr=1.0, epochs=K + Wm + 40,
                   independent=True, history_dtype=torch.bfloat16 if args.history_bf16 else torch.float32)
        if args.streams > 1:
            job = pipelines.NstJobGroup(c_host.to(dev), s_host.to(dev), vgg, dev, streams=args.streams, **jkw)
            subjobs = job.jobs
        else:
            job = pipelines.NstJob(c_host.to(dev), s_host.to(dev), vgg, dev, **jkw)
            subjobs = [job]
        torch.cuda.synchronize()
        x0 = torch.cat([j.x for j in subjobs]).clone()
        for _ in range(Wm):
            job.tick()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        lib.isx_prof_enable(1)
        launches0 = lib.isx_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if args.streams > 1:
            job.fork(dev)
        for _ in range(K):
            job.tick()
        if args.streams > 1:
            job.join(dev)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = int(lib.isx_launch_count() - launches0)
        prof = (ctypes.c_double * 9)()
        _lib.call("isx_prof_collect", prof, 9)
        lib.isx_prof_enable(0)
        clocks = sampler.stop() if rank == 0 else None
        moved = float((torch.cat([j.x for j in subjobs]) - x0).abs().mean())
        loss_first = float(sum(j.hist_s[0].sum() for j in subjobs))
        loss_last = float(sum(j.hist_s[j.ticks - 1].sum() for j in subjobs))
        hist_slots = subjobs[0].cfg.history
        del job, subjobs
        torch.cuda.empty_cache()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max / 1e3)

    # ------------------------------------------------------------------ end-to-end leg (public API, host buffers)
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty_like(c_host).pin_memory()
        barrier()
        t0 = time.perf_counter()
        x, _, c_hist, s_hist = iris_b200.nst(c_host, s_host, BN_loss=False, c_loss_weight=1.0, s_loss_weight=1e6,
                                              epochs=K, vgg=vgg, use_tqdm=False, device=str(dev), independent=True,
                                              x_hist_stride=0, streams=args.streams,
                                              history_dtype=torch.bfloat16 if args.history_bf16 else torch.float32)
        x_host.copy_(x, non_blocking=True)  # result image back into pinned host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        evals = len(s_hist)
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_max = float(t.item())
        h2d = c_host.numel() * 4 + s_host.numel() * 4
        d2h = x_host.numel() * 4 + 2 * evals * B * 8
        e2e = {"value": world * B * evals / dt_max, "unit": UNIT, "h2d_bytes_per_step": h2d / evals,
               "d2h_bytes_per_step": d2h / evals, "evals": evals,
               "note": "iris_b200.nst() from pinned host tensors to host results: targets, %d evaluations, final image "
                       "and per-evaluation losses copied back; wall clock" % evals}
        del x
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ secondary metric: Gram-feature images/s
    # BASELINE config[2]: style features (mean/std + Gram upper triangles of relu1_1..relu4_1) of synthetic eyes for
    # the iris classifier, image list sharded contiguously over the ranks, ONE all-gather of the rows at the end.
    feat = None
    if not args.no_features:
        from iris_b200 import features

        n_per_gpu = args.feature_images
        n_total = n_per_gpu * world
        base, _ = __import__("iris_b200").synthetic.synthetic_batch(list(range(16)), H, W)   # 16 distinct eyes, tiled
        base = torch.from_numpy(base).pin_memory()

        class Tiled:
            """n_total eyes = the 16 synthetic frames repeated; a slice is served from pinned memory (cached per phase)."""
            def __init__(self):
                self.cache = {}

            def __len__(self):
                return n_total

            def __getitem__(self, sl):
                key = (sl.start % base.shape[0], sl.stop - sl.start)
                if key not in self.cache:
                    idx = torch.arange(sl.start, sl.stop) % base.shape[0]
                    self.cache[key] = base[idx].pin_memory()
                return self.cache[key]

        imgs = Tiled()
        vgg_feat = iris_b200.VGG19(content_layers=[], weights=vgg.host_weights)  # forward stops at relu4_1
        features.extract_features_sharded(vgg_feat, imgs, batch=32, device=dev)  # warm-up (workspaces, NCCL)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        rows = features.extract_features_sharded(vgg_feat, imgs, batch=32, device=dev)
        f1.record()
        barrier()
        tf = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        feat = {"metric": "Gram-feature images/sec @640x400 (relu1_1..relu4_1: mean/std + Gram upper triangles)",
                "value": n_total / (float(tf.item()) / 1e3), "unit": "images/s", "n_images": n_total,
                "feature_dim": int(rows.shape[1]), "all_gather_bytes": int(rows.numel() * 4),
                "flops_per_image": 131.96e9,
                "note": "host (pinned) -> device copies of the frames inside the timed region; one NCCL all-gather"}
        del rows
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline + baselines
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    if not peak_tf:
        peak_tf, peak_src = 1400.0, "B200_PROFILING.md fallback, sustained (of fallback)"
    # L-BFGS passes: algorithmic HBM bytes = history streamed twice (16*m*N) + 36*N per image-step, with
    # m = number of stored pairs at that tick (one pair per iteration, capped at the 100 slots)
    N_img = 3 * H * W
    esz = 2 if args.history_bf16 else 4
    lbfgs_bytes = 0.0
    for t in range(Wm, Wm + K):
        m_t = min(max(t - 1, 0), hist_slots)
        lbfgs_bytes += B * N_img * (4.0 * esz * m_t + 36.0)
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    lbfgs_gbs = lbfgs_bytes / (prof[7] / 1e3) / 1e9 if prof[7] > 0 else None
    # DRAM traffic of the conv family per launch, from the committed ncu capture of this command (profiles/)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")))
        if tr.get("batch") == B:
            traffic = tr["dram_bytes_per_conv_launch"]
    except Exception:
        pass
    conv_n, conv_ms, conv_flops = prof[0], prof[1], prof[2]
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    roofline = {
        "kernel": "conv family: conv_halo / conv_c64 / conv_tc / conv1_1 head+tail (tcgen05 implicit-GEMM conv fwd/dgrad/Gram-bwd)", "bound": "tensor",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None,
        # algorithmic DRAM bytes of the 19 conv launches of one evaluation (bf16 NHWC in + out, ReLU-mask and Gram-operand
        # reads of the dgrads): 587 MB per 640x400 image
        "traffic": traffic, "algorithmic_dram_bytes_per_launch": 587e6 * B / 19.0,
        "peak_source": peak_src, "launches": int(conv_n), "avg_launch_ms": conv_ms / max(conv_n, 1),
        "share_of_step": conv_ms / ms,
        "other": {"gram_tc_ms_share": prof[4] / ms, "lbfgs_ms_share": prof[7] / ms,
                  "gram_tflops": prof[5] / (prof[4] / 1e3) / 1e12 if prof[4] > 0 else None,
                  "lbfgs_hbm": {"bound": "hbm", "achieved": lbfgs_gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": lbfgs_gbs / hbm_peak if lbfgs_gbs else None,
                                "note": "lbfgs_dots + reduce + control + lbfgs_update, algorithmic bytes / CUDA-event time"}},
    }
    cpu = None
    if not args.no_cpu_baseline:
        rate, done, dtc, threads = cpu_reference_rate(40, 2)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d closure evaluations + L-BFGS iterations of ONE 640x400 masked eye (%.1f s; oracle/nst_oracle.py, "
                         "torch CPU fp32, %d threads)" % (done, dtc, threads)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "BASELINE config[1]: batch of %d synthetic OpenEDS2019-shaped 640x400 eyes per GPU, iris-masked, "
                               "3-channel, random-init VGG-19, Gram style loss (relu1_1..relu4_1) + content relu4_2, "
                               "alpha=1 beta=1e6, L-BFGS(lr=1, history 100), every image its own problem" % B,
                   "batch_per_gpu": B, "image": "3x%dx%d" % (H, W), "l2": "inputs larger than L2 (activations %.1f GB per step)"
                   % (B * 139e6 / 1e9), "history_slots": hist_slots, "history_dtype": "bf16" if args.history_bf16 else "f32", "streams": args.streams,
                   "flops_per_image_step": FLOPS_PER_IMAGE_STEP,
                   "model_tflops": value * FLOPS_PER_IMAGE_STEP / 1e12 / world},
        "e2e": e2e, "secondary": feat, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "sanity": {"image_moved_mae": moved, "s_loss_first": loss_first, "s_loss_last": loss_last},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
