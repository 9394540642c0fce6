#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: clips/sec, raw audio -> point clouds -> Set
Transformer logits (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

A step is one pass of the whole path over one batch of synthetic audio.  Workload at every N (weak
scaling, one process per GPU, no data-path collective): BASELINE config 2 -- FST Set Transformer
(ISAB+PMA, D=64, H=8, M=64) on spectrogram point clouds, batch 256 clips of 1 s @ 16 kHz per GPU,
n_fft 2048 / hop 1024 -> 16 frame clouds of 1025 (f, mag) points per clip = 4096 clouds per step.

`value`  : clips/s with the audio already resident in HBM (device-timed with CUDA events, max over ranks)
`e2e`    : clips/s through the public host-buffer interface (AudioSetPipeline.submit_host / wait_host: every
           step copies its pinned host audio H2D, runs the kernels and copies its logits D2H inside the timed
           region; the copy of step i+1 overlaps the kernels of step i, the host reads each step's logits)
`roofline`: dominant kernel (by device time inside the timed region, measured live with CUDA events on
           the launching stream) against the measured peak in MEASURED_PEAKS.json
`cpu_baseline`: the CPU oracle port of the reference path on the host cores, bounded sample
`--impl reference`: that same CPU port as its own arm (the reference is pure Python and cannot travel).
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
# keep stdout to the ONE JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "clips/sec audio->point-cloud->encoder"
UNIT = "clips/s"
CLIPS_PER_STEP = 256
N_SAMPLES = 16000
FS = 16000
N_FFT = 2048
WORKLOAD = ("FST Set Transformer (ISAB+PMA, D=64,H=8,M=64,C=10) on spectrogram point clouds: batch 256 "
            "synthetic 1 s 16 kHz clips per GPU, STFT 2048/1024 -> 16 frame clouds x 1025 (f,mag) points per clip")


def st_flops_per_cloud(n, d_in=2, D=64, M=64, S=1, C=10):
    """Algorithmic encoder FLOPs per cloud."""
    return 2 * (n * (3 * d_in * D + 8 * M * D + 7 * D * D + 2 * S * D) + 8 * M * D * D + 2 * S * D * D + S * D * C)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_path(steps, warmup, clips_per_sample, min_seconds=0.0):
    """The reference's CPU path (oracle port: librosa-0.8 STFT restatement -> ESC_pc clouds -> ST
    forward with the shipped FST weights, torch CPU with every host thread)."""
    import numpy as np
    import torch
    from oracle import pcaudio_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = orc.strip_module_prefix({k: torch.from_numpy(v) for k, v in
                                 np.load(os.path.join(ROOT, "tests", "golden", "fst_weights.npz")).items()})
    audio = orc.synth_audio(clips_per_sample, N_SAMPLES, FS, seed=202)

    def one_pass():
        with torch.no_grad():
            for c in range(clips_per_sample):
                clouds = torch.from_numpy(orc.clip_frame_clouds(audio[c], FS, N_FFT))
                orc.st_forward(w, clouds, 8)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    done = 0
    while done < steps or (time.perf_counter() - t0) < min_seconds:
        one_pass()
        done += 1
        if done >= steps and min_seconds == 0.0:
            break
    dt = time.perf_counter() - t0
    return {"value": clips_per_sample * done / dt, "seconds": dt, "passes": done, "cores": cores,
            "threads": torch.get_num_threads()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = 8          # bounded sample of the 256-clip batch: 8 clips = 128 frame clouds per step
    r = cpu_reference_path(args.steps, args.warmup, clips)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["passes"], "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / r["passes"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{clips} of the 256 clips per step (CPU-bounded)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": f"{clips} clips/step x {r['passes']} steps, torch CPU {r['threads']} threads"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ other BASELINE configs
def run_extra_configs(args, dev, rank, world, local):
    """BASELINE.json configs 1, 3, 4, 5 and the fp32 parity path of the headline, OUTSIDE the timed headline: each record is
    device-timed (CUDA events, >= 3 warm-up passes, max over ranks), weak scaling for the inference configs (the same
    per-GPU work on every rank, value = aggregate), strong scaling of one optimisation step for config 5 (global batch 256,
    one NCCL all-reduce of the flat gradient per step when N > 1).  Inputs exceed L2 or rotate through distinct batches."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import pcaudio_b200 as pca
    from pcaudio_b200 import parallel

    G = os.path.join(ROOT, "tests", "golden")
    recs = []

    def st_model(tag, d_in, precision):
        w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, f"{tag}_weights.npz")).items()}
        st = pca.ST(dim_input=d_in, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
        st.load_state_dict(w)
        return st.set_precision(precision)

    def timed(fn, steps, warmup=3):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def synth(n_clips, n_samples, seed):
        gen = torch.Generator(device=dev).manual_seed(seed + rank)
        t = torch.arange(n_samples, device=dev, dtype=torch.float32) / FS
        x = 0.05 * torch.randn(n_clips, n_samples, generator=gen, device=dev)
        for _j in range(4):
            a = torch.empty(n_clips, 1, device=dev).uniform_(0.05, 0.4, generator=gen)
            f = torch.empty(n_clips, 1, device=dev).uniform_(50.0, FS / 2 - 50.0, generator=gen)
            ph = torch.empty(n_clips, 1, device=dev).uniform_(0.0, 6.2831853, generator=gen)
            x = x + a * torch.sin(6.2831853 * f * t + ph)
        return x.clamp_(-1.0, 1.0).contiguous()

    def pipeline_record(name, tag, d_in, n_clips, n_samples, n_fft, mode, ntemp, top_k, precision, steps, pool_n, seed):
        st = st_model(tag, d_in, precision)
        cfg = pca.AudioConfig(sampling_rate=FS, window_size=n_fft, n_samples=n_samples, mode=mode, Ntemp=ntemp, top_k=top_k,
                              precision=precision)
        pipe = pca.AudioSetPipeline(st, cfg, dev)
        pool = [synth(n_clips, n_samples, seed + 17 * i) for i in range(pool_n)]
        sampler = ClockSampler(local)
        sampler.start()
        ms = timed(lambda i: pipe(pool[i % pool_n]), steps)
        clocks = sampler.stop()
        clouds = n_clips * pipe.clouds_per_clip
        fl = clouds * st_flops_per_cloud(pipe.points_per_cloud, d_in)
        recs.append({"config": name, "precision": precision, "clips_per_pass_per_gpu": n_clips, "passes": steps,
                     "clips_streamed_per_gpu": n_clips * steps, "clouds_per_clip": pipe.clouds_per_clip,
                     "points_per_cloud": pipe.points_per_cloud, "ms_per_pass": ms, "value": world * n_clips / ms * 1e3,
                     "unit": "clips/s", "clouds_per_s": world * clouds / ms * 1e3, "encoder_tflops_per_gpu": fl / ms / 1e9,
                     "n_gpus": world, "scaling": "weak", "clocks": clocks})
        del pool, pipe, st
        torch.cuda.empty_cache()

    # fp32 parity class of the headline workload (same 256-clip batches; the reference's own arithmetic is fp32)
    pipeline_record("2 (headline workload) fp32 parity path: FST 256 clips x 16 frame clouds x 1025 points", "fst", 2,
                    CLIPS_PER_STEP, N_SAMPLES, N_FFT, 2, 10, 0, "fp32", 3, 3, 202)
    # config 1: 3ST (temporal model, the reference's CPU-runnable case): batch 16 x 1 s, clip-as-cloud and reference chunking
    pipeline_record("1: 3ST batch 16 x 1 s clips, clip-as-cloud N=16384", "3st", 3, 16, 16000, 1024, 3, 32, 0, "bf16", 20, 8, 101)
    pipeline_record("1: 3ST batch 16 x 1 s clips, 3 chunk clouds x 5120 points", "3st", 3, 16, 16000, 1024, 3, 10, 0, "bf16", 20, 8, 101)
    # config 3: 3ST on 4 s clips, one GPU's shard of the data-parallel batch
    pipeline_record("3: 3ST 4 s clips, 12 chunk clouds x 5120 points, 64 clips per GPU", "3st", 3, 64, 64000, 1024, 3, 10, 0, "bf16", 6, 3, 303)
    pipeline_record("3: 3ST 4 s clips, clip-as-cloud N=64512, 32 clips per GPU", "3st", 3, 32, 64000, 1024, 3, 126, 0, "bf16", 6, 3, 303)
    # config 4: top-K sweep on the 16 384-point clip cloud (STFT + top-K kernel + encoder), 1e5 clips streamed per GPU and K
    for K in (256, 512, 1024, 2048, 4096, 8192):
        pipeline_record(f"4: sweep, 1 s clips -> top-K={K} of 16384 points (front end + top-K + encoder)", "3st", 3, 10000, 16000, 1024,
                        3, 32, K, "bf16", 10, 2, 404 + K)

    # config 5: main_pointcloud.SetTransformer training step (fwd + CE + bwd + all-reduce + Adam), global batch 256
    torch.manual_seed(505)
    model = pca.SetTransformer(dim_hidden=256, num_heads=4, num_inds=16).to(dev).train()
    N, d_in, Cc, Bg = 1000, 3, 40, 256
    lo, hi = parallel.shard_range(Bg, rank, world)
    gen = torch.Generator().manual_seed(505)
    pool = []
    for _i in range(4):
        X = torch.randn(Bg, N, d_in, generator=gen)
        X = (X - X.mean(dim=1, keepdim=True)) / X.std(dim=1, keepdim=True).clamp_min(1e-6)
        y = torch.randint(0, Cc, (Bg,), generator=gen)
        pool.append((X[lo:hi].to(dev), y[lo:hi].to(dev)))
    tr = pca.SetTrainer(model, lr=1e-3)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda i: tr.step(*pool[i % len(pool)]), 10)
    clocks = sampler.stop()
    flops_fwd = 2 * (N * (3 * 3 * 256 + 8 * 16 * 256 + 7 * 256 * 256 + 2 * 256) + 8 * 16 * 256 * 256 + 2 * 256 * 256 + 256 * 40)
    recs.append({"config": "5: main_pointcloud.SetTransformer(256, 4 heads, 16 inducing points) training step on ModelNet40-shaped "
                           "1000-point clouds: fwd + CE + bwd + NCCL all-reduce of the flat gradient + Adam",
                 "precision": "fp32", "global_batch": Bg, "local_batch": hi - lo, "ms_per_step": ms, "value": Bg / ms * 1e3,
                 "unit": "clouds/s", "achieved_tflops": 3 * flops_fwd * Bg / ms / 1e9, "n_gpus": world, "scaling": "strong",
                 "allreduce_bytes_per_step": int(sum(p.numel() for p in model.parameters()) * 4) if world > 1 else 0, "clocks": clocks})
    # the same model, inference forward (eval mode, no dropout) on the local shard: every linear layer and every attention
    # contraction as split-bf16 tcgen05 GEMMs (fp32 parity class)
    model.eval()
    Xs = [pl[0] for pl in pool]
    sampler = ClockSampler(local)
    sampler.start()
    with torch.no_grad():
        ms = timed(lambda i: model(Xs[i % len(Xs)]), 10)
    clocks = sampler.stop()
    recs.append({"config": "5 (inference): main_pointcloud.SetTransformer(256, 4 heads, 16 inducing points) forward on 1000-point clouds",
                 "precision": "fp32 (split-bf16 tensor-core GEMMs + attention)", "global_batch": Bg, "local_batch": hi - lo,
                 "ms_per_pass": ms, "value": Bg / ms * 1e3, "unit": "clouds/s", "achieved_tflops": flops_fwd * Bg / ms / 1e9,
                 "n_gpus": world, "scaling": "strong", "clocks": clocks})
    return recs


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import pcaudio_b200 as pca
    from pcaudio_b200 import _lib
    from pcaudio_b200.parallel import init_distributed

    rank, world, local = init_distributed("nccl")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)

    w = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests", "golden", "fst_weights.npz")).items()}
    model = pca.ST(dim_input=2, num_outputs=1, dim_output=10, num_inds=64, dim_hidden=64, num_heads=8).to(dev)
    model.load_state_dict(w)
    cfg = pca.AudioConfig(sampling_rate=FS, window_size=N_FFT, n_samples=N_SAMPLES, mode=2, precision=args.precision)
    pipe = pca.AudioSetPipeline(model, cfg, dev)
    clouds_per_step = CLIPS_PER_STEP * pipe.clouds_per_clip

    # synthetic audio generated on the device; a pool of distinct batches larger than L2 (126 MB) is
    # rotated so no step re-reads inputs that are still cached
    pool_n = 10
    gen = torch.Generator(device=dev).manual_seed(202 + rank)
    t = torch.arange(N_SAMPLES, device=dev, dtype=torch.float32) / FS
    pool = []
    for _ in range(pool_n):
        x = 0.05 * torch.randn(CLIPS_PER_STEP, N_SAMPLES, generator=gen, device=dev)
        for _j in range(4):
            a = torch.empty(CLIPS_PER_STEP, 1, device=dev).uniform_(0.05, 0.4, generator=gen)
            f = torch.empty(CLIPS_PER_STEP, 1, device=dev).uniform_(50.0, FS / 2 - 50.0, generator=gen)
            ph = torch.empty(CLIPS_PER_STEP, 1, device=dev).uniform_(0.0, 6.2831853, generator=gen)
            x = x + a * torch.sin(6.2831853 * f * t + ph)
        pool.append(x.clamp_(-1.0, 1.0).contiguous())
    host_pool = [p.cpu().pin_memory() for p in pool[:4]]
    host_out = torch.empty((clouds_per_step, 1, 10), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, tail=None):
        for i in range(warmup):
            fn(i)
        if tail is not None:
            tail()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if tail is not None:
            tail()                        # the last batch's logits are read inside the timed region too
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    last = {}

    def step_dev(i):
        last["logits"] = pipe(pool[i % pool_n])

    host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
    pending = []

    def step_host(i):
        # public pipelined host-buffer API: batch i is submitted (its H2D copy overlaps the kernels of batch i-1), then
        # the logits of batch i-1 are read on the host.  Every step copies its 16.4 MB in and its logits out.
        pending.append((pipe.submit_host(host_pool[i % len(host_pool)], host_outs[i & 1]), i & 1))
        if len(pending) > 1:
            t, slot = pending.pop(0)
            pipe.wait_host(t)
            last["host"] = float(host_outs[slot][0, 0, 0])

    def drain_host():
        while pending:
            t, slot = pending.pop(0)
            pipe.wait_host(t)
            last["host"] = float(host_outs[slot][0, 0, 0])

    # ---- device-resident timing (value)
    sampler = ClockSampler(local)
    sampler.start()                       # spans warm-up, the timed region and the e2e region (all under load)
    for i in range(args.warmup):
        step_dev(i)
    launches_warm = _lib.launch_count()
    ms = timed(step_dev, args.steps, 0)
    launches = _lib.launch_count() - launches_warm
    assert torch.isfinite(last["logits"]).all()
    value = world * CLIPS_PER_STEP * args.steps / (ms / 1e3)

    # ---- end-to-end timing through the host-buffer entry point (e2e)
    ms_e2e = timed(step_host, args.steps, min(3, args.warmup), tail=drain_host)
    e2e_value = world * CLIPS_PER_STEP * args.steps / (ms_e2e / 1e3)
    clocks = sampler.stop()

    # ---- per-kernel device times inside a (separately) timed region -> roofline of the dominant kernel
    _lib.profile_enable(True)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        step_dev(i)
    torch.cuda.synchronize(dev)
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    peaks = load_peaks()
    roofline, kernels = None, {}
    if rep:
        tot = sum(v["ms"] for v in rep.values())
        for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
            kernels[k] = {"launches_per_step": v["launches"] / prof_steps, "ms_per_step": v["ms"] / prof_steps,
                          "share": v["ms"] / tot}
        name, top = max(rep.items(), key=lambda kv: kv[1]["ms"])
        # DRAM traffic per launch of that kernel from the committed ncu --set full capture (profiles/), if any
        traffic, traffic_detail = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic_detail = json.load(open(tpath)).get(name)
            if isinstance(traffic_detail, dict):     # contract: a number (dram read + write bytes per launch) or null
                traffic = traffic_detail.get("dram_bytes_per_launch")
            else:
                traffic, traffic_detail = traffic_detail, None
        avg_s = top["ms"] / 1e3 / top["launches"]
        compute = top["flops"] > 0
        if compute:
            ach = top["flops"] / top["launches"] / avg_s / 1e12
            peak = peaks["tf_sustained"]
            roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": traffic,
                        "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)",
                        "avg_launch_ms": avg_s * 1e3, "share_of_step": top["ms"] / tot}
        else:
            ach = top["bytes"] / top["launches"] / avg_s / 1e9
            peak = peaks["hbm_gbs"]
            roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic, "peak_source": f"{peaks['source']} copy bandwidth",
                        "avg_launch_ms": avg_s * 1e3, "share_of_step": top["ms"] / tot}
    if roofline is not None and traffic_detail is not None:
        roofline["traffic_detail"] = traffic_detail
    # The softmax kernels are bound by the MUFU (ex2) pipe, not by the tensor pipe (head dim 8: 512 exponentials per
    # point and MAB against ~31 kFLOP): report that unit next to the contract's tensor roofline.  Peak = 16 ex2/clk/SM
    # (tools/microbench_mufu.cu, measured) x SMs x the SM clock sampled under load.
    if roofline is not None and roofline["kernel"] in ("mab_apply_tc_kernel", "mab_reduce_tc_kernel"):
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        ex2_per_launch = clouds_per_step * pipe.points_per_cloud * 8 * 64
        ach = ex2_per_launch / (roofline["avg_launch_ms"] / 1e3)
        roofline["binding_unit"] = {"unit": "MUFU ex2", "per_launch": ex2_per_launch, "achieved_per_s": ach,
                                    "peak_per_s": 16.0 * n_sm * clk, "frac": ach / (16.0 * n_sm * clk),
                                    "peak_source": "16 ex2/clk/SM measured (tools/microbench_mufu.cu) x SMs x sampled SM clock"}
    enc_flops = clouds_per_step * st_flops_per_cloud(pipe.points_per_cloud)
    whole = {"algorithmic_tflop_per_step": enc_flops / 1e12,
             "achieved_tflops": world * enc_flops / (ms / args.steps / 1e3) / 1e12 / world,
             "frac_of_bf16_sustained": enc_flops / (ms / args.steps / 1e3) / 1e12 / peaks["tf_sustained"]}

    extra = None
    if not args.no_extra:
        del pool, host_pool
        torch.cuda.empty_cache()
        extra = run_extra_configs(args, dev, rank, world, local)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_path(1, 1, 8, min_seconds=10.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"8 clips (128 frame clouds) per pass, {r['passes']} passes in {r['seconds']:.1f} s, "
                         f"torch CPU {r['threads']} threads"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step_per_gpu": CLIPS_PER_STEP,
                       "clouds_per_step_per_gpu": clouds_per_step, "points_per_cloud": pipe.points_per_cloud,
                       "encoder_precision": args.precision, "parallelism": f"dp{world} (clips sharded, no collective)",
                       "l2": f"inputs rotate through a pool of {pool_n} distinct batches ({pool_n * CLIPS_PER_STEP * N_SAMPLES * 4 / 1e6:.0f} MB > 126 MB L2); "
                             "per-step intermediates exceed L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": CLIPS_PER_STEP * pipe.h2d_bytes_per_clip,
                    "d2h_bytes_per_step": CLIPS_PER_STEP * pipe.d2h_bytes_per_clip},
            "gpu_launches": int(launches),
            "roofline": roofline, "whole_step": whole, "kernels": kernels, "cpu_baseline": cpu, "extra_configs": extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PCA_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs records (BASELINE configs 1, 3, 4, 5 and the fp32 headline)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                                   # timing rule: W >= 3 (both arms use the same count)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
