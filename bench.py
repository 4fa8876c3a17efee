#!/usr/bin/env python
"""bench.py -- RA-LENet training-step throughput on B200 (BASELINE.json metric: ECG windows/s).

    python bench.py --gpus N --steps K --warmup W            # one rank per GPU under torchrun for N > 1
    python bench.py --impl reference ...                     # the reference algorithm's CPU path (oracle port)

Workload: BASELINE.json configs[1] "RA-LENet single-lead training step (fwd+bwd+Adam),
batch 256 x 1 x 512" realised reference-natively as 256 windows of 2 x 256 samples (the reference's conv1 is
hard-wired to 2 leads x 256 samples), model = transformer.ralenet(high_level_enhence=True), MSE loss,
Adam lr 1e-3, fp32.  One "step" = forward + loss + backward + (gradient all-reduce) + Adam on one batch of
synthetic windows.  N > 1: weak scaling, 256 windows per GPU, BatchNorm statistics and the flat gradient buffer
all-reduced over NCCL (equals the single-process step on the global batch).

Prints ONE JSON line (rank 0).  `value` = windows/s with inputs resident in HBM, CUDA-event timed, max over
ranks; `e2e` = the same through FusedTrainer.step_host with pinned HOST buffers (H2D of the batch and D2H of the
loss inside the timed region); `roofline` = the dominant kernel, timed live per launch with CUDA events
(ralenet_profile_*); `cpu_baseline` = the oracle port of the reference algorithm on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RA-LENet train-step throughput (fwd+bwd+Adam)"
UNIT = "windows/s"
PER_GPU_BATCH = 256
WORKLOAD = ("configs[1]: RA-LENet training step (fwd+bwd+Adam, MSE, lr 1e-3), {b} windows x 2 leads x 256 samples "
            "per GPU (= 256 x 1 x 512 samples; reference-native shape, SURVEY F1), transformer.ralenet("
            "high_level_enhence=True), fp32")


# ------------------------------------------------------------------------------------------------------
def algorithmic_work(label: str, B: int, L0: int = 256):
    """(flops, bytes) per launch of a kernel label 'name<C>' / 'wgrad_kernel<N,K>' -- 2*MAC of the matmul-shaped
    work only, and the fp32 activation tensors that must cross HBM."""
    name, _, tag = label.partition("<")
    tags = [int(t) for t in tag.rstrip(">").split(",")] if tag else []
    N = 8 * L0                       # floats per window per activation tensor
    name = {"ffn_fwd_cluster": "ffn_fwd_kernel", "ffn_bwd_cluster": "ffn_bwd_kernel",
            "ffn_fwd_umma": "ffn_fwd_kernel", "ffn_bwd_umma": "ffn_bwd_kernel"}.get(name, name)
    if name in ("attn_fwd_kernel", "attn_bwd_kernel", "ffn_fwd_kernel", "ffn_bwd_kernel"):
        C = tags[0]
        L = N // C
        if name == "attn_fwd_kernel":
            return B * (8 * L * C * C + 4 * L * L * C), B * 2 * N * 4
        if name == "attn_bwd_kernel":
            return B * (8 * L * C * C + 8 * L * L * C), B * 3 * N * 4
        if name == "ffn_fwd_kernel":
            return B * 16 * L * C * C, B * 2 * N * 4
        return B * 16 * L * C * C, B * 3 * N * 4
    if name in ("patch_fwd_kernel", "patch_bwd_kernel"):
        CN = tags[0]
        rows = N // CN
        return B * 2 * rows * CN * CN, B * (2 if name == "patch_fwd_kernel" else 3) * N * 4
    if name == "wgrad_kernel":
        Nn, K = tags
        # M is not in the label: every wgrad of this network has M*min-dim... recover from N*K and the stage
        return None, None
    return 0, None


def sample_clocks(stop_evt, path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except OSError:
        return
    stop_evt.wait()
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()


def parse_clocks(path, gpu_index):
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9 or f[0] != str(gpu_index):
                continue
            sm.append(float(f[1])); mx.append(float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except OSError:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    busy = sorted(sm)[len(sm) // 2:]          # upper half = samples under load
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def cpu_train_step_factory(B, threads):
    """the reference algorithm's CPU training step (oracle port): fwd, MSE, manual bwd, Adam on every tensor."""
    import torch
    from ecg_denoise_b200 import synth
    from oracle import ralenet_oracle as O
    from oracle import synth_weights as SW
    torch.set_num_threads(threads)
    sd = SW.make_state_dict("rw", 1, 2023)
    noisy, clean = synth.make_batch(B, 2, 256, seed=2023)
    x, t = torch.from_numpy(noisy), torch.from_numpy(clean)
    names = O.trainable_keys(sd)
    m = {n: torch.zeros_like(sd[n]) for n in names}
    v = {n: torch.zeros_like(sd[n]) for n in names}
    state = {"step": 0}

    def step():
        state["step"] += 1
        with torch.no_grad():
            out, ctx, new_stats = O.ralenet_fwd(x, sd, training=True)
            loss, dout = O.mse_loss_fwd_bwd(out, t)
            _, grads = O.ralenet_bwd(dout, ctx, sd)
            for n in names:
                O.adam_step(sd[n], grads[n], m[n], v[n], state["step"])
            sd["conv1.2.running_mean"], sd["conv1.2.running_var"] = new_stats[0], new_stats[1]
        return float(loss)
    return step


def time_cpu_baseline(budget_s=12.0, B=32):
    threads = os.cpu_count() or 1
    step = cpu_train_step_factory(B, threads)
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    return {"value": n * B / el, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} train steps (fwd+bwd+Adam) of the oracle port on {B} x 2 x 256 windows, "
                      f"{el:.1f} s, torch CPU fp32, {threads} threads"}


def run_reference(args):
    """--impl reference: the reference algorithm's own CPU path (the oracle port -- the reference is Python
    and does not exist on the GPU box) on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = 32
    step = cpu_train_step_factory(B, threads)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    val = args.steps * B / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(b=PER_GPU_BATCH),
                   "sample": f"each step = {B} of the {PER_GPU_BATCH} windows"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} train steps of the oracle port on {B} x 2 x 256 windows"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_inference(args):
    """configs[4]: batch-sharded inference on 48 synthetic 650000-sample 2-lead records (30 min @ 360 Hz) cut into
    non-overlapping 256-sample windows (the reference's cut, local_utils/local_utils.py:53): 2539 windows per record,
    121,872 in total, sharded by record across ranks with no communication.  One step = all records of this rank:
    z-norm + window gather, eval forward, stitch."""
    import torch
    import torch.distributed as dist
    from ecg_denoise_b200 import inference
    from ecg_denoise_b200.model import transformer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R_total, T = 48, 650000
    R = R_total // world
    torch.manual_seed(2023)
    model = transformer.ralenet(high_level_enhence=True)
    for rw in (model.rwattn1, model.rwattn2, model.rwattn3, model.rwattn4):
        rw.parameters_normalize()
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(100 + rank)
    host = torch.randn(R, 2, T, generator=g).pin_memory()
    recs = host.to(dev)
    nper = inference.windows_per_record(T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        y = inference.denoise_records(model, recs, batch=4096)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = inference.denoise_records(model, recs, batch=4096)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    out_host = torch.empty_like(host).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        recs.copy_(host, non_blocking=True)
        y = inference.denoise_records(model, recs, batch=4096)
        out_host.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        nwin = world * R * nper
        print(json.dumps({
            "metric": "RA-LENet inference throughput (records -> windows -> denoise -> stitch)",
            "value": nwin * args.steps / (float(t[0]) * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(t[0]) / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: {world * R} records x 2 leads x {T} samples -> {nwin} windows of 2 x 256, "
                                   "eval forward, batch-sharded by record, no communication"},
            "e2e": {"value": nwin * args.steps / (float(t[1]) * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": R * 2 * T * 4, "d2h_bytes_per_step": R * 2 * T * 4}}))


def run_train12(args):
    """configs[2]: 12-lead fine-tuning step (ralenet_12leads.newrale: 4 x Conv1d k13 around the frozen RA-LENet
    core, Transfer_learning.py:71-75).  64 LUDB-shaped records x 12 leads x 5000 samples per GPU, zero-padded to
    5120 and cut into 20 windows of 256 -> 1280 windows of 12 x 256 per step; MSE, torch.optim.Adam(lr 1e-3) over
    the 2,210 trainable parameters, through the drop-in nn.Module API (autograd nodes -> C ABI)."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from ecg_denoise_b200 import _lib
    from ecg_denoise_b200.model import ralenet_12leads
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(2023)
    core = ralenet_12leads.ralenet(high_level_enhence=True)
    for rw in (core.rwattn1, core.rwattn2, core.rwattn3, core.rwattn4):
        rw.parameters_normalize()
    model = ralenet_12leads.newrale(core).to(dev).train()
    if world > 1:       # SyncBN-equivalent statistics of the core's stem, as in the single-lead path
        core._plan.reduce_fn = lambda t: dist.all_reduce(t)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    R, T, NW = 64, 5000, 20
    g = torch.Generator().manual_seed(300 + rank)
    rec = torch.zeros(2, R, 12, NW * 256)
    rec[..., :T] = torch.randn(2, R, 12, T, generator=g)
    # (R, 12, 5120) -> (R * 20, 12, 256) windows
    hwin = rec.view(2, R, 12, NW, 256).permute(0, 1, 3, 2, 4).reshape(2, R * NW, 12, 256).contiguous().pin_memory()
    hx, ht = hwin[0], hwin[1]
    dx, dt = hx.to(dev), ht.to(dev)
    B = R * NW

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model(x), t)
        loss.backward()
        if world > 1:
            for p in params:
                dist.all_reduce(p.grad)
                p.grad.div_(world)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dx, dt)
    barrier()
    _lib.launch_count(reset=True)
    step(dx, dt)
    torch.cuda.synchronize()
    launches = _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(dx, dt)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sx, st_ = torch.empty_like(dx), torch.empty_like(dt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sx.copy_(hx, non_blocking=True)
        st_.copy_(ht, non_blocking=True)
        host_loss = step(sx, st_).item()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({
            "metric": "RA-LENet 12-lead fine-tune step throughput (fwd+bwd+Adam, frozen core)",
            "value": world * B * args.steps / (float(t[0]) * 1e-3), "unit": "12-lead windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[0]) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: newrale(ralenet(high_level_enhence=True)), {R} records x 12 x {T} per GPU "
                                   f"-> {B} windows of 12 x 256 per step (reference-native shape, SURVEY F2), MSE, "
                                   "torch.optim.Adam lr 1e-3 on the 4 Conv1d k13 layers, core frozen",
                       "global_batch": world * B, "parallelism": f"dp{world}" if world > 1 else "single"},
            "e2e": {"value": world * B * args.steps / (float(t[1]) * 1e-3), "unit": "12-lead windows/s",
                    "h2d_bytes_per_step": 2 * B * 12 * 256 * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "final_loss": float(host_loss)}))


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="windows per GPU")
    ap.add_argument("--graph", default="on", choices=["on", "off"], help="replay the step from a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--dump-kernels", default=None, help="write the per-kernel timing table (JSON) here")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "train12"],
                    help="train = configs[1] (default, the headline); infer = configs[4] record inference; "
                         "train12 = configs[2] 12-lead fine-tuning")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "infer":
        return run_inference(args)
    if args.workload == "train12":
        return run_train12(args)

    import torch
    import torch.distributed as dist
    from ecg_denoise_b200 import _lib, synth
    from ecg_denoise_b200.engine import FusedTrainer
    from ecg_denoise_b200.model import transformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    # random-init weights of the benchmark architecture (seed 2023, main.py:24); R-wave tables randomised with
    # the reference's own parameters_normalize() so the bias path does real work
    torch.manual_seed(2023)
    model = transformer.ralenet(high_level_enhence=True)
    for rw in (model.rwattn1, model.rwattn2, model.rwattn3, model.rwattn4):
        rw.parameters_normalize()
    model = model.to(dev)
    NB = 4                                       # rotating input batches (different windows every step)
    noisy, clean = synth.make_batch(NB * B, 2, 256, seed=2023 + 17 * rank)
    hx = torch.from_numpy(noisy).view(NB, B, 2, 256).pin_memory()
    ht = torch.from_numpy(clean).view(NB, B, 2, 256).pin_memory()
    dx, dt = hx.to(dev), ht.to(dev)

    trainer = FusedTrainer(model, lr=1e-3, use_graph=(args.graph == "on"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- launches per step (counted once, un-graphed) ---------------------------------------------------
    eager = FusedTrainer(model, lr=1e-3, use_graph=False)
    eager.step(dx[0], dt[0])
    torch.cuda.synchronize()
    _lib.launch_count(reset=True)
    eager.step(dx[1], dt[1])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count(reset=True)

    # ---- device-resident timing -------------------------------------------------------------------------
    for i in range(args.warmup):
        trainer.step(dx[i % NB], dt[i % NB])
    barrier()
    stop = threading.Event()
    clk_path = os.path.join(tempfile.gettempdir(), f"clocks_{os.getpid()}.csv")
    th = threading.Thread(target=sample_clocks, args=(stop, clk_path), daemon=True)
    if rank == 0:
        th.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss, rmse, snr, out = trainer.step(dx[i % NB], dt[i % NB])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    final_loss = float(loss.item())

    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss ---------------------------------------
    for i in range(3):
        trainer.step_host(hx[i % NB], ht[i % NB]).item()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for i in range(args.steps):
        l = trainer.step_host(hx[i % NB], ht[i % NB])
        host_loss = l.item()                      # device -> host read of the step's result, every step
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    if rank == 0:
        stop.set()
        th.join(timeout=10)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- per-kernel timing (roofline pass; un-graphed, events after every launch) ------------------------
    # (under torch.distributed every rank runs it -- the eager step contains the collectives -- and rank 0 reports)
    roofline, table = None, None
    if not args.no_profile:
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        agg = {}
        reps = 3
        for r in range(reps):
            torch.cuda.synchronize()
            _lib.check(lib.ralenet_profile_begin(ctypes.c_void_p(st)))
            eager.step(dx[r % NB], dt[r % NB])
            n_max = 8192
            labels = ctypes.create_string_buffer(n_max * 48)
            msb = (ctypes.c_float * n_max)()
            n = lib.ralenet_profile_end(labels, 48, msb, n_max)
            for i in range(n):
                lab = labels.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
                a = agg.setdefault(lab, [0, 0.0])
                a[0] += 1
                a[1] += msb[i]
        total = sum(a[1] for a in agg.values())
        table = sorted(((lab, a[0] // reps, a[1] / a[0], a[1] / reps) for lab, a in agg.items()), key=lambda r: -r[3])
        top, flops, nbytes = None, None, None
        for row in table:                      # dominant kernel with matmul-shaped algorithmic work
            f, nb = algorithmic_work(row[0], B)
            if f:
                top, flops, nbytes = row, f, nb
                break
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
        if top is not None:
            ach = flops / (top[2] * 1e-3) / 1e12
            traffic, traffic_src = None, None
            try:     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
                tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                e = tr.get(top[0].split(",")[0].rstrip(">") + ">")
                if e and B == PER_GPU_BATCH:
                    traffic, traffic_src = e["dram_bytes_per_launch"], "profiles/" + e["source"]
            except (OSError, ValueError, KeyError):
                pass
            roofline = {"kernel": top[0], "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": ach / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src,
                        "avg_launch_ms": top[2], "launches_per_step": top[1], "share_of_step": top[3] / (total / reps),
                        "note": "algorithmic FLOPs (2*MAC of the projections and of q.k^T / p.v and their adjoints) "
                                "over the live per-launch time.  The kernel runs split-precision (3xTF32) mma.sync "
                                "with head_dim 4: it is issue-bound on the softmax / split arithmetic around the "
                                "MMAs (ncu: tensor pipe ~28 % active, issue slots ~55 %), not HBM-bound; the "
                                "denominator is the bf16 dense peak the north star names",
                        "hbm_algorithmic_GBs": (nbytes / (top[2] * 1e-3) / 1e9) if nbytes else None}
        if args.dump_kernels and rank == 0:
            with open(args.dump_kernels, "w") as f:
                json.dump({"per_step_ms_sum": total / reps, "kernels": [
                    {"label": l, "launches_per_step": c, "avg_ms": a, "ms_per_step": s} for l, c, a, s in table]}, f,
                    indent=1)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = time_cpu_baseline()

    if rank == 0:
        clocks = parse_clocks(clk_path, local_rank)
        bytes_in = 2 * B * 2 * 256 * 4
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(b=B), "global_batch": world * B,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "cuda_graph": args.graph == "on",
                       "l2": "no explicit flush: each step streams 466 MB of saved activations (> 126 MB L2) and "
                             "rotates over 4 different input batches"},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bytes_in,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clocks, "final_loss": final_loss,
            "roofline": roofline, "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    trainer.close()
    eager.close()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
