#!/usr/bin/env python
"""Benchmark of the CTC hot path (BASELINE.json metric: CTC fwd+bwd utterances/s & HBM GB/s % peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One "step" = one forward+backward pass of CTCLoss(reduce=True, size_average=True,
after_logsoftmax=False) over one synthetic batch of the named workload (default BASELINE
configs[1]: LibriSpeech-shaped B=64, T=400, V=29, L<=200, fp32), i.e. the reference call stack
modules/ctc_loss.py:25-57 -> functions/forward_backward.py:6-35 -> src/losses/*.cpp.

Our arm (default):
  value      whole-job utterances/s with the inputs resident in HBM, K steps timed with CUDA events
             on the launching stream between barriers, max over ranks;
  e2e        the same metric through the reference-facing engine call on HOST (pinned) buffers:
             CTCLossEngine.compute() -> e2e_ctc_engine_loss_host() copies the logits in, runs the
             kernels, copies losses + gradient back -- all inside the timed region;
  roofline   the dominant kernel (the alpha/beta lattice kernel) timed live with CUDA events by the
             library's profile hooks: algorithmic bytes (logits read + gradient write)
             per launch / mean launch time, against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  the reference CPU engine (oracle/_ref, else the C port) on this box's host cores,
             rank 0 / N=1 only, bounded sample.
N > 1 (under torchrun): every rank owns its own batch (weak scaling, no data-path collective); the
only exchange is the 16-byte all-reduce of {loss sum, count} (end2end_b200.distributed).

--impl reference times the reference's own CPU implementation on the same workload and prints the
same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 8))

import torch  # noqa: E402

WORKLOADS = {
    # name: (B, T, V, Lmin, Lmax, seed, dtype, full_length, description)
    "c1": (4, 50, 28, 10, 29, 0, "f32", True, "README CTCLoss example B=4 T=50 V=28 L 10-29 fp32"),
    "c2": (64, 400, 29, 100, 200, 1, "f32", False, "LibriSpeech-shaped char CTC B=64 T=400 V=29 L 100-200 fp32"),
    "c3": (1024, 128, 96, 20, 40, 2, "bf16", False, "OCR lines B=1024 T=128 V=96 L 20-40 bf16"),
    "c4": (128, 250, 1024, 40, 80, 3, "f32", False, "subword CTC B=128 T=250 V=1024 L 40-80 fp32"),
    "c5": (2048, 1600, 29, 300, 600, 4, "f32", False, "long-form ASR B=2048 T=1600 V=29 L 300-600 fp32"),
}
L2_BYTES = 126 * 2 ** 20


def make_inputs(B, T, V, Lmin, Lmax, seed, dtype, full_length):
    """SURVEY.md 8(d) draw order: logits, target lengths, targets (blank 0 never a target), frames."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, V, generator=g).to(dtype)
    tl = torch.randint(Lmin, Lmax + 1, (B,), generator=g)
    tg = torch.randint(1, V, (B, Lmax), generator=g)
    ll = torch.full((B,), T, dtype=torch.int64) if full_length else torch.randint(3 * T // 4, T + 1, (B,), generator=g)
    return x, tg, ll, tl


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload):
    """dram bytes/launch of the dominant kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload, {}).get("lattice_dram_bytes_per_launch")
    return None


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def reference_step_fn(x, tg, ll, tl):
    """The reference's CPU path for one step (module forward + backward), driven as the reference's
    Python drives its engine (oracle.ctc_loss_module restates those lines)."""
    import oracle
    eng = oracle.engine(0, prefer="reference")

    def step():
        leaf = x.detach().clone().requires_grad_()
        loss = oracle.ctc_loss_module(eng, leaf, tg, ll, tl, reduce=True, size_average=True, after_logsoftmax=False)
        loss.backward()
        return float(loss)
    return step, eng.kind


def run_reference(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return
    B, T, V, Lmin, Lmax, seed, dt, full, desc = WORKLOADS[wl]
    sample_B = min(B, args.ref_batch)       # the reference keeps ~46 MB of fp64 lattices live per long utterance
    x, tg, ll, tl = make_inputs(sample_B, T, V, Lmin, Lmax, seed, torch.float32, full)
    if dt == "bf16":
        x = x.to(torch.bfloat16).float()      # the oracle for bf16 logits is the reference on logits.float()
    step, kind = reference_step_fn(x, tg, ll, tl)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt_s = time.perf_counter() - t0
    value = sample_B * args.steps / dt_s
    cores = os.cpu_count()
    sample = "%d utterances of %s per step, %d steps, %d host threads (OMP_NUM_THREADS=%s; the engine spawns one thread per utterance)" % (
        sample_B, wl, args.steps, cores, os.environ.get("OMP_NUM_THREADS"))
    print(json.dumps({
        "impl": "reference", "metric": "ctc_fwd_bwd_utterances_per_s", "value": value, "unit": "utterances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (wl, desc), "batch_per_step": sample_B, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "utterances/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, wl):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CTC engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # The only collective of this path is a scalar all-reduce: NVLink SHARP multicast buys nothing, and its
        # set-up makes a TWO-rank communicator on a larger NVSwitch box 2-3x slower per step (measured:
        # profiles/r01/scale_r01e.md).  NCCL caches the setting at the first communicator, so it is set here.
        if world == 2:      # with 4 or 8 ranks the in-switch reduction helps (8 ranks: 0.167 vs 0.219 ms/step)
            os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        dist.init_process_group("nccl", device_id=dev)
    from end2end_b200 import CTCLoss, CTCLossEngine, _lib
    from end2end_b200.distributed import ShardedCTCLoss

    B, T, V, Lmin, Lmax, seed, dt, full, desc = WORKLOADS[wl]
    if args.batch:
        B = args.batch
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
    esize = 2 if dt == "bf16" else 4
    batch_bytes = B * T * V * esize
    # distinct batches so that the streamed inputs exceed L2 between reuses (no flush needed)
    n_rot = max(2, min(args.max_rotate, -(-2 * L2_BYTES // batch_bytes)))
    if batch_bytes * n_rot > 24 * 2 ** 30:
        n_rot = max(1, (24 * 2 ** 30) // batch_bytes)
    batches = []
    for i in range(n_rot):
        x, tg, ll, tl = make_inputs(B, T, V, Lmin, Lmax, seed + 1000 * i + 7919 * rank, dtype, full)
        batches.append((x.to(dev).requires_grad_(), tg.to(dev), ll.to(dev), tl.to(dev)))
    if world > 1:
        crit = ShardedCTCLoss(reduce=True, size_average=True, after_logsoftmax=False, global_batch=B * world)
    else:
        crit = CTCLoss(reduce=True, size_average=True, after_logsoftmax=False)

    def step(i):
        x, tg, ll, tl = batches[i % n_rot]
        x.grad = None
        loss = crit(x, tg, ll, tl)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # warm-up: at least W steps, and at least one pass over the rotating batches so that every batch's gradient
    # buffer exists before the timed region (a first-touch cudaMalloc inside it would time the allocator)
    for i in range(max(3, args.warmup, n_rot)):
        step(i)
    barrier()
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    _lib.profile_enable(False)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    barrier()
    w1 = time.perf_counter()
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(loss).item(), "non-finite loss in the timed region"
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(t_ms.item())
    clocks = sampler.summary(w0, w1) if sampler else None

    # ---- per-kernel device times, live, with the library's event hooks (separate short pass) ----
    _lib.profile_enable(True)
    _lib.profile_read()
    psteps = min(args.steps, 20)
    for i in range(psteps):
        step(i)
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    kern = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items() if v[1]}
    dominant = max(kern, key=kern.get)
    alg_bytes = B * T * V * 2 * esize                 # logits read + gradient write per launch
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kern[dominant] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": recorded_traffic(wl), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern,
                "step_frac": alg_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the engine call on HOST buffers (pinned), copies inside the timing ----
    eng = CTCLossEngine(0)
    hx, htg, hll, htl = make_inputs(B, T, V, Lmin, Lmax, seed + 7919 * rank, dtype, full)
    hx = hx.pin_memory()
    for _ in range(3):
        eng.compute(hx, htg, hll, htl, from_logits=True)
    barrier()
    esteps = min(args.steps, 50)
    t0 = time.perf_counter()
    for _ in range(esteps):
        hl, hg = eng.compute(hx, htg, hll, htl, from_logits=True)
        float(hl[0])                                   # the result is on the host when compute() returns
    e2e_s = time.perf_counter() - t0
    h2d, d2h = eng.last_host_traffic()
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_e, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * esteps / float(t_e.item())

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_B = min(B, args.ref_batch)
        cx, ctg, cll, ctl = make_inputs(sample_B, T, V, Lmin, Lmax, seed, torch.float32, full)
        if dt == "bf16":
            cx = cx.to(torch.bfloat16).float()
        cstep, kind = reference_step_fn(cx, ctg, cll, ctl)
        cstep()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.cpu_seconds:
            cstep(); n += 1
        dt_s = time.perf_counter() - t0
        cpu_baseline = {"value": sample_B * n / dt_s, "unit": "utterances/s", "cores": os.cpu_count(), "kind": kind,
                        "sample": "%d utterances of %s x %d passes in %.1f s, one host thread per utterance "
                                  "(the reference's pool) on %d cores" % (sample_B, wl, n, dt_s, os.cpu_count())}

    if rank == 0:
        print(json.dumps({
            "metric": "ctc_fwd_bwd_utterances_per_s", "value": world * B * args.steps / (ms * 1e-3),
            "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (wl, desc), "batch_per_gpu": B, "io_dtype": dt,
                       "loss": "CTCLoss(reduce=True, size_average=True, after_logsoftmax=False)",
                       "l2": "rotating %d distinct resident batches (%.0f MB streamed between reuses > 126 MB L2)"
                             % (n_rot, n_rot * batch_bytes * 2 / 2 ** 20),
                       "parallelism": "batch-sharded x%d, 16-byte loss all-reduce" % world if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "utterances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "CTCLossEngine.compute(host pinned tensors, from_logits=True) -> e2e_ctc_engine_loss_host"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's batch size (experiments)")
    ap.add_argument("--max-rotate", type=int, default=128)
    ap.add_argument("--ref-batch", type=int, default=64, help="utterances per reference-CPU step (bounded sample)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _, _, world = dist_env()
    if args.gpus != world and world > 1:
        args.gpus = world
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
