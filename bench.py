#!/usr/bin/env python
"""Benchmark of the hot path (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on rank 0.  `value` = whole-job VAE training throughput (sequences/s) with the batch
resident in HBM; `e2e` = the same step driven from pinned HOST buffers (H2D of the batch + D2H of the
per-sample losses inside the timed region).  `roofline` describes the dominant kernel of the step,
`rasteriser` the K1 note-event rasteriser on BASELINE config 2 (roll GB/s), `cpu_baseline` the oracle
(CPU restatement of the reference step) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# stdout carries exactly one line (the JSON): everything else any library prints (NCCL_DEBUG banners, warnings) goes to
# stderr through this dup2; NCCL_DEBUG itself is left as the caller set it
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "vae_train_sequences_per_sec"
UNIT = "sequences/s"
HEADLINE = "bf16p3f"     # precision mode of the headline line (engine.PRECISIONS); its parity: tests/test_parity_bench_gpu.py


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="sequences per GPU per step")
    ap.add_argument("--seq-len", type=int, default=64)
    ap.add_argument("--dec-type", default="lstm", choices=["lstm", "transformer"])
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--precision", default=HEADLINE, choices=["fp32", "fp32x3", "tf32x3f", "bf16x3f", "bf16p3f", "tf32", "bf16"],
                    help="engine precision mode (musicstyletransfer_b200/engine.py PRECISIONS): bf16p3f (default) = fp32-class forward "
                         "(GEMM operands as bf16 hi + lo planes, compensated attention scores), TF32 backward; tf32x3f = the same "
                         "with 3xTF32 GEMMs; tf32 = every product single-pass TF32; "
                         "fp32x3 = strict fp32 on the tensor cores; fp32 = exact FFMA; bf16 = BASELINE config 4")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="rows per oracle step of the CPU arm (0 = the GPU arm's per-GPU batch, i.e. the same step)")
    ap.add_argument("--mode", default="train", choices=["train", "sweep", "style"],
                    help="train: the contract line (default); sweep: BASELINE config 3 batch / sequence-length sweep, one JSON "
                         "line per point; style: BASELINE config 5 style-transfer inference")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (falls back to "
                         "nccl if symmetric memory cannot be set up), nccl = ncclAllReduce of the gradient arena + full Adam")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of a step from the host instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", action="store_true", help="also print the step's per-entry-point time table (CUDA events "
                    "around every libmsx call of three eager steps) to stderr")
    ap.add_argument("--gemm-table", action="store_true", help="also print a per-shape table of the step's GEMM launches to stderr")
    ap.add_argument("--no-raster", action="store_true")
    ap.add_argument("--no-variants", "--no-bf16-variant", dest="no_variants", action="store_true",
                    help="skip the sub-object legs the default tf32 line carries: bf16 variant (BASELINE config 4), strict-fp32 "
                         "(3xTF32) variant, strong-scaling points, the batch-32 step")
    return ap.parse_args()


def workload_name(args):
    return ("VarAutoEncoder train step (fwd+bwd+Adam), scripts/train-vae.sh model: enc 2x256/8h, Z=256, dec %s 1x128, "
            "dropout %.1f, B=%d per GPU, L=%d (T=%d), synthetic 4/4 token rows"
            % (args.dec_type, args.dropout, args.batch, args.seq_len, args.seq_len + 1))


PRECISION_NOTE = {
    "fp32": "fp32 storage, GEMMs fp32 FFMA",
    "fp32x3": "fp32 storage, GEMMs on tcgen05 with 3xTF32 operand splitting (hi/lo, three MMAs per k-block, fp32 accumulate): "
              "fp32-equivalent products; attention / LSTM on the exact FFMA kernels",
    "tf32x3f": "fp32 storage; FORWARD GEMMs 3xTF32 on tcgen05 (fp32-equivalent products) and attention scores compensated the "
               "same way -> loss / KL / latent means within 1e-3 of the fp32 oracle with 10x margin (measured 1.0e-4); BACKWARD "
               "GEMMs, attention and the LSTM recurrence single-pass TF32 (fp32 accumulate)",
    "bf16p3f": "fp32 storage; FORWARD encoder GEMMs on msx_gemm_tc_p3: both operands arrive as bf16 hi / lo planes written by "
               "the producing kernels (embedding, LayerNorm, attention, FF1 epilogue; weights split once per step), three walks "
               "hi*hi + hi*lo + lo*hi on tcgen05 kind::f16, ~2^-17 per product, attention scores compensated the same way; "
               "backward GEMMs, LSTM recurrence single-pass TF32 (two weight gradients read the bf16 hi plane)",
    "bf16x3f": "fp32 storage; FORWARD encoder GEMMs with bf16x3 products on tcgen05 kind::f16 (operands split into bf16 hi + lo "
               "inside the kernel, ~2^-17 per product) and attention scores 3xTF32 -> loss / KL / latent means within 1e-3 of the "
               "fp32 oracle with margin; BACKWARD GEMMs, decoder GEMMs, attention and the LSTM recurrence single-pass TF32",
    "tf32": "fp32 storage, every tensor-core product single-pass TF32 (fp32 accumulate); latent means deviate 1.3e-3 from the "
            "fp32 oracle at the bench shape (over the 1e-3 bar, hence a variant and not the headline)",
    "bf16": "fp32 master weights / residual stream / LN / softmax / losses / Adam, Transformer-layer GEMM operands bf16 in HBM on "
            "tcgen05 kind::f16 (fp32 accumulate), other GEMMs TF32",
}


def config_dict(args, world):
    """The workload description both arms print (same dict -> the driver can pair the lines)."""
    return {"workload": workload_name(args), "global_batch": args.batch * world, "batch_per_gpu": args.batch,
            "seq_len": args.seq_len, "parallelism": "dp%d" % world}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  nvidia-smi takes up to a
    second to start on an 8-GPU box, so the sampler is started before the warm-up and the samples are filtered to the
    timed region by their timestamps (mark_start / mark_end)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.06)                    # let the sample that covers the end of the region land
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), [n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or r[0]) + 0.05]
        use = inside or rows[-3:]           # clock skew / too short a region: fall back to the last samples taken under load
        if use:
            sm = sorted(r[1] for r in use)
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[2] for r in use),
                   "reasons": sorted({n for r in use for n in r[3]}), "samples": len(use), "in_region": bool(inside)}
        return out


# ------------------------------------------------------------------------------------- reference arm
def oracle_step_rate(args, steps, warmup, threads=None):
    """Times the oracle (CPU restatement of Trainer._step, trainer.py:155-179) on a bounded sample."""
    import torch
    from musicstyletransfer_b200 import synth
    from oracle import model as om
    # torchrun exports OMP_NUM_THREADS=1: set the thread count explicitly so the CPU arm always uses every host core
    torch.set_num_threads(threads or os.cpu_count() or 1)
    cfg = om.Cfg(dec_type=args.dec_type, enc_dropout=args.dropout, dec_dropout=args.dropout)
    p = om.init_params(cfg, seed=0)
    opt = om.Adam(p, lr=3e-4, clip_gradient=1.0)
    Bc = args.cpu_batch or args.batch
    tok, lens, cls, lab = synth.token_rows_4_4(Bc * 2, args.seq_len, seed=1)
    t = lambda a: torch.from_numpy(a).float()
    batches = [(t(tok[i * Bc:(i + 1) * Bc]), t(lens[i * Bc:(i + 1) * Bc]), t(cls[i * Bc:(i + 1) * Bc]),
                t(lab[i * Bc:(i + 1) * Bc])) for i in range(2)]
    g = torch.Generator().manual_seed(0)
    times = []
    for i in range(warmup + steps):
        tk, ln, cl, lb = batches[i % 2]
        eps = torch.randn(Bc, cfg.latent, generator=g)
        t0 = time.perf_counter()
        om.train_step(cfg, p, opt, tk, ln, cl, lb, eps, masks="random" if args.dropout > 0 else None)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return Bc * len(times) / total, total / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 1))
    rate, per_step, cores = oracle_step_rate(args, steps, warmup)
    Bc = args.cpu_batch or args.batch
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world),
        "note": "reference = CPU oracle (the reference's MXNet 1.3 stack cannot run here, SURVEY.md §8(c)), all host "
                "threads on rank 0; each step is one %d-row batch of the workload (the GPU arm's per-GPU step)" % Bc,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d oracle steps of %d rows (torch-CPU fp32, %d host threads)" % (steps, Bc, cores)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------- our arm
def bench_rasteriser(peaks):
    import torch
    from musicstyletransfer_b200 import featurise, synth
    dtick, pitch, vel, offs = synth.note_events()
    dev = "cuda"
    d = [torch.from_numpy(a).to(dev) for a in (dtick, pitch, vel, offs)]
    n = offs.size - 1
    # two output sets (2 x 277 MB) alternate so that consecutive launches never hit the 126 MB L2
    outs = [(torch.empty((n, 65), dtype=torch.int32, device=dev), torch.empty((n, 64, 128), dtype=torch.uint8, device=dev),
             torch.empty((n,), dtype=torch.int32, device=dev)) for _ in range(2)]
    for i in range(4):
        featurise.rasterize(*d, out=outs[i % 2])
    torch.cuda.synchronize()
    reps = 20
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps):
        evs[i][0].record()
        featurise.rasterize(*d, out=outs[i % 2])
        evs[i][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    med = ms[len(ms) // 2]
    E = dtick.size
    bytes_alg = 6 * E + 4 * (n + 1) + n * 64 * 128 + n * 65 * 4 + n * 4
    gbs = bytes_alg / (med * 1e-3) / 1e9
    cpu = None
    try:        # CPU baseline of the same workload: the oracle's C restatement on all host cores
        from oracle import raster_c
        if raster_c.available():
            th = os.cpu_count() or 1
            raster_c.rasterize_batch(dtick, pitch, vel, offs, threads=th)
            t0 = time.perf_counter()
            for _ in range(5):
                raster_c.rasterize_batch(dtick, pitch, vel, offs, threads=th)
            sec = (time.perf_counter() - t0) / 5
            cpu = {"value": bytes_alg / sec / 1e9, "unit": "GB/s", "cores": th, "kind": "port",
                   "sample": "5 full passes of config 2 through oracle/raster.c (OpenMP, incl. output allocation)",
                   "events_per_s": E / sec}
    except Exception as exc:        # the baseline is informative only
        cpu = {"error": str(exc)}
    return {"cpu_baseline": cpu, "workload": "BASELINE config 2: 1,048,576 note events, 32768 sequences -> tokens int32[N,65] + roll uint8[N,64,128]",
            "ms": med, "events_per_s": E / (med * 1e-3), "roll_GBps": gbs,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": bytes_alg,
                         "peak_source": peaks["src"]}}


def bench_from_midi(args, cfg, dev):
    """End-to-end featurisation leg (single GPU): Standard MIDI File bytes -> C++ parser (host) -> K1 token streams ->
    A2 rows (device) -> one pass of graph-replayed train steps over the shuffled rows, batches gathered on the device.
    The MIDI bytes are synthetic (built before the clock starts); everything after them is inside the timed region."""
    import numpy as np
    import torch
    from musicstyletransfer_b200 import featurise, synth
    from musicstyletransfer_b200.engine import VAEEngine
    B, L = args.batch, args.seq_len
    blobs, classes = synth.midi_files(n_files=512, ev_per_file=2048, seed=3)
    order_c = np.argsort(np.asarray(classes), kind="stable")                    # tracks grouped by class (data.py:137-155)
    blobs = [blobs[i] for i in order_c]
    cls_sorted = np.asarray(classes, np.int32)[order_c]
    class_start = np.searchsorted(cls_sorted, np.arange(cfg.num_classes + 1)).astype(np.int32)
    eng = VAEEngine(cfg, dev, seed=0, precision=args.precision)
    train = eng.train_step if args.no_graph else eng.train_step_graphed
    res = {}
    passes = []
    import gc
    gc.collect()
    torch.cuda.empty_cache()                                                    # the variants before this leg leave a fragmented cache
    for rep in range(4):                                                        # pass 0 warms up (graph capture, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        soas = [featurise.parse_smf(b)[1][0] for b in blobs]
        t1 = time.perf_counter()
        tokens, n_tokens = featurise.tokenize_tracks_device(soas, dev)
        rows = featurise.build_rows(tokens, n_tokens, torch.from_numpy(cls_sorted).to(dev), torch.from_numpy(class_start).to(dev), L)
        R = int(rows[0].shape[0])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        perm = np.random.RandomState(rep).permutation(R)
        n_batches = R // B
        for i in range(n_batches):
            idx = torch.from_numpy(perm[i * B:(i + 1) * B].astype(np.int32)).to(dev, non_blocking=True)
            tk, lb, cl, ln = featurise.gather_batch(rows, idx, L + 1)
            train(tk, ln, cl, lb, kl_weight=1.0, global_batch=B, lr=3e-4, clip_gradient=1.0)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        res = {"value": n_batches * B / (t3 - t0), "unit": UNIT, "rows": R, "batches": n_batches, "batch": B,
               "midi_bytes": int(sum(len(b) for b in blobs)), "note_events": int(sum(len(s[0]) for s in soas)),
               "stage_ms": {"parse_smf_host": (t1 - t0) * 1e3, "tokenise_and_rows_device": (t2 - t1) * 1e3,
                            "train_steps": (t3 - t2) * 1e3},
               "what": "512 synthetic single-track .mid files (bytes in host memory) -> msx_smf_parse -> msx_rasterize -> "
                       "msx_rows_plan/build -> msx_rows_gather_batch + graph-replayed train steps over every full batch; wall "
                       "clock from the first parsed byte to the last step; median of three passes after one warm-up pass"}
        if rep > 0:
            passes.append(res)
    res = sorted(passes, key=lambda r: r["value"])[len(passes) // 2]
    res["passes_sequences_per_s"] = [p_["value"] for p_ in passes]
    eng._graphs.clear()
    del eng
    torch.cuda.synchronize()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from musicstyletransfer_b200 import lib, synth
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    peaks = measured_peaks()

    cfg = VAEConfig(dec_type=args.dec_type, enc_dropout=args.dropout, dec_dropout=args.dropout)
    eng = VAEEngine(cfg, dev, seed=0, precision=args.precision)
    B, L = args.batch, args.seq_len
    T = L + 1
    n_batches = 4
    tok, lens, cls, lab = synth.token_rows_4_4(B * n_batches, L, seed=100 + rank)
    host = []
    for i in range(n_batches):
        sl = slice(i * B, (i + 1) * B)
        host.append(tuple(torch.from_numpy(a[sl].copy()).pin_memory() for a in (tok, lens, cls, lab)))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    gbatch = B * world

    def allreduce(g):
        dist.all_reduce(g)

    ar = allreduce if world > 1 else None
    dp_exchange = "none" if world == 1 else "nccl all-reduce + Adam on every rank"
    if world > 1 and args.dp == "peer":
        try:
            eng.enable_peer_optimizer()
            ar = "peer"
            dp_exchange = "fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory (msx_adam_nvlink_step)"
        except Exception as exc:        # symmetric memory unavailable: keep the NCCL exchange, say so
            print("bench: peer optimiser unavailable (%s: %s), using NCCL all-reduce" % (type(exc).__name__, exc), file=sys.stderr)

    train = eng.train_step if args.no_graph else eng.train_step_graphed

    def step_resident(i):
        tk, ln, cl, lb = resident[i % n_batches]
        return train(tk, ln, cl, lb, kl_weight=1.0, global_batch=gbatch, lr=3e-4, clip_gradient=1.0, allreduce=ar)

    def step_eager(i):
        tk, ln, cl, lb = resident[i % n_batches]
        return eng.train_step(tk, ln, cl, lb, kl_weight=1.0, global_batch=gbatch, lr=3e-4, clip_gradient=1.0, allreduce=ar)

    stage = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    loss_host = torch.empty((2, B), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        hb = host[i % n_batches]
        db = stage[i % 2]
        for dst, src in zip(db, hb):
            dst.copy_(src, non_blocking=True)
        out = train(db[0], db[1], db[2], db[3], kl_weight=1.0, global_batch=gbatch, lr=3e-4, clip_gradient=1.0, allreduce=ar)
        loss_host[0].copy_(out["ce"], non_blocking=True)
        loss_host[1].copy_(out["kl"], non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the step's result is read on the host
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks else None
        for i in range(warmup):
            fn(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.LAUNCHES
        if sampler:
            sampler.mark_start()
        s.record()
        for i in range(steps):
            fn(warmup + i)
        e.record()
        barrier()
        if sampler:
            sampler.mark_end()
        launches = lib.LAUNCHES - l0
        clocks = sampler.stop() if sampler else None
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    W, K = max(args.warmup, 3), args.steps
    ms, launches, clocks = timed(step_resident, K, W, sample_clocks=True)
    value = gbatch * K / (ms * 1e-3)
    ms_e2e, _, _ = timed(step_e2e, K, 3)
    e2e_value = gbatch * K / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = loss_host.numel() * loss_host.element_size()

    # ---- dominant kernel (GEMM) roofline: CUDA events around every GEMM launch of a few extra steps.
    # Every rank runs these steps (they contain the gradient all-reduce); only rank 0 records events.
    roofline = roofline_other = None
    psteps = 3
    prof = {"match": "msx_gemm", "events": []}
    if rank == 0:
        lib._profile = prof
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(psteps):
        step_eager(i)
    e.record()
    torch.cuda.synchronize()
    lib._profile = None
    if rank == 0:
        gemm_ms = sum(a.elapsed_time(b) for a, b, _ in prof["events"])
        gemm_flops = sum(f[0] for _, _, f in prof["events"])
        gemm_bytes = sum(f[1] for _, _, f in prof["events"])
        step_ms = s.elapsed_time(e)
        n_launch = max(1, len(prof["events"]))
        if args.gemm_table:
            tab = {}
            for a, b, f in prof["events"]:
                t = tab.setdefault(f[2] if len(f) > 2 else "?", [0, 0.0, 0.0, 0.0])
                t[0] += 1; t[1] += a.elapsed_time(b); t[2] += f[0]; t[3] += f[1]
            for k, t in sorted(tab.items(), key=lambda kv: -kv[1][1]):
                print("gemm %-58s n/step=%4.1f %8.1f us  %7.1f TFLOP/s %7.1f GB/s" % (
                    k, t[0] / psteps, 1e3 * t[1] / t[0], t[2] / (t[1] * 1e-3) / 1e12, t[3] / (t[1] * 1e-3) / 1e9), file=sys.stderr)
        # Two GEMM kernels share a step in the mixed modes: the single-pass one (HBM-bound: fp32-in / fp32-out GEMMs with
        # K, N <= 1024 carry 64-102 flop per algorithmic byte, below the TF32 ridge point ~700 TFLOP/s / 6.55 TB/s = 107) and
        # the 3xTF32 one (tensor-bound: it executes 3 MMAs per algorithmic multiply-add).  `roofline` describes the group
        # with the larger share of the step, `roofline_other` the other one.
        KN = {"tf32": "gemm_tc2_kernel / gemm_tc_kernel (msx_gemm_tc: tcgen05 kind::tf32, cta_group::2 pair tiles, TMA)",
              "bf16": "gemm_tc2_kernel / gemm_tc_kernel (msx_gemm_tc_bf16: tcgen05 kind::f16, bf16 operands, cta_group::2 pair tiles, TMA)",
              "tf32x3": "gemm_tc2x3_kernel (msx_gemm_tc_x3: tcgen05 kind::tf32 with in-kernel hi/lo operand splitting, three "
                        "MMAs per k-block, cta_group::2 pair tiles, TMA)",
              "bf16x3": "gemm_tc2b3_kernel (msx_gemm_tc_b3: fp32 operands split into bf16 hi + lo inside the kernel, three tcgen05 "
                        "kind::f16 MMAs per k-step, cta_group::2 pair tiles, TMA)",
              "bf16p3": "gemm_tc2_kernel / gemm_tc_kernel (msx_gemm_tc_p3: operands as bf16 hi + lo planes from the producing kernels, "
                        "three walks hi*hi + hi*lo + lo*hi of the reduction on tcgen05 kind::f16, cta_group::2 pair tiles, TMA)",
              "ffma": "sgemm_kernel (msx_gemm_f32, fp32 FFMA path)", "f32": "sgemm_kernel (msx_gemm_f32, fp32 FFMA path)"}
        groups = {}
        for a_, b_, f_ in prof["events"]:
            kind = (f_[2].split(" ")[0] if len(f_) > 2 else "?")
            gsum = groups.setdefault(kind, [0, 0.0, 0.0, 0.0])
            gsum[0] += 1; gsum[1] += a_.elapsed_time(b_); gsum[2] += f_[0]; gsum[3] += f_[1]
        tf32_peak = peaks["tflops"] / 2.0
        rl = []
        for kind, (cnt, tms, fl, by) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
            tfk = fl / (tms * 1e-3) / 1e12
            gbk = by / (tms * 1e-3) / 1e9
            common = {"kernel": KN.get(kind, kind), "share_of_step": tms / step_ms, "launches_per_step": cnt / psteps,
                      "avg_launch_ms": tms / cnt, "algorithmic_bytes_per_launch": by / cnt, "algorithmic_flops_per_launch": fl / cnt,
                      "measured": "CUDA events around every launch of %d eager steps (the graph-replayed step is what `value` times)" % psteps}
            if kind in ("bf16x3", "bf16p3"):
                traffic3 = None
                tpath3 = os.path.join(REPO, "profiles", "traffic_gemm_p3.json")
                if kind == "bf16p3" and os.path.exists(tpath3):
                    with open(tpath3) as f:
                        traffic3 = json.load(f).get("dram_bytes_per_launch")
                # three kind::f16 MMAs per multiply-add = 1.5 TF32-equivalents: on these shapes the kernel is back under the
                # HBM roof of its fp32 operand / result bytes
                rl.append(dict(common, bound="hbm", achieved=gbk, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbk / peaks["hbm_gbs"],
                               traffic=traffic3, traffic_source="one ncu --set full capture of the three large forward shapes "
                               "(profiles/traffic_gemm_p3.json), mean per launch; the bench's mean also covers the five small "
                               "top-layer launches" if traffic3 else None,
                               peak_source=peaks["src"] + " HBM copy bandwidth",
                               tensor={"achieved": tfk, "executed_tflops": 3 * tfk, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                       "executed_frac": 3 * tfk / peaks["tflops"],
                                       "peak_source": peaks["src"] + " bf16 sustained (cuBLAS); the kernel executes 3 bf16 MMAs "
                                       "per algorithmic multiply-add"}))
            elif kind == "tf32x3":
                rl.append(dict(common, bound="tensor", achieved=tfk, peak=tf32_peak, unit="TFLOP/s", frac=tfk / tf32_peak,
                               executed_tflops=3 * tfk, executed_frac=3 * tfk / tf32_peak, traffic=None,
                               peak_source=peaks["src"] + " bf16 sustained (cuBLAS) / 2 = TF32 dense; `achieved` counts the "
                               "algorithmic 2MNK, the kernel executes 3 MMAs per multiply-add (executed_*)",
                               hbm={"achieved": gbk, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbk / peaks["hbm_gbs"]}))
            else:
                traffic = None
                tpath = os.path.join(REPO, "profiles", "traffic_gemm_bf16.json" if kind == "bf16" else "traffic_gemm.json")
                if kind in ("tf32", "bf16") and os.path.exists(tpath):
                    with open(tpath) as f:
                        traffic = json.load(f).get("dram_bytes_per_launch")
                rl.append(dict(common, bound="hbm", achieved=gbk, peak=peaks["hbm_gbs"], unit="GB/s", frac=gbk / peaks["hbm_gbs"],
                               traffic=traffic, traffic_source="one ncu --set full capture of a round-1 step's GEMM launches "
                               "(profiles/traffic_gemm*.json), per launch" if traffic else None,
                               peak_source=peaks["src"] + " HBM copy bandwidth",
                               tensor={"achieved": tfk, "peak": tf32_peak if kind != "bf16" else peaks["tflops"], "unit": "TFLOP/s",
                                       "frac": tfk / (tf32_peak if kind != "bf16" else peaks["tflops"])}))
        roofline = rl[0] if rl else None
        roofline_other = rl[1:] or None
        if roofline:
            roofline["gemm_share_of_step"] = gemm_ms / step_ms
            roofline["gemm_flops_per_step"] = gemm_flops / psteps
    if args.kernel_table and rank == 0 and world == 1:
        kp = {"match": "msx_", "events": [], "names": True}
        lib._profile = kp
        for i in range(psteps):
            step_eager(i)
        torch.cuda.synchronize()
        lib._profile = None
        tab = {}
        for a, b, f in kp["events"]:
            name = f if isinstance(f, str) else (f[2].split(" ")[0] + " gemm" if isinstance(f, tuple) and len(f) > 2 else "?")
            t = tab.setdefault(name, [0, 0.0])
            t[0] += 1; t[1] += a.elapsed_time(b)
        tot = sum(t[1] for t in tab.values())
        for k, t in sorted(tab.items(), key=lambda kv: -kv[1][1]):
            print("kernel %-28s n/step=%5.1f %8.1f us/call %8.3f ms/step %5.1f%%" % (
                k, t[0] / psteps, 1e3 * t[1] / t[0], t[1] / psteps, 100 * t[1] / tot), file=sys.stderr)
    if world > 1:
        barrier()

    # ---- data parallel: every rank must hold bit-identical parameters after the timed steps (the exchange kernel
    # writes each rank's updated slice into every peer's arena; NCCL path: identical all-reduced gradients)
    ranks_identical = None
    if world > 1:
        w = eng.arena.w
        chk = torch.stack([w.double().sum(), w.view(torch.int32).to(torch.int64).sum().double()])
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        ranks_identical = bool(all(torch.equal(c, allchk[0]) for c in allchk))

    def time_variant(precision, B_v, gbatch_v, steps_v, warm_v, seed_off=0, L_v=None):
        """Same model / data generator / timing protocol on a fresh engine: `precision`, B_v rows per GPU per step,
        gradients averaged over gbatch_v rows, rows of L_v tokens (default: the headline's).  Data parallel runs take the
        same exchange as the headline."""
        L_v = L if L_v is None else L_v
        eng_v = VAEEngine(cfg, dev, seed=0, precision=precision)
        ar_v = None
        exch = "none"
        if world > 1:
            ar_v, exch = allreduce, "nccl all-reduce + Adam on every rank"
            if args.dp == "peer" and ar == "peer":
                try:
                    eng_v.enable_peer_optimizer()
                    ar_v, exch = "peer", "fused NVLink optimiser step (msx_adam_nvlink_step)"
                except Exception as exc:
                    print("bench: peer optimiser unavailable for the %s variant (%s)" % (precision, exc), file=sys.stderr)
        tok_v, lens_v, cls_v, lab_v = synth.token_rows_4_4(B_v * n_batches, L_v, seed=200 + seed_off + rank)
        res_v = [tuple(torch.from_numpy(a[i * B_v:(i + 1) * B_v].copy()).to(dev) for a in (tok_v, lens_v, cls_v, lab_v))
                 for i in range(n_batches)]
        train_v = eng_v.train_step if args.no_graph else eng_v.train_step_graphed

        def step_v(i):
            tk, ln, cl, lb = res_v[i % n_batches]
            return train_v(tk, ln, cl, lb, kl_weight=1.0, global_batch=gbatch_v, lr=3e-4, clip_gradient=1.0, allreduce=ar_v)

        ms_v, _, _ = timed(step_v, steps_v, warm_v)
        out = {"value": gbatch_v * steps_v / (ms_v * 1e-3), "unit": UNIT, "ms_per_step": ms_v / steps_v,
               "batch_per_gpu": B_v, "global_batch": gbatch_v, "steps": steps_v, "dp_exchange": exch}
        eng_v._graphs.clear()
        del eng_v, train_v, res_v
        torch.cuda.synchronize()
        if world > 1:
            barrier()
        return out

    # ---- bf16 variant (BASELINE config 4), stated separately: same model, batch, data and timing protocol with
    # precision="bf16"; its tolerances are the bf16 ones of tests/test_engine_gpu.py, not the fp32 bar of the headline.
    bf16_variant = fp32_variant = b32 = x3_variant = seq_sweep = None
    strong = []
    tf32_variant = None
    if args.precision == HEADLINE and not args.no_variants:
        # ---- round 2's first headline: the same forward accuracy class with the operand split INSIDE the GEMM (3xTF32)
        x3_variant = time_variant("tf32x3f", B, gbatch, K, W)
        x3_variant.update({"dtype": "f32 forward (3xTF32) / tf32 backward", "what": PRECISION_NOTE["tf32x3f"]})
        tf32_variant = time_variant("tf32", B, gbatch, K, W)
        tf32_variant.update({"dtype": "tf32", "what": PRECISION_NOTE["tf32"]})
        bf16_variant = time_variant("bf16", B, gbatch, K, W)
        bf16_variant.update({"dtype": "bf16", "what": PRECISION_NOTE["bf16"],
                             "parity": "vs fp32 oracle: loss 1e-4, KL 1e-3, latent means 1e-2, gradients 10 % worst tensor / "
                                       "3 % mean (tests/test_engine_gpu.py::test_bf16_variant_step_vs_oracle)"})
        # ---- strict-fp32 variant: the reference's own precision (trainer.py:155-179 is fp32 end to end)
        fp32_variant = time_variant("fp32x3", B, gbatch, max(5, K // 2), 3)
        fp32_variant.update({"dtype": "f32", "what": PRECISION_NOTE["fp32x3"],
                             "parity": "loss / KL / latent means 1e-5, every gradient within 1e-3 of its scale vs the fp32 oracle "
                                       "(tests/test_parity_bench_gpu.py::test_bench_shape_gradients_vs_oracle[fp32x3])"})
        # ---- strong scaling: the global batch is fixed, each GPU takes 1/N of it
        for gb in (2048, 256):
            if gb % world == 0 and not (world == 1 and gb == B):
                sv = time_variant(args.precision, gb // world, gb, K, W, seed_off=7)
                sv["scaling"] = "strong"
                strong.append(sv)
        # ---- the reference's own configuration of record (scripts/train-vae.sh: batch 32), one graph launch per step
        if world == 1:
            b32 = time_variant(args.precision, 32, 32, 200, 10, seed_off=11)
            b32["what"] = "scripts/train-vae.sh batch size (32 rows, L=64), CUDA-graph replay"
        # ---- BASELINE config 3's sequence-length axis at the headline batch: the long-row tcgen05 attention
        # (csrc/attention_tc_long.cu) takes T = L + 1 = 129 / 257; tokens/s next to the headline's
        if world == 1 and L == 64:
            seq_sweep = []
            for L_s in (128, 256):
                sv = time_variant(args.precision, B, gbatch, 10, 3, seed_off=13, L_v=L_s)
                sv.update({"seq_len": L_s, "tokens_per_s": sv["value"] * L_s})
                seq_sweep.append(sv)
    e2e_from_midi = None
    if world == 1 and not args.no_variants and args.dec_type == "lstm":
        e2e_from_midi = bench_from_midi(args, cfg, dev)

    if rank != 0:
        if world > 1:
            eng._graphs.clear()
            torch.cuda.synchronize()
            dist.destroy_process_group()
        return

    raster = None
    if not args.no_raster:
        raster = bench_rasteriser(peaks)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, per_step, cores = oracle_step_rate(args, steps=3, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "3 oracle steps of %d rows, the GPU arm's own step (torch-CPU fp32 restatement of Trainer._step, "
                         "%d host threads)" % (args.cpu_batch or args.batch, cores),
               "ms_per_step": per_step * 1e3}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "fp32x3": "f32", "tf32x3f": "f32 forward (3xTF32) / tf32 backward", "bf16x3f": "f32 storage, bf16x3 / 3xTF32 forward, tf32 backward", "bf16p3f": "f32 storage, forward products from bf16 hi+lo planes (~2^-17) / tf32 backward", "tf32": "tf32", "bf16": "bf16"}[args.precision],
        "data": "synthetic",
        "config": config_dict(args, world),
        "run": {"precision": PRECISION_NOTE[args.precision], "cuda_graph": not args.no_graph, "dp_exchange": dp_exchange, "ranks_identical": ranks_identical,
                "l2": "per-step working set (activations ~%.1f GB) exceeds the 126 MB L2; 4 input batches rotate" %
                      (B * T * 4 * 40e3 / 1e9 / 10)},
        "roofline": roofline, "roofline_other": roofline_other, "rasteriser": raster, "cpu_baseline": cpu, "tf32x3f_variant": x3_variant, "tf32_variant": tf32_variant, "bf16_variant": bf16_variant,
        "fp32_variant": fp32_variant, "strong_scaling": strong or None, "b32": b32, "seq_len_sweep": seq_sweep,
        "e2e_from_midi": e2e_from_midi,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": launches, "clocks": clocks,
    }
    emit(line)
    if world > 1:
        eng._graphs.clear()
        torch.cuda.synchronize()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------- BASELINE config 3 / 5
def run_sweep(args):
    """Config 3: fp32-storage training throughput over batch and sequence length on one GPU (device-resident batch)."""
    import torch
    from musicstyletransfer_b200 import lib, synth
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    cfg = VAEConfig(dec_type=args.dec_type, enc_dropout=args.dropout, dec_dropout=args.dropout)
    for L in (64, 128, 256, 512):
        for B in (32, 128, 512, 2048, 8192):
            if L == 512 and B > 2048:                       # 4 M positions: the saved activations alone would be ~130 GB
                continue
            eng = VAEEngine(cfg, dev, seed=0, precision=args.precision)
            tok, lens, cls, lab = synth.token_rows_4_4(B * 2, L, seed=7)
            bat = [tuple(torch.from_numpy(a[i * B:(i + 1) * B].copy()).to(dev) for a in (tok, lens, cls, lab)) for i in range(2)]
            steps = max(5, min(200, int(60000 / max(B, 64))))
            train = eng.train_step if args.no_graph else eng.train_step_graphed
            for i in range(3):
                train(*bat[i % 2], kl_weight=1.0, global_batch=B, lr=3e-4, clip_gradient=1.0)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(steps):
                train(*bat[i % 2], kl_weight=1.0, global_batch=B, lr=3e-4, clip_gradient=1.0)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / steps
            emit({"metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B,
                              "seq_len": L, "T": L + 1, "steps": steps, "precision": args.precision, "cuda_graph": not args.no_graph,
                              "config": {"workload": "config 3 sweep point, device-resident batch, " + args.dec_type}})
            del eng, bat
            torch.cuda.empty_cache()


def run_style(args):
    """Config 5: style-transfer inference (encode the source with the target class, z = means, decode 2T steps with
    multinomial sampling) for every class over a large synthetic batch; the oracle times a small sample on the CPU."""
    import numpy as np
    import torch
    from musicstyletransfer_b200 import lib, synth
    from musicstyletransfer_b200.engine import VAEConfig, VAEEngine
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    cfg = VAEConfig(dec_type=args.dec_type)
    eng = VAEEngine(cfg, dev, seed=0, precision=args.precision)
    B, L = 8192, args.seq_len
    tok, lens, cls, lab = synth.token_rows_4_4(B, L, seed=11)
    tk, ln = torch.from_numpy(tok).to(dev), torch.from_numpy(lens).to(dev)
    targets = [torch.full((B,), c, dtype=torch.int32, device=dev) for c in range(cfg.num_classes)]

    def one_pass():
        n = 0
        for c in range(cfg.num_classes):                      # sampler.py:95 / :126: the class vector is overwritten per class
            seqs, _ = eng.style_transfer(tk, ln, targets[c], seed=3)
            n += seqs.shape[1]
        return n
    one_pass()
    torch.cuda.synchronize()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        ntok = one_pass()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / reps
    line = {"metric": "style_transfer_sequences_per_sec", "value": B * cfg.num_classes / sec, "unit": "sequences/s",
            "tokens_per_s": B * ntok / sec, "ms_per_pass": sec * 1e3, "n_gpus": 1, "higher_is_better": True,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "style transfer: %d rows x %d classes, L=%d, decoder %s, up to 2T=%d sampled steps per class; "
                                   "wall clock including the host-side stop test" % (B, cfg.num_classes, L, args.dec_type, 2 * (L + 1))}}
    if not args.no_cpu_baseline:
        from oracle import model as om
        cfg_o = om.Cfg(dec_type=args.dec_type)
        p = om.init_params(cfg_o, seed=0)
        Bc = 16
        rng = np.random.default_rng(0)
        u = torch.from_numpy(rng.random((2 * (L + 1), Bc), dtype=np.float32))
        fn = om.style_transfer_lstm if args.dec_type == "lstm" else om.style_transfer_transformer
        t0 = time.perf_counter()
        for c in range(cfg_o.num_classes):
            fn(cfg_o, p, torch.from_numpy(tok[:Bc]).float(), torch.full((Bc,), float(c)), u)
        csec = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Bc * cfg_o.num_classes / csec, "unit": "sequences/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": "%d rows x %d classes through the oracle" % (Bc, cfg_o.num_classes)}
    emit(line)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sweep":
        run_sweep(args)
    elif args.mode == "style":
        run_style(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
