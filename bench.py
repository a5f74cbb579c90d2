#!/usr/bin/env python
"""Throughput bench of the PASSION training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one forward + PASSION loss + backward + AdamW(amsgrad) update on one synthetic batch of
BraTS-shaped 4x80^3 crops, B = 2 per GPU (BASELINE.json configs[1]); at N > 1 (torchrun, one rank per GPU,
NCCL) each rank takes its own B = 2 shard of the global batch (weak scaling, configs[2]).
Prints ONE JSON line (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same metric
through the public API with pinned HOST buffers (H2D of x / uint8 label map / mask every step, the H2D of batch i+1
overlapping step i, and a D2H read of every step's loss, one step behind, inside the timed region).  `roofline` is the live
CUDA-event measurement of the dominant kernel family.
`--impl reference` times the UNMODIFIED reference model + criterions staged under baseline/_ref (PyTorch fp32 on the host
cores, same config block); where that tree is missing it falls back to the oracle port and says so (`kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (4x80^3 crops, fwd+bwd+PASSION loss+AdamW)"
S_CROP = 80
B_PER_GPU = 2
# algorithmic work per sample (dense as the reference executes it)
FLOP_PER_SAMPLE = 557.8e9
BYTES_PER_SAMPLE = 14.7e9


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r01_*_metrics.txt, profiles/r02_*_metrics.txt), keyed by kernel family; None = not captured yet.
NCU_TRAFFIC_BYTES = {
    "conv3d_fwd_tc": {"class": "c16->8 k3 s1 80x80x80 n8 g1", "bytes": 131.29e6 + 40.21e6, "source": "profiles/r01_conv_tc_kdstack_metrics.txt"},
    "conv3d_dgrad_tc": {"class": "c16->8 k3 s1 80x80x80 n8 g1", "bytes": 131.29e6 + 40.21e6, "source": "profiles/r01_conv_tc_kdstack_metrics.txt"},
    "conv3d_wgrad_tc": {"class": "c16->8 k3 s1 80x80x80 n10 g1", "bytes": 245.85e6 + 3.83e6, "source": "profiles/r02_wgrad_rs_metrics.txt"},
    "inorm_lrelu_bwd": {"class": "c8 (n10, 80^3): reduce pass 163.86e6 + 4.24e6, apply pass 163.86e6 + 46.95e6 (write-back partly deferred)",
                        "bytes": 163.86e6 + 4.24e6 + 163.86e6 + 46.95e6, "source": "profiles/r01_final_tuned_kernels_metrics.txt"},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        import statistics
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def mask_ids_for(n_samples, offset=0):
    """Missing-modality masks drawn from the reference's imbalanced mr2468 table."""
    import csv
    import numpy as np
    with open(os.path.join(ROOT, "tests", "golden", "Brats2020_imb_split_mr2468.csv")) as f:
        table = [int(r["mask_id"]) for r in csv.DictReader(f)]
    rs = np.random.RandomState(1037)
    ids = [table[i] for i in rs.randint(0, len(table), offset + n_samples)]
    return ids[offset:]


MASK_ARRAY = [[False, False, False, True], [False, True, False, False], [False, False, True, False], [True, False, False, False],
              [False, True, False, True], [False, True, True, False], [True, False, True, False], [False, False, True, True],
              [True, False, False, True], [True, True, False, False], [True, True, True, False], [True, False, True, True],
              [True, True, False, True], [False, True, True, True], [True, True, True, True]]     # datasets_nii.py:27-30


def synth_host_batches(rank, n_batches, B, S):
    """Pinned-host synthetic batches in the reference loader's format: x f32 [B,4,S,S,S], one-hot target f64, mask bool."""
    import numpy as np
    import torch
    rs = np.random.RandomState(1037 + rank)
    out = []
    for i in range(n_batches):
        x = torch.from_numpy(rs.standard_normal((B, 4, S, S, S)).astype(np.float32))
        y = rs.randint(0, 4, (B, S, S, S))
        target = torch.from_numpy(np.ascontiguousarray(np.eye(4)[y].transpose(0, 4, 1, 2, 3)))
        ids = mask_ids_for(B, offset=(rank * n_batches + i) * B)
        mask = torch.tensor([MASK_ARRAY[j] for j in ids])
        out.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in (x, target, mask)))
    return out


def workload_config(model, B, S, world):
    """The `config` block, identical in both arms (the driver compares them)."""
    label = {"rfnet": "RFNet", "mmformer": "mmFormer"}[model]
    return {"workload": f"{label}+PASSION train step, B={B}/GPU, 4x{S}^3 crops, idt masks from mr2468, temp 4, AdamW amsgrad",
            "global_batch": B * world, "parallelism": f"dp{world}"}


def modal_weight():
    """iter_per_epoch / modal_num of the mr2468 table (train.py:163-171): 219 / (90, 135, 184, 43)."""
    import torch
    return torch.tensor([219 / 90.0, 219 / 135.0, 219 / 184.0, 219 / 43.0])


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from passion_b200 import _lib, ops
    from passion_b200.engine import Trainer
    from passion_b200.models import build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")            # keeps NCCL's version banner off stdout (one JSON line only)
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "nccl_bench_%h_%p.log"))
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    torch.manual_seed(1037)
    model = build_model(args.model, num_cls=4, crop=args.size).to(dev)
    model.compute_dtype = dtype
    use_graph = not args.no_graph
    trainer = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True,
                      modal_weight=modal_weight(), use_graph=use_graph)
    B, S = args.batch, args.size
    nb = 2
    host = synth_host_batches(rank, nb, B, S)
    devb = [tuple(t.to(dev) for t in b) for b in host]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.step(*devb[i % nb])
    sync()

    # ---- timed region 1: inputs resident in HBM (whole step, incl. the NCCL all-reduces, replayed as one CUDA graph)
    clocks = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        loss, _ = trainer.step(*devb[i % nb])
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    # ---- the same steps once more, eagerly, with every kernel launch bracketed by CUDA events on the launching
    #      stream: per-kernel durations for the roofline block and the launch count (events cannot sit inside a graph)
    trainer._eager_step(*devb[0])                       # untimed: lets the caching allocator settle for eager execution
    torch.cuda.synchronize()
    timer = ops.KernelTimer()
    ops.TIMER = timer
    l0 = _lib.launch_count()
    n_prof = min(args.steps, 3)
    for i in range(n_prof):
        # park the GPU behind a ~150 ms spin so the host enqueues the whole step ahead of it: the event pairs then
        # bracket back-to-back kernel executions instead of host launch latency
        torch.cuda._sleep(int(3e8))
        trainer._eager_step(*devb[i % nb])
        torch.cuda.synchronize()
    sync()
    ops.TIMER = None
    launches = (_lib.launch_count() - l0) // n_prof * args.steps
    ops.check_tc_errors()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)

    # ---- timed region 2: end to end through the public API from pinned host buffers
    #      The target travels as the uint8 label map [B,S,S,S] (the compact form Model.forward / criterions accept next to the
    #      reference loader's float64 one-hot, which is np.eye(4)[label] of the same map: 1 MB instead of 33 MB per batch).
    from passion_b200.engine import DevicePrefetcher
    host = [(x, t.argmax(1).to(torch.uint8).contiguous().pin_memory(), m) for x, t, m in host]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    for batch in DevicePrefetcher((host[i % nb] for i in range(2)), dev, like=host[0]):      # untimed: one-time setup of the path
        trainer.step(*batch)                                                                 # (re-captures the step for this target format)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    feed = DevicePrefetcher((host[i % nb] for i in range(args.steps)), dev, like=host[0])   # H2D of batch i+1 overlaps step i
    sync()
    e2.record()
    last = 0.0
    # D2H read of EVERY step's result, one step behind: the loss of step i is copied into pinned host memory right behind the
    # step on the compute stream and read by the host after step i+1 has been enqueued (what a training loop's logging does),
    # so the host's launch latency does not sit between two steps.  The last loss is read inside the timed region.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    k = 0
    for batch in feed:
        loss, _ = trainer.step(*batch)
        feed.release(batch)
        loss_host[k % 2].copy_(loss.detach().reshape(()).float(), non_blocking=True)
        loss_ready[k % 2].record()
        if k > 0:
            loss_ready[(k - 1) % 2].synchronize()
            last = float(loss_host[(k - 1) % 2])
        k += 1
    if k > 0:
        loss_ready[(k - 1) % 2].synchronize()
        last = float(loss_host[(k - 1) % 2])
    e3.record()
    sync()
    clk = clocks.stop()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2_total = float(ms2)

    # ---- timed region 3 (additional): the device-side sample pipeline.  Synthetic cases resident in HBM;
    #      per step the host draws the reference's transform parameters and sends ~10 KB, one kernel builds x and the uint8
    #      label map (passion_b200/data.py), the same training step follows, the loss is read back.
    resident = None
    try:
        if world > 1:                                    # single-GPU figure: at N > 1 the step graph holds NCCL work and is not re-captured here
            raise StopIteration
        import random
        import numpy as np
        from passion_b200.data import AugmentSampler, DeviceAugment, ResidentCases
        rs = np.random.RandomState(77 + rank)
        cases = ResidentCases(dev)
        shp = (S + 40, S + 40, S + 24)
        for _ in range(2):
            cases.add(rs.standard_normal(shp + (4,)).astype(np.float32), rs.randint(0, 4, shp).astype(np.uint8))
        smp = AugmentSampler((S, S, S), py_rng=random.Random(1037 + rank), np_rng=np.random.RandomState(1037 + rank))
        aug = DeviceAugment(dev, size=(S, S, S), batch=B)
        masks = [b[2] for b in devb]

        def resident_step(i):
            ids = [(i + b) % len(cases) for b in range(B)]
            x, labels, _ = aug(cases, ids, [smp.sample(shp) for _ in ids])
            return trainer.step(x, labels, masks[i % nb])

        for i in range(2):                               # untimed: re-capture of the step for the label-map input format
            resident_step(i)
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e4.record()
        for i in range(args.steps):
            loss, _ = resident_step(i)
            last_r = float(loss.item())
        e5.record()
        sync()
        ms3 = torch.tensor([e4.elapsed_time(e5)], device=dev)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        resident = {"value": round(args.steps * B * world / (float(ms3) / 1e3), 3), "unit": "samples/s",
                    "ms_per_step": round(float(ms3) / args.steps, 3), "h2d_bytes_per_step": aug.param_bytes(),
                    "d2h_bytes_per_step": 4, "last_loss": last_r,
                    "what": "cases resident in HBM; crop+rotation+intensity+flip+label encoding as one kernel per batch "
                            "(bit-exact vs the reference's numpy/scipy transforms), uint8 label-map target"}
    except StopIteration:
        resident = None
    except Exception as exc:                             # an additional figure must never take the headline line down
        resident = {"error": repr(exc)[:300]}

    def finish():
        """Leave without tearing NCCL down: destroying a process group while captured graphs still hold its
        communicator can block; every rank has passed the final barrier, so exiting directly is safe."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            os._exit(0)

    if rank != 0:
        finish()
        return
    samples = args.steps * B * world
    value = samples / (ms_total / 1e3)
    e2e = samples / (ms2_total / 1e3)
    hbm_peak, tc_peak, how = peaks()
    summ = timer.summary()
    fam = {}
    for (name, key), d in summ.items():
        f = fam.setdefault(name, dict(calls=0, ms=0.0, bytes=0, flops=0))
        for k in f:
            f[k] += d[k]
    for f in fam.values():                      # normalise the instrumented pass to the timed region's step count
        for k in ("calls", "ms", "bytes", "flops"):
            f[k] = f[k] * args.steps / n_prof
    for d in summ.values():
        for k in ("calls", "ms", "bytes", "flops"):
            d[k] = d[k] * args.steps / n_prof
    top_name = max(fam, key=lambda k: fam[k]["ms"])
    top = fam[top_name]
    top_cls = max(((k, d) for k, d in summ.items() if k[0] == top_name), key=lambda kv: kv[1]["ms"])
    ach = top["bytes"] / (top["ms"] / 1e3) / 1e9
    ach_tf = top["flops"] / (top["ms"] / 1e3) / 1e12
    # the tcgen05 conv kernels are limited by the tensor pipe (ncu: pipe 65-90 % busy, dram bytes = algorithmic bytes), every
    # other family streams: report the family against the roof that actually bounds it, and the other figure beside it
    tensor_bound = top_name.endswith("_tc")
    roofline = {"bound": "tensor" if tensor_bound else "hbm", "kernel": top_name,
                "achieved": round(ach_tf, 2) if tensor_bound else round(ach, 1),
                "peak": tc_peak if tensor_bound else hbm_peak, "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": round(ach_tf / tc_peak, 4) if tensor_bound else round(ach / hbm_peak, 4),
                "hbm_GBps": round(ach, 1), "hbm_frac": round(ach / hbm_peak, 4),
                "traffic": (NCU_TRAFFIC_BYTES.get(top_name) or {}).get("bytes"),
                "traffic_detail": NCU_TRAFFIC_BYTES.get(top_name), "peak_source": how,
                "launches": int(top["calls"]), "avg_launch_ms": round(top["ms"] / top["calls"], 4),
                "share_of_step": round(top["ms"] / ms_total, 3),
                "top_class": {"key": top_cls[0][1], "ms_per_launch": round(top_cls[1]["ms"] / top_cls[1]["calls"], 4),
                              "GBps": round(top_cls[1]["bytes"] / (top_cls[1]["ms"] / 1e3) / 1e9, 1),
                              "TFLOPs": round(top_cls[1]["flops"] / (top_cls[1]["ms"] / 1e3) / 1e12, 2)},
                "step_hbm_frac": round(value / world * BYTES_PER_SAMPLE / 1e9 / hbm_peak, 4),
                "step_tc_frac": round(value / world * FLOP_PER_SAMPLE / 1e12 / tc_peak, 4),
                "families_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}
    metric = METRIC if S == S_CROP else METRIC.replace("4x80^3", f"4x{S}^3")
    out = {"metric": metric, "value": round(value, 3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if dtype == torch.bfloat16 else "f32", "data": "synthetic",
           "config": workload_config(args.model, B, S, world),
           "notes": {"step_execution": "one CUDA graph replay per step" if use_graph else "eager launches",
                     "roofline_region": "eager re-run of the same steps with CUDA events around every kernel launch",
                     "l2": "per-step working set (~6 GiB of activations) exceeds the 126 MB L2; no explicit flush"},
           "clocks": clk,
           "e2e": {"value": round(e2e, 3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                   "ms_per_step": round(ms2_total / args.steps, 3), "last_loss": last,
                   "target_format": "uint8 label map (x f32 + labels u8 + mask from pinned host memory every step)",
                   "loss_readback": "every step's loss is copied to pinned host memory and read one step behind"},
           "e2e_resident_cases": resident,
           "gpu_launches": int(launches), "roofline": roofline}
    dump = os.environ.get("PB_DUMP_KERNELS")
    if dump:
        rows = sorted(((k[0], k[1], d["calls"] // args.steps, d["ms"] / args.steps, d["bytes"] / (d["ms"] / 1e3) / 1e9,
                       d["flops"] / (d["ms"] / 1e3) / 1e12) for k, d in summ.items()), key=lambda r: -r[3])
        with open(dump, "w") as f:
            f.write("kernel | class | launches/step | ms/step | GB/s (algorithmic) | TFLOP/s\n")
            for r in rows:
                f.write(f"{r[0]} | {r[1]} | {int(r[2])} | {r[3]:.3f} | {r[4]:.1f} | {r[5]:.2f}\n")
    if args.model != "rfnet":
        # SURVEY.md §8d's per-sample flop / byte figures are RFNet's; the secondary backbone reports kernel families only
        roofline.pop("step_hbm_frac"); roofline.pop("step_tc_frac")
    if world == 1 and args.model == "rfnet" and not args.no_extras:
        del trainer, model
        torch.cuda.empty_cache()
        out["extra_configs"] = extra_configs(dev)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], ref_first = cpu_baseline(budget_s=30.0, model=args.model, S=S)
        try:
            out["parity"] = parity_block(ref_first, args.model, S)
        except Exception as exc:                         # an additional block must never take the headline line down
            out["parity"] = {"error": repr(exc)[:300]}
    emit(out)
    finish()


# ----------------------------------------------------------------------------------------- CPU reference arm
def _oracle_step_fn(S, B=1, model="rfnet"):
    """One training step of the reference algorithm's CPU port (oracle/, torch fp32 on the host cores)."""
    import torch
    from oracle import mmformer_oracle, rfnet_oracle, synth, train_step_oracle
    torch.set_num_threads(os.cpu_count())
    fwd = rfnet_oracle.forward if model == "rfnet" else mmformer_oracle.forward
    sd = synth.make_state_dict(1037, None if model == "rfnet" else synth.mmformer_param_shapes(patch=S // 16))
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW([{"params": list(P.values()), "lr": 2e-4, "weight_decay": 1e-4}], betas=(0.9, 0.999),
                            eps=1e-8, amsgrad=True)
    x, target, mask, _ = synth.make_batch(B, S, seed=1037, labels="U", mask_ids=mask_ids_for(B))
    beta, mw = torch.ones(4), modal_weight()

    def step():
        outs = fwd(P, x, mask, target, 4.0)
        loss, _ = train_step_oracle.loss_mix(outs, target, mask, beta, mw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss), outs
    return step, sd


def _reference_step_fn(S, B, model="rfnet"):
    """-> (step, state_dict, kind).  kind = "reference": the UNMODIFIED reference model + criterions staged under
    baseline/_ref (baseline/run_cpu_reference.py); "port": the oracle restatement, when the staged tree is absent or for the
    secondary backbone."""
    import torch
    from oracle import synth
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import run_cpu_reference as rcr
    if model == "rfnet" and rcr.available():
        x, target, mask, _ = synth.make_batch(B, S, seed=1037, labels="U", mask_ids=mask_ids_for(B))
        sd = synth.make_state_dict(1037)
        step, _ = rcr.step_fn(x, target, mask, torch.ones(4), modal_weight(), temp=4.0, state_dict=sd)
        return step, sd, "reference"
    step, sd = _oracle_step_fn(S, B, model)
    return step, sd, "port"


def cpu_baseline(budget_s=30.0, model="rfnet", S=S_CROP):
    """One sample (B = 1) of the bench workload per step on the host cores: 1 warm-up + up to 3 timed steps inside the
    budget, median reported.  Returns (json block, first-step outputs for the parity block)."""
    import torch
    t_start = time.time()
    step, sd, kind = _reference_step_fn(S, 1, model)
    t0 = time.time()
    loss0, outs0 = step()                               # warm-up (library initialisation); its loss is at the initial weights
    first = time.time() - t0
    ts = []
    while len(ts) < 3 and (not ts or time.time() - t_start + first < budget_s):
        t0 = time.time()
        step()
        ts.append(time.time() - t0)
    ts.sort()
    dt = ts[len(ts) // 2]
    what = ("the unmodified reference model + criterions (baseline/_ref), train.py's step restated" if kind == "reference"
            else "oracle/ (PyTorch-CPU restatement of the reference)")
    blk = {"value": round(1.0 / dt, 4), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
           "sample": f"median of {len(ts)} steps after 1 warm-up ({first:.1f} s), B=1, 4x{S}^3, fp32, {what}, {dt:.2f} s/step"}
    return blk, (loss0, [o.detach() for o in outs0], sd)


def parity_block(ref_first, model_name, S):
    """First-step quantities of OUR arm (bf16 production mode and fp32 check mode) against the CPU arm's first step on the
    same weights and the same B = 1 batch, at the bench's crop size."""
    import torch
    from oracle import synth
    from passion_b200.models import build_model
    from passion_b200.train_step import loss_mix
    loss_ref, outs_ref, sd = ref_first
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    x, target, mask, _ = synth.make_batch(1, S, seed=1037, labels="U", mask_ids=mask_ids_for(1))
    names = ["fuse_prob", "prm_loss", "sep_loss", "kl_loss", "proto_loss", "dist"]
    out = {"what": f"first step, B=1, 4x{S}^3, same weights/batch as cpu_baseline; rel-L2 of Model.forward outputs and |dloss|/loss",
           "loss_cpu": round(loss_ref, 6)}
    for tag, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        model = build_model(model_name, num_cls=4, crop=S).to(dev)
        model.load_state_dict(sd)
        model.is_training, model.use_passion, model.mask_type = True, True, "idt"
        model.compute_dtype = dt
        with torch.no_grad():
            outs = model(x.to(dev), mask.to(dev), target=target.to(dev), temp=4.0)
            loss, _ = loss_mix(outs, target.to(dev), mask.to(dev), torch.ones(4, device=dev), modal_weight().to(dev), mask_type="idt")
        rels = {}
        for n, a, b in zip(names, outs, outs_ref):
            a, b = a.double().cpu(), b.double()
            rels[n] = float(f"{float((a - b).norm() / b.norm().clamp_min(1e-30)):.3e}")
        out[tag] = {"loss": round(float(loss), 6), "loss_rel": float(f"{abs(float(loss) - loss_ref) / abs(loss_ref):.3e}"), "outputs_rel_l2": rels}
        del model
    return out


def extra_configs(dev, budget_s=60.0):
    """BASELINE.json configs[3] and configs[4] as additional, driver-visible figures (N = 1; never part of `value`):
    mmFormer + PASSION on one 4x128^3 crop per GPU (the per-GPU share of configs[3]'s 8-GPU job) and the 15-mask sliding-window
    inference sweep over a 240x240x155 volume.  Each is bounded to a few seconds; failures are reported, not raised."""
    import numpy as np
    import torch
    from passion_b200 import ops
    from passion_b200.engine import Trainer
    from passion_b200.models import build_model
    from passion_b200.predict import _windows, predict_all_masks
    out = {}
    t_start = time.time()
    try:
        torch.manual_seed(1037)
        model = build_model("mmformer", num_cls=4, crop=128).to(dev)
        model.compute_dtype = torch.bfloat16
        tr = Trainer(model, lr=2e-4, weight_decay=1e-4, temp=4.0, mask_type="idt", use_passion=True, modal_weight=modal_weight(), use_graph=True)
        batch = tuple(t.to(dev) for t in synth_host_batches(0, 1, 1, 128)[0])
        for _ in range(3):
            tr.step(*batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            tr.step(*batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out["mmformer_128"] = {"workload": "mmFormer+PASSION train step, B=1/GPU, 4x128^3 crop (configs[3], per-GPU share), bf16, CUDA graph",
                               "ms_per_step": round(ms, 2), "samples_per_s": round(1e3 / ms, 2), "steps": n}
        del tr, model, batch
        torch.cuda.empty_cache()
    except Exception as exc:
        out["mmformer_128"] = {"error": repr(exc)[:300]}
    try:
        if time.time() - t_start < budget_s:
            torch.manual_seed(1037)
            model = build_model("rfnet", num_cls=4).to(dev)
            model.compute_dtype = torch.bfloat16
            rs = np.random.RandomState(7)
            shape = (240, 240, 155)
            x = torch.from_numpy(rs.standard_normal((1, 4) + shape).astype(np.float32)).to(dev)
            predict_all_masks(model, x, patch_size=128)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 2
            e0.record()
            for _ in range(reps):
                predict_all_masks(model, x, patch_size=128)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out["inference_15_masks"] = {"workload": "15 missing-modality subsets, sliding window 128^3 (50 % overlap) over 240x240x155 (configs[4]), "
                                                     "RFNet bf16, encoders once per window + one batch-15 decoder pass",
                                         "windows_per_volume": len(_windows(shape, 128)), "ms_per_volume": round(ms, 1),
                                         "volumes_per_s": round(1e3 / ms, 3), "reps": reps}
            ops.check_tc_errors()
    except Exception as exc:
        out["inference_15_masks"] = {"error": repr(exc)[:300]}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the host cores — the unmodified reference
    model + criterions from baseline/_ref (staged by __graft_entry__.build(); /root/reference itself does not exist on the GPU
    box), same config as our arm: B = 2, 4x80^3.  If a full batch per step would push the whole --steps/--warmup run past
    ~5 minutes, each step is a bounded sample (B = 1, half a batch) and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    total = args.steps + args.warmup
    S, B = args.size, args.batch
    step1, _, kind = _reference_step_fn(S, 1, args.model)
    step1()                                              # library warm-up
    t0 = time.time(); step1(); t1 = time.time() - t0     # one B = 1 step: the probe that sizes the run
    del step1
    b_run = B if t1 * B * total <= 330.0 else 1
    step, _, kind = _reference_step_fn(S, b_run, args.model)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    value = args.steps * b_run / dt
    cores = torch.get_num_threads()
    impl = ("unmodified reference model + criterions (baseline/_ref/code: models/rfnet.py, utils/criterions.py), train.py:228-280 "
            "step restated, PyTorch fp32 CPU" if kind == "reference" else "CPU port of the reference algorithm (oracle/, PyTorch fp32)")
    sample = (f"{args.steps} steps of B={b_run} x 4x{S}^3 (" + ("the full per-GPU batch" if b_run == B else f"a bounded sample: {b_run} of the {B} samples of a batch")
              + f"), fp32, {cores} threads; {impl}")
    out = {"impl": "reference", "metric": METRIC if S == S_CROP else METRIC.replace("4x80^3", f"4x{S}^3"), "value": round(value, 4),
           "unit": "samples/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.model, B, S, 1),
           "cpu_baseline": {"value": round(value, 4), "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 4), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def _claim_stdout():
    """Everything that native libraries print on fd 1 (NCCL's version banner, for one) goes to stderr from here on; the ONE
    JSON line is written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--size", type=int, default=S_CROP)
    ap.add_argument("--model", default="rfnet", choices=["rfnet", "mmformer"],
                    help="backbone; the headline metric (BASELINE.json configs[1]) is rfnet")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the additional configs[3] / configs[4] figures (N = 1)")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying one CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_ours(args)


if __name__ == "__main__":
    main()
