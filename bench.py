#!/usr/bin/env python
"""bench.py -- candidate sites/sec (inference), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: 4M synthetic (33,4,4) candidate sites, clairvoyante_v3
forward, fp32, per GPU (the site list shards with no collective: every rank runs the same
4M-site pass on its own GPU => weak scaling).  One "step" = one pass of the hot path over
those sites.
  value : whole-job sites/s with the inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public `Clairvoyante` object with pinned HOST buffers,
          H2D of every input byte and D2H of every result inside the timed region
  roofline : dominant kernel, algorithmic FLOPs / live CUDA-event duration (the path is
          compute-bound: 3,708 FLOP per HBM byte); HBM view alongside
  cpu_baseline : the oracle's torch-CPU fp32 restatement of the TF graph on the host cores
--impl reference times that CPU restatement alone (the reference itself needs TensorFlow
1.12 + Python 2 and cannot run here or on the GPU box).
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

import numpy as np  # noqa: E402

FLOPS_PER_SITE = {  # SURVEY.md 8d / BASELINE.md section 2 (2*MAC)
    "v3": dict(conv1=67584, conv2=950272, conv3=3833856, fc4=3096576, tail=112896 + 6720, total=8067904),
    "v3_slim": dict(conv1=33792, conv2=405504, conv3=2703360, fc4=304128, tail=1296 + 720, total=3448800),
}
HBM_BYTES_PER_SITE = 2176          # 2112 in + 64 out, fp32 I/O
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.45: 148 SMs x 128 FMA lanes x 2 x clocks.max.sm
METRIC = "candidate sites/sec (inference)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        # kernels here are timed inside a long step (back-to-back launches): the sustained cuBLAS figure is the denominator
        sus = d.get("bf16_tflops_sustained")
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=sus or d["bf16_tflops"], bf16_burst=d["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json, %s)" % ("sustained" if sus else "burst"))
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def best_thread_count(pred, x, cores):
    """The CPU arm gets the thread count that serves it best: oneDNN/MKL on a 100+-core host
    with batch-1000 tensors can be far slower with every core than with a subset, so probe a
    few counts (0.6 s each) and keep the fastest.  Reported in `cores`/`sample`."""
    import torch
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    best, best_rate = cands[0], 0.0
    for c in cands:
        torch.set_num_threads(c)
        pred.predict(x)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 0.6:
            pred.predict(x)
            n += len(x)
        rate = n / (time.perf_counter() - t0)
        if rate > best_rate:
            best, best_rate = c, rate
    torch.set_num_threads(best)
    return best


def cpu_reference_rate(variant, W, budget_s, batch=1000):
    """sites/s of the torch-CPU fp32 restatement (oracle) with all host threads, batch =
    param.predictBatchSize (param.py:12); ~budget_s seconds of CPU work."""
    import torch
    from oracle import cv_oracle_torch as OT
    from clairvoyante_b200 import synth
    cores = host_cores()
    pred = OT.CpuPredictor(W, variant)
    x = synth.make_sites(batch, seed=99)
    threads = best_thread_count(pred, x, cores)
    for _ in range(3):
        pred.predict(x)
    n = 0
    t0 = time.perf_counter()
    while True:
        pred.predict(x)
        n += batch
        if time.perf_counter() - t0 >= budget_s and n >= 8 * batch:
            break
    dt = time.perf_counter() - t0
    return n / dt, cores, n, threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (here: its restatement,
    see module docstring) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from clairvoyante_b200 import initializers
    W = initializers.init_weights(args.variant, seed=0)
    import torch
    from oracle import cv_oracle_torch as OT
    from clairvoyante_b200 import synth
    cores = host_cores()
    pred = OT.CpuPredictor(W, args.variant)
    sample = 16 * 1000                                     # sites per step: 16 batches of predictBatchSize
    xs = [synth.make_sites(1000, seed=100 + i) for i in range(16)]
    threads = best_thread_count(pred, xs[0], cores)
    def step():
        for x in xs:
            pred.predict(x)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = dict(metric=METRIC, value=v, unit="sites/s", impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=dict(workload="configs[1]: 4M synthetic (33,4,4) sites, %s forward fp32" % args.variant,
                            sample="bounded: %d sites per step in batches of 1000 (param.predictBatchSize)" % sample),
                cpu_baseline=dict(value=v, unit="sites/s", cores=cores, kind="port",
                                  sample="%d sites/step x %d steps, torch-CPU fp32 restatement of the TF graph, best of probed thread counts = %d threads"
                                         % (sample, args.steps, threads)),
                e2e=dict(value=v, unit="sites/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="reference needs TensorFlow 1.12/Python 2 (absent): oracle/cv_oracle_torch.py stands in (kind=port)")
    print(json.dumps(line))


def source_hash():
    from clairvoyante_b200 import _lib
    return _lib.source_hash()


def load_traffic(variant, kernel):
    """ncu DRAM bytes per launch of `kernel` from profiles/traffic.json -- only if that file was captured from THIS build
    (it records the hash of csrc/ + include/ that tools/capture_round.sh saw); a stale file reads as null + a note"""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return None, "profiles/traffic.json absent"
    d = json.load(open(tp))
    if d.get("source_hash") != source_hash():
        return None, "profiles/traffic.json is from another build (source_hash %s != %s): re-run tools/capture_round.sh" % (
            d.get("source_hash"), source_hash())
    return d.get(variant, {}).get(kernel), "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, build %s" % d["source_hash"]


def inference_leg(args, cv, variant, W, local, rank, world, barrier, max_over_ranks, steps, warmup, full, compute=None):
    """device-resident pass + per-kernel timing + the two e2e feeds for one variant; `full` adds roofline detail"""
    import torch
    from clairvoyante_b200 import synth, utils_v2
    m = cv.Clairvoyante(device=local)
    compute = compute or args.compute
    if compute != "auto":
        m.setComputeMode(compute)
    m.setWeights(W)
    tensor = m.computeMode != "fp32"
    n = args.sites
    pool_n = min(65536, n)
    pool = synth.make_sites(pool_n, seed=1000 + rank)
    reps = (n + pool_n - 1) // pool_n
    xd = torch.from_numpy(pool).cuda().repeat(reps, 1, 1, 1)[:n].contiguous()      # 8.45 GB for 4M sites >> 126 MB L2
    od = torch.empty((n, 16), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        m.predictDevice(xd.data_ptr(), n, od.data_ptr(), None, stream)

    for _ in range(warmup):
        step()
    barrier()
    clocks = ClockSampler(local)
    l0 = m.kernelLaunches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = m.kernelLaunches() - l0
    clk = clocks.stop()
    value = world * n * steps / (ms / 1e3)
    out = dict(value=value, ms_per_step=ms / steps, launches=launches, clocks=clk, tensor=tensor, n=n, pool_n=pool_n)

    # ---- per-kernel durations, live, CUDA events on the launching stream (separate short pass so
    #      the event records do not sit inside the headline timing)
    m.profileBegin()
    for _ in range(2):
        step()
    prof = m.profileRead()
    fl = dict(FLOPS_PER_SITE[variant])
    conv2_separate = prof["conv2"][0] > 0.01 * prof["front"][0]      # tcgen05 conv2 runs as its own kernel
    fl["front"] = fl["conv1"] if conv2_separate else fl["conv1"] + fl["conv2"]
    if not conv2_separate:
        prof.pop("conv2")
    kern = {}
    for k, (tms, cnt) in prof.items():
        kern[k] = dict(ms_per_launch=tms / max(cnt, 1), launches=cnt, share=tms / max(sum(v[0] for v in prof.values()), 1e-9),
                       tflops=fl[k] * (2 * n) / (tms / 1e3) / 1e12)
    out["kernels"], out["flops"] = kern, fl
    del xd, od
    torch.cuda.empty_cache()

    # ---- e2e: public API, pinned HOST input, H2D of every input byte + D2H of every result inside the timed region.
    # One e2e step = the same `ne` sites, fed as predict() calls of at most 1 Mi sites from a pinned buffer (keeps the
    # pinned footprint per rank small -- 8 ranks share one host).  Two feeds:
    #   counts : the batch as the repo's own producers hand it out (utils_v2.GetTensor / CreateTensor.GetTensorFromAlignments
    #            yield a CountBatch: float32 tensors + the raw uint8 / int16 counts) -> cvb_predict_host_counts_*: a quarter of
    #            the fp32 bytes cross the link, the first kernel widens and subtracts the reference channel (utils_v2.py:46)
    #   fp32   : a caller that only holds the channel-subtracted float32 tensors (the reference's own callers)
    ne = args.e2e_sites or n
    call_n = min(ne, 1 << 20)
    ncalls = (ne + call_n - 1) // call_n
    ne = ncalls * call_n
    raw = synth.make_counts(pool_n, seed=1000 + rank)
    packed = utils_v2.pack_counts(raw.astype(np.float32), subtracted=False)
    assert packed is not None and np.array_equal(packed.astype(np.int16), raw)
    feeds = {}
    for name, dtype in (("counts", packed.dtype), ("fp32", np.float32)):
        xh = torch.empty((call_n, 33, 4, 4), dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
        xh_np = xh.numpy()
        src = packed if name == "counts" else pool
        for i in range(0, call_n, pool_n):
            k = min(pool_n, call_n - i)
            xh_np[i:i + k] = src[:k]
        checksum = 0.0
        for _ in range(2):
            m.predict(xh_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            for _c in range(ncalls):
                base, z, t, l = m.predict(xh_np)
                checksum += float(t[0, 0])
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        feeds[name] = dict(value=world * ne * steps / dt, unit="sites/s", h2d_bytes_per_step=world * ne * 528 * xh_np.itemsize,
                           d2h_bytes_per_step=world * ne * 16 * 4, ms_per_step=dt / steps * 1e3, checksum=checksum,
                           input_dtype=str(xh_np.dtype))
        del xh, xh_np
    e2e = dict(feeds["counts"])
    e2e.update(sites_per_step=world * ne, sites_per_gpu_per_step=ne, calls_per_step=ncalls,
               api="Clairvoyante.predict(X) on a pinned NumPy batch of RAW %s counts (what utils_v2.GetTensor / "
                   "CreateTensor.GetTensorFromAlignments attach to every batch), %d sites per call; four per-head float32 "
                   "arrays come back" % (feeds["counts"]["input_dtype"], call_n),
               fp32_input=dict(feeds["fp32"], api="Clairvoyante.predict(X) on pinned channel-subtracted float32 X"))
    out["e2e"] = e2e
    m.close()
    return out


def train_leg(args, cv, W, local, rank, world, dist, barrier, max_over_ranks):
    """BASELINE configs[3] / README.md:307-319: v3 training, fwd + loss + bwd + Adam per step through the public API
    (Clairvoyante.train at N = 1, parallel.DataParallelTrainer -- NCCL all-reduce inside the library -- at N > 1) on the
    reference's trainBatchSize = 10,000 tensors GLOBAL (strong scaling) and on 10,000 per GPU (weak); host arrays, H2D inside"""
    import torch
    from clairvoyante_b200 import parallel, synth, utils_v2
    m = cv.Clairvoyante(device=local)
    m.init(seed=1)
    m.setWeights(W)
    tr = parallel.DataParallelTrainer(m, dist if world > 1 else None)
    res = {}
    for label, total in (("global_batch_10000", 10000), ("per_gpu_batch_10000", 10000 * world)):
        if world == 1 and label == "per_gpu_batch_10000":
            res[label] = res["global_batch_10000"]
            continue
        x, y = synth.make_labeled_sites(total, seed=77)
        x = utils_v2.with_counts(x)      # what the training drivers hand over (_driver.TrainingSet.fetch): float32 + raw counts
        for _ in range(3):
            tr.train(x, y, seed=1)
        barrier()
        l0 = m.kernelLaunches()
        steps = 10
        t0 = time.perf_counter()
        for i in range(steps):
            loss, _ = tr.train(x, y, seed=2 + i)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        res[label] = dict(value=total * steps / dt, unit="tensors/s", ms_per_step=dt / steps * 1e3, global_batch=total,
                          per_gpu_batch=total // world, launches_per_step=(m.kernelLaunches() - l0) / steps, loss_per_tensor=float(loss) / total)
    m.close()
    return dict(metric="training tensors/sec (fwd + loss + bwd + TF-Adam per step, clairvoyante_v3, dropout 0.5, lambda 1e-3)",
                api="Clairvoyante.train(X, Y)" if world == 1 else "parallel.DataParallelTrainer.train (cvb_allreduce_init: ncclAllReduce on the library's stream)",
                dtype="bf16x3 on tcgen05 (split-bf16 operands, fp32 accumulate, fp32 master weights)", data="synthetic labelled sites",
                reference_published="README.md:307-319: 90 s per 11 M tensors on a V100 (122 k/s), ~3.8 k/s on 28 Xeon cores", **res)


def small_batch_leg(cv, W, local, batch=1000):
    """BASELINE configs[4]'s left end, the reference drivers' real call pattern: predictBatchSize = 1000 sites per call
    (param.py:12, callVar.py:184).  Device-resident calls, the synchronous host API on a pageable CountBatch, and the same
    feed with three calls in flight (predictSubmit / predictCollect)."""
    import time

    import torch

    from clairvoyante_b200 import synth, utils_v2
    m = cv.Clairvoyante(device=local)
    m.setWeights(W)
    x = synth.make_sites(batch, 5)
    feed = utils_v2.with_counts(x.copy())
    xd = torch.from_numpy(x).cuda()
    od = torch.empty((batch, 16), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    reps = 300
    for _ in range(10):
        m.predictDevice(xd.data_ptr(), batch, od.data_ptr(), None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        m.predictDevice(xd.data_ptr(), batch, od.data_ptr(), None, st)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / reps
    out = dict(batch=batch, config="configs[4] left end: %d sites per call (the reference's predictBatchSize)" % batch,
               device=dict(value=batch / dev_us * 1e6, unit="sites/s", us_per_call=dev_us))
    for name, arr in (("host_api_fp32_pageable", x), ("host_api_counts", feed)):
        for _ in range(5):
            m.predict(arr)
        t0 = time.perf_counter()
        for _ in range(reps):
            m.predict(arr)
        us = (time.perf_counter() - t0) * 1e6 / reps
        out[name] = dict(value=batch / us * 1e6, unit="sites/s", us_per_call=us, api="Clairvoyante.predict, one call at a time")
    for _ in m.predictStream([feed] * 8, depth=3):
        pass
    t0 = time.perf_counter()
    for _ in m.predictStream([feed] * reps, depth=3):
        pass
    us = (time.perf_counter() - t0) * 1e6 / reps
    out["host_api_counts_pipelined"] = dict(value=batch / us * 1e6, unit="sites/s", us_per_call=us,
                                            api="Clairvoyante.predictStream (cvb_predict_submit / cvb_predict_collect), 3 calls in flight")
    m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="v3", choices=["v3", "v3_slim"])
    ap.add_argument("--sites", type=int, default=4 * 1024 * 1024, help="sites per GPU per step")
    ap.add_argument("--e2e-sites", type=int, default=None, help="sites per GPU per e2e step (default: --sites)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--compute", default="auto", choices=["auto", "fp32", "fp16x3", "fp16"],
                    help="arithmetic path: fp16x3 = conv3/FC4 on tcgen05 with split-fp16 operands + fp32 accumulate "
                         "(fp32-equivalent, logits within 1e-3 of fp64); fp32 = all-SIMT fp32")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip the v3_slim and training blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d" % (args.gpus, world, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # rank 0 prints ONE JSON line on stdout:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner (printed at WARN too) goes to stderr
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from clairvoyante_b200 import initializers
    from clairvoyante_b200 import clairvoyante_v3 as cv3, clairvoyante_v3_slim as cvs
    cv = cv3 if args.variant == "v3" else cvs
    W = initializers.init_weights(args.variant, seed=0)
    r = inference_leg(args, cv, args.variant, W, local, rank, world, barrier, max_over_ranks, args.steps, args.warmup, True)
    tensor, n, kern, fl, value = r["tensor"], r["n"], r["kernels"], r["flops"], r["value"]
    dom = max(kern, key=lambda k: kern[k]["share"])
    pk = peaks()
    traffic, traffic_note = load_traffic(args.variant, dom)
    sites_per_launch = 2 * n / max(kern[dom]["launches"], 1)
    tc_kernels = ("conv2", "conv3", "fc4", "tail") if args.variant == "v3" else ("conv2", "conv3", "fc4")   # slim: conv1 and the tail stay SIMT
    on_tensor = tensor and dom in tc_kernels
    if on_tensor:
        # the kernel issues 3 fp16 MMAs per algorithmic fp32 MAC (split operands); achieved counts ALGORITHMIC flops
        r_bound, r_peak = "tensor", pk["bf16_tflops"]
        r_src = "%s bf16 dense GEMM; this kernel issues 3x the algorithmic flops as fp16 tcgen05.mma (split-fp16)" % pk["source"]
    else:
        r_bound, r_peak = "fp32_fma", FP32_NOMINAL_TFLOPS
        r_src = "nominal fp32 FMA (148 SM x 128 lanes x 2 x 1.965 GHz); MEASURED_PEAKS.json has no fp32-SIMT figure"
    for k in kern:
        kern[k]["pipe"] = "tcgen05 (3x split-fp16)" if (tensor and k in tc_kernels) else "fp32 SIMT"
    roofline = dict(kernel=dom, bound=r_bound, achieved=kern[dom]["tflops"], peak=r_peak, unit="TFLOP/s",
                    frac=kern[dom]["tflops"] / r_peak, peak_source=r_src,
                    issued_frac=(3.0 if on_tensor else 1.0) * kern[dom]["tflops"] / r_peak,   # tensor-pipe occupancy: MMAs issued / peak
                    algorithmic_flops_per_launch=fl[dom] * sites_per_launch, ms_per_launch=kern[dom]["ms_per_launch"],
                    traffic=traffic, traffic_source=traffic_note, kernels=kern,
                    whole_pass=dict(tflops=value / world * fl["total"] / 1e12, frac_fp32_nominal=value / world * fl["total"] / 1e12 / FP32_NOMINAL_TFLOPS),
                    hbm=dict(achieved_gbs=value / world * HBM_BYTES_PER_SITE / 1e9, peak_gbs=pk["hbm_gbs"],
                             frac=value / world * HBM_BYTES_PER_SITE / 1e9 / pk["hbm_gbs"], peak_source=pk["source"],
                             note="not the binding bound: 3,708 FLOP per algorithmic HBM byte"))

    slim = train = small = None
    if not args.no_extra and world == 1:
        small = small_batch_leg(cv, W, local)
    if not args.no_extra:
        if args.variant == "v3":
            # BASELINE configs[2]: v3_slim, site list sharded 1 -> N GPUs, recorded at every N the driver runs -- in the
            # fp32-equivalent arithmetic (3x split fp16, logits within 1e-3) and in the plain fp16 the config names
            slim = dict(metric=METRIC, unit="sites/s",
                        config="configs[2]: v3_slim inference, %d sites per GPU, site-list sharded, no collective" % args.sites)
            Ws = initializers.init_weights("v3_slim", seed=0)
            for mode in ("fp16x3", "fp16"):
                s = inference_leg(args, cvs, "v3_slim", Ws, local, rank, world, barrier, max_over_ranks, max(2, args.steps // 2), 2,
                                  False, compute=mode)
                slim[mode] = dict(value=s["value"], ms_per_step=s["ms_per_step"], vs_v3=s["value"] / value, e2e=s["e2e"],
                                  gpu_launches=s["launches"],
                                  tolerance="|logit - fp64| <= 1e-3" if mode == "fp16x3" else "|logit - fp64| <= 2e-3 * max |logit|",
                                  kernels={k: dict(ms_per_launch=v["ms_per_launch"], share=v["share"]) for k, v in s["kernels"].items()})
            slim["value"], slim["vs_v3"] = slim["fp16"]["value"], slim["fp16"]["vs_v3"]
            train = train_leg(args, cv3, W, local, rank, world, dist, barrier, max_over_ranks)

    cpu = None
    if rank == 0 and world == 1:
        v, cores, nsamp, thr = cpu_reference_rate(args.variant, W, args.cpu_seconds)
        cpu = dict(value=v, unit="sites/s", cores=cores, kind="port",
                   sample="%d sites in batches of 1000, torch-CPU fp32 restatement of the TF graph (oracle/cv_oracle_torch.py), best of probed thread counts = %d threads"
                          % (nsamp, thr))
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="sites/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 (3x split-fp16 on tcgen05, f32 accumulate)" if tensor else "f32",
                    data="synthetic",
                    config=dict(workload="configs[1]: 4M synthetic (33,4,4) candidate sites, clairvoyante_%s forward, fp32" % args.variant,
                                sites_per_gpu_per_step=n, unique_sites=r["pool_n"], parallelism="site-list sharding, no collective",
                                l2="inputs (%.2f GB/GPU) exceed the 126 MB L2; no flush needed" % (n * 2112 / 1e9),
                                weights="reference initialisers, seed 0",
                                compute_mode=(("fp32-equivalent: %s on tcgen05 with 3x split-fp16 operands and fp32 accumulate, "
                                               "rest fp32 SIMT" % ("conv2+conv3+FC4+FC5/heads" if args.variant == "v3" else "conv2+conv3+FC4"))
                                              if tensor else "fp32 SIMT")),
                    clocks=r["clocks"], e2e=r["e2e"], gpu_launches=r["launches"], roofline=roofline, cpu_baseline=cpu,
                    slim=slim, train=train, small_batch=small, build=source_hash())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
