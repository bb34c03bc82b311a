"""BASELINE.json configs[4]: batch-size sweep 256 -> 65536, v3 (and v3_slim) inference, 1 x B200: sites/s device-resident and
through the host API (pageable float32 as the reference's callers hold it, the repo's own CountBatch feed, and that feed with
three batches in flight through predictSubmit / predictCollect), per-kernel
duration at every batch, and the achieved fraction of the roofline per batch -- the dominant kernel's algorithmic TFLOP/s over
the measured dense bf16 figure (the pass is compute-bound: 3,708 FLOP per algorithmic HBM byte), with the HBM view beside it.
Writes gpurun_out/batch_sweep.json (kept as profiles/rNN_batch_sweep.json).    python tools/batch_sweep.py [variant ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import FLOPS_PER_SITE, HBM_BYTES_PER_SITE, peaks  # noqa: E402
from clairvoyante_b200 import clairvoyante_v3 as cv, clairvoyante_v3_slim as cvs, initializers as I, synth, utils_v2  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
pk = peaks()
res = dict(peaks=pk, note="frac = algorithmic TFLOP/s of the pass (and of its dominant kernel) / measured sustained dense bf16 "
                          "TFLOP/s; the tensor kernels issue 3x the algorithmic flops (split fp16), so 0.33 is their ceiling")
want = sys.argv[1:] or ["v3", "v3_slim"]
for variant, mod in (("v3", cv), ("v3_slim", cvs)):
    if variant not in want:
        continue
    fl = FLOPS_PER_SITE[variant]
    m = mod.Clairvoyante()
    m.setWeights(I.init_weights(variant, 0))
    pool = synth.make_sites(65536, 1)
    counts = utils_v2.pack_counts(pool)
    xd = torch.from_numpy(pool).cuda()
    od = torch.empty((65536, 16), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for b in (256, 512, 1000, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        reps = max(4, min(200, (1 << 21) // b))
        for _ in range(3):
            m.predictDevice(xd.data_ptr(), b, od.data_ptr(), None, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            m.predictDevice(xd.data_ptr(), b, od.data_ptr(), None, st)
        e1.record()
        torch.cuda.synchronize()
        dev = b * reps / (e0.elapsed_time(e1) / 1e3)
        m.profileBegin()
        for _ in range(4):
            m.predictDevice(xd.data_ptr(), b, od.data_ptr(), None, st)
        prof = {k: v[0] / max(v[1], 1) for k, v in m.profileRead().items() if v[1] and v[0] > 0}
        host = {}
        for name, arr in (("fp32_pageable", pool[:b].copy()), ("counts", utils_v2.with_counts(pool[:b].copy()))):
            m.predict(arr)
            k = max(5, reps // 4)
            t0 = time.perf_counter()
            for _ in range(k):
                m.predict(arr)
            host[name] = b * k / (time.perf_counter() - t0)
        if b <= 16384:      # the same feed with three batches in flight (predictSubmit / predictCollect)
            arr = utils_v2.with_counts(pool[:b].copy())
            k = max(8, reps // 2)
            for _ in m.predictStream([arr] * 4, depth=3):
                pass
            t0 = time.perf_counter()
            for _ in m.predictStream([arr] * k, depth=3):
                pass
            host["counts_pipelined"] = b * k / (time.perf_counter() - t0)
        dom = max(prof, key=prof.get)
        chunks = -(-b // 18944) if variant == "v3" else -(-b // 33152)
        fdom = fl["conv1"] + fl["conv2"] if (dom == "front" and "conv2" not in prof) else fl.get(dom, fl["conv1"])
        row = dict(batch=b, device_sites_per_s=dev, us_per_call_device=1e6 * b / dev, host_api_sites_per_s=host,
                   kernel_us={k: round(v * 1e3, 2) for k, v in prof.items()}, dominant_kernel=dom,
                   roofline=dict(pass_tflops=dev * fl["total"] / 1e12, pass_frac=dev * fl["total"] / 1e12 / pk["bf16_tflops"],
                                 dominant_tflops=fdom * (b / chunks) / (prof[dom] / 1e3) / 1e12,
                                 dominant_frac=fdom * (b / chunks) / (prof[dom] / 1e3) / 1e12 / pk["bf16_tflops"],
                                 hbm_gbs=dev * HBM_BYTES_PER_SITE / 1e9, hbm_frac=dev * HBM_BYTES_PER_SITE / 1e9 / pk["hbm_gbs"]))
        rows.append(row)
        print(variant, b, "device %.2f M/s (%.1f us/call)  host fp32 %.2f M/s  counts %.2f M/s  pipelined %.2f M/s  pass frac %.3f  %s" %
              (dev / 1e6, 1e6 * b / dev, host["fp32_pageable"] / 1e6, host["counts"] / 1e6,
               host.get("counts_pipelined", 0) / 1e6, row["roofline"]["pass_frac"], row["kernel_us"]), flush=True)
    res[variant] = rows
    m.close()
json.dump(res, open("gpurun_out/batch_sweep.json", "w"), indent=1)
