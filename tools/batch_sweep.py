"""BASELINE.json configs[4]: batch-size sweep 256 -> 65536, v3 inference, 1 x B200: sites/s (device-resident and through
the host API).  Writes gpurun_out/batch_sweep.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clairvoyante_b200 import clairvoyante_v3 as cv, clairvoyante_v3_slim as cvs, initializers as I, synth

os.makedirs("gpurun_out", exist_ok=True)
res = {}
for variant, mod, modes in (("v3", cv, ("fp16x3", "fp32")), ("v3_slim", cvs, ("fp32",))):
    W = I.init_weights(variant, 0)
    m = mod.Clairvoyante(); m.setWeights(W)
    pool = synth.make_sites(65536, 1)
    xd = torch.from_numpy(pool).cuda()
    xh = torch.from_numpy(pool).pin_memory().numpy()
    od = torch.empty((65536, 16), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for mode in modes:
        m.setComputeMode(mode)
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
            e1.record(); torch.cuda.synchronize()
            dev = b * reps / (e0.elapsed_time(e1) / 1e3)
            m.predict(xh[:b])
            t0 = time.perf_counter()
            for _ in range(max(3, reps // 4)):
                m.predict(xh[:b])
            host = b * max(3, reps // 4) / (time.perf_counter() - t0)
            rows.append(dict(batch=b, device_sites_per_s=dev, host_api_sites_per_s=host, us_per_call_device=1e6 * b / dev))
            print(variant, mode, b, "device %.2f M/s  host-API %.2f M/s  (%.1f us/call)" % (dev / 1e6, host / 1e6, 1e6 * b / dev), flush=True)
        res["%s/%s" % (variant, mode)] = rows
    m.close()
json.dump(res, open("gpurun_out/batch_sweep.json", "w"), indent=1)
