"""Encoder-only timing: lgr_encode_rows on a resident batch (device time per launch)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import torch

lgr = bench.load_package()
out = {}
for k, R in ((256, 16384), (1024, 4096), (2048, 2048), (8192, 512)):
    n = 4 * k
    ex = lgr.make_executor(max(k - 192, 1), k)
    src = ex.make_device_buffer(R * k * 32); dst = ex.make_device_buffer(R * n * 32)
    ex.synth(src, 3, 0, R, k)
    for _ in range(3):
        ex.encode_rows(src, R, dst)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        ex.encode_rows(src, R, dst)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    import math
    lk = math.log2(k)
    mm = R * k * ((lk - 1) / 2 + 3 + 3 * (lk - 1) / 2)          # iNTT_k + 3 computed cosets (nominal, k <= 2048; the tile engine adds four-step twists)
    out["k%d" % k] = {"rows": R, "ms": ms, "elements_per_s": R * k / (ms * 1e-3), "nominal_montmul_per_s": mm / (ms * 1e-3)}
    ex.close()
print(json.dumps(out))
