"""BASELINE config 2: 2^20-point NTT then iNTT on one B200 (device timing, L2 flushed between launches)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import torch

lgr = bench.load_package()
dev = torch.device("cuda", 0)
ex = lgr.make_executor(64, 256)
peaks, src = bench.measured_peaks()
hbm = float(peaks["hbm_gbs"])
out = {}
for logn in (12, 16, 20, 22):
    N = 1 << logn
    buf = ex.make_device_buffer(N * 32)
    ex.synth(buf, 2, 0, 1, N)
    w = lgr.root_of_unity(logn)
    for _ in range(3):
        ex.ntt_pow2(buf, logn, 1, w, False); ex.ntt_pow2(buf, logn, 1, w, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tf = ti = 0.0
    reps = 10
    for _ in range(reps):
        flush.fill_(1); e0.record(); ex.ntt_pow2(buf, logn, 1, w, False); e1.record(); torch.cuda.synchronize(); tf += e0.elapsed_time(e1)
        flush.fill_(2); e0.record(); ex.ntt_pow2(buf, logn, 1, w, True); e1.record(); torch.cuda.synchronize(); ti += e0.elapsed_time(e1)
    by = 2 * N * 32
    for name, t in (("forward", tf / reps), ("inverse", ti / reps)):
        out["ntt_2^%d_%s" % (logn, name)] = {"ms": t, "algorithmic_bytes": by, "achieved_gbs": by / (t * 1e-3) / 1e9, "frac_hbm": by / (t * 1e-3) / 1e9 / hbm,
                                              "mulmods_per_s": (N // 2) * logn / (t * 1e-3)}
print(json.dumps({"peak_gbs": hbm, "peak_source": src, "results": out}, indent=1))
