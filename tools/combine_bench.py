"""Stage-2 row combiners over a resident codeword tile (SURVEY 8d iii): device timing + HBM roofline."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import torch

lgr = bench.load_package()
dev = torch.device("cuda", 0)
peaks, src = bench.measured_peaks()
hbm = float(peaks["hbm_gbs"])
out = {}
for k, T in ((256, 1 << 16), (8192, 1 << 11)):
    n = 4 * k
    ex = lgr.make_executor(max(k - 192, 1), k)
    a = ex.make_device_buffer(T * n * 32); b = ex.make_device_buffer(T * n * 32)
    ex.synth(a, 7, 0, T, n); ex.synth(b, 8, 0, T, n)
    acc = ex.make_codeword_buffer()
    rs = lgr.ints_to_array([(i * 0x9E3779B97F4A7C15 + 12345) % lgr.P for i in range(T)])   # converted once, outside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    z = ex.make_device_buffer(T * n * 32); ex.synth(z, 9, 0, T, n)
    for name, fn, by in (("combine_code", lambda: ex.combine_code(a, T, rs, acc), T * n * 32), ("combine_linear", lambda: ex.combine_linear(a, b, T, acc), 2 * T * n * 32),
                         ("combine_quad", lambda: ex.combine_quad(a, b, z, T, rs, acc), 3 * T * n * 32)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        reps = 5
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        out["%s_k%d_T%d" % (name, k, T)] = {"ms": t, "algorithmic_bytes": by, "achieved_gbs": by / (t * 1e-3) / 1e9, "frac_hbm": by / (t * 1e-3) / 1e9 / hbm,
                                            "note": "tile of %.1f GiB > L2, read once per sweep" % (T * n * 32 / 2**30)}
    ex.close()
print(json.dumps({"peak_gbs": hbm, "peak_source": src, "results": out}, indent=1))
