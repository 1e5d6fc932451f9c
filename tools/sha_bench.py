"""Column-hash kernels alone on a resident codeword tile: ms per launch and cycles per 64-byte block per column
(1965 MHz assumed), for the chain-kernel knobs given in the environment (run on a B200)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
lgr = ge._load_package()
res = {}
for k, T in ((256, 8192), (64, 16384), (512, 4096), (1024, 2048), (8192, 512)):
    n = 4 * k
    ex = lgr.make_executor(max(k - 192, 1), k)
    tile = ex.make_device_buffer(T * n * 32)
    ex.synth(tile, 1, 0, T, n)
    sha = ex.make_device_buffer(lgr.lib().lgr_sha_ctx_bytes(n))
    lib = lgr.lib()
    import ctypes as C
    lib.lgr_sha_init(ex._ctx, sha.ptr(), C.c_uint32(n))
    def run():
        lib.lgr_sha_update_rows(ex._ctx, sha.ptr(), C.c_uint32(n), tile.ptr(), C.c_uint64(n), C.c_uint32(T))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    for _ in range(5):
        run()
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res["k%d_T%d" % (k, T)] = {"ms": ms, "cycles_per_block": ms * 1e-3 * 1.965e9 / (T / 2), "compress_per_s": n * (T / 2) / (ms * 1e-3)}
    ex.close()
print(json.dumps(res, indent=1))
