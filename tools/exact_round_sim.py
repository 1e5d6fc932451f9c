"""One rank's share of the exact multi-GPU layout at G ranks, replayed on ONE GPU: per round encode T rows slab-major and absorb
G chunks of T rows x n/G columns (what arrives from the G ranks) -- the kernel mix that decides exact_k8192's scaling, without
needing G GPUs to tune it.  python tools/exact_round_sim.py [k] [G] [rounds]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import torch

lgr = bench.load_package()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 16
n = 4 * k
slab = n // G
T = max(2, (1 << 23) // n)
dev = torch.device("cuda", 0)
ex = lgr.make_executor(k - 192, k)
rows = ex.make_device_buffer(T * k * 32)
ex.synth(rows, 3, 0, T, k)
send = [torch.empty(T * n * 8, dtype=torch.int32, device=dev) for _ in range(2)]
enc_s, hash_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev, priority=-1)
enc_done = [torch.cuda.Event() for _ in range(2)]
hash_done = [torch.cuda.Event() for _ in range(2)]
hash_go = [torch.cuda.Event() for _ in range(2)]
GATE = os.environ.get("LGR_EXACT_GATE", "1") != "0"      # release encode(r+1) together with hash(r) so that the hash CTAs are placed first
ex.sha256_init(slab)
ctx = ex.make_device_buffer(ex.sha256_context_bytes(slab))
dig = ex.make_device_buffer(slab * 32)
bind = ex.bind_sha256_context(ctx, dig)
ex.sha256_digest_init(bind)
torch.cuda.synchronize()


ONLY = os.environ.get("LGR_SIM_ONLY", "")            # "enc" / "hash": time one side alone


def run(nr):
    for r in range(nr):
        b = r & 1
        with torch.cuda.stream(enc_s):
            if r >= 2:
                enc_s.wait_event(hash_done[b])
            if GATE and r >= 1:
                enc_s.wait_event(hash_go[(r - 1) & 1])
            ex.use_torch_stream()
            base = send[b].data_ptr()
            if ONLY != "hash":
                ex.encode_rows_slabs(rows, T, [base + h * T * slab * 32 for h in range(G)])
            enc_done[b].record(enc_s)
        with torch.cuda.stream(hash_s):
            hash_s.wait_event(enc_done[b])
            hash_go[b].record(hash_s)
            ex.use_torch_stream()
            ex.sha256_init(slab)
            v = send[b].view(G, T * slab * 8)
            if ONLY != "enc":
                if os.environ.get("LGR_SIM_ONE_LAUNCH", "1") != "0":     # the G chunks are one contiguous [G*T][slab] matrix
                    ex.sha256_digest_update_rows(bind, ex.wrap(send[b]), G * T, slab)
                else:
                    for h in range(G):
                        ex.sha256_digest_update_rows(bind, ex.wrap(v[h]), T, slab)
            hash_done[b].record(hash_s)
    torch.cuda.synchronize()


run(4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(rounds)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / rounds
print(json.dumps({"k": k, "G": G, "tile_rows": T, "slab_columns": slab, "ms_per_round": ms, "projected_elements_per_s_at_G": G * T * k / (ms * 1e-3)}))
