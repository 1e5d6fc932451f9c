"""2+ GPU check of the bit-exact multi-GPU layout (sharding.commit_exact) against a single-GPU commit of
the same global matrix.  Run: torchrun --nproc-per-node G tools/exact_check.py [k] [tile_rows] [total_rows]"""
import importlib.util, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lgr = bench.load_package()
spec = importlib.util.spec_from_file_location("lgr_sharding", os.path.join(ROOT, "ligero-prover_b200", "sharding.py"))
sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
total = int(sys.argv[3]) if len(sys.argv) > 3 else 64 * 5 + 17
n = 4 * k
ex = lgr.Executor(local); ex.ntt_init(max(k - 192, 1), k, n)
num_tiles = (total + T - 1) // T
mine = sh.tiles_of_rank(num_tiles, world, rank)
bufs = []
for t in mine:
    rows = min(T, total - t * T)
    b = ex.make_device_buffer(T * k * 32)
    ex.synth(b, 3, t * T, rows, k)                       # global rows [t*T, t*T+rows)
    bufs.append((b, rows))
eng = sh.make_gpu_engine(ex, T, world, rank, dist if world > 1 else None) if world > 1 else sh.GpuEngine(ex, T, world)
leaves = sh.commit_exact(eng, lambda i: bufs[i], total, T, world, rank, dist if world > 1 else None)
nodes = ex.make_device_buffer((2 * n - 1) * 32)
ex.merkle_build(ex.wrap(leaves.contiguous()), n, nodes)
root = ex.copy_to_host(nodes, np.uint8)[:32].tobytes().hex()
# single-GPU reference commitment of the whole matrix
whole = ex.make_device_buffer(total * k * 32)
ex.synth(whole, 3, 0, total, k)
d1 = ex.make_device_buffer(n * 32); n1 = ex.make_device_buffer((2 * n - 1) * 32)
ex.encode_commit(whole, total, d1, n1)
want = ex.copy_to_host(n1, np.uint8)[:32].tobytes().hex()
ok = root == want and np.array_equal(ex.copy_to_host(d1, np.uint8), leaves.cpu().numpy().view(np.uint8).reshape(-1))
print("rank", rank, eng.transport, "exact layout root", root[:16], "single-GPU root", want[:16], "MATCH" if ok else "MISMATCH", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
