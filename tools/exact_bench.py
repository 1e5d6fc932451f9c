"""Strong-scaling timing of the exact multi-GPU layout alone (bench.run_exact), for tuning: torchrun --nproc-per-node G tools/exact_bench.py [k] [log_rows]"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lgr = bench.load_package()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
log_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 15
dev = torch.device("cuda", local)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    out = bench.run_exact(lgr, torch, dist, dev, stream, rank, world, k, 1 << log_rows, 6554.2, steps=5)
if rank == 0:
    print(json.dumps({kk: out[kk] for kk in ("k", "rows_total", "transport", "value", "ms_per_step", "speedup_vs_single_gpu", "root_equals_single_gpu")}), flush=True)
dist.barrier()
dist.destroy_process_group()
