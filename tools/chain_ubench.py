"""Cycles per SHA-256 compression of one warp for every round formulation in csrc/ubench.cu (run on a B200)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
lgr = ge._load_package()
ex = lgr.make_executor(64, 256)
res = {}
for v in range(3, 26):
    for wpc in (1, 4):
        res["variant%d_warps%d" % (v, wpc)] = ex.ubench_chain(v, wpc, 32)
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/chain_ubench.json", "w"), indent=1)
