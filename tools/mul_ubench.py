"""Multiplier micro-benchmarks on a B200 (include/lgr_ubench.h): IMAD.WIDE, the IMAD Montgomery multiplication, the Shoup
variant, DFMA, and the FP64-pipe Montgomery multiplication alone and mixed with IMAD warps.  Writes gpurun_out/mul_ubench.json."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
lgr = bench.load_package()
ex = lgr.make_executor(64, 256)
names = {0: "imad_wide_per_s", 3: "imad_lo_per_s", 4: "dfma_per_s", 1: "mont_imad_per_s", 5: "shoup_per_s", 6: "mont_fp64_per_s",
         7: "mont_mixed_1imad_1fp64_per_s", 8: "mont_mixed_3imad_1fp64_per_s"}
res = {}
for w, name in names.items():
    res[name] = ex.ubench(w)
    print(name, "%.4g" % res[name])
res["fp64_over_imad"] = res["mont_fp64_per_s"] / res["mont_imad_per_s"]
res["best_mixed_over_imad"] = max(res["mont_mixed_1imad_1fp64_per_s"], res["mont_mixed_3imad_1fp64_per_s"]) / res["mont_imad_per_s"]
res["mont_vs_sha_overlap"] = ex.ubench_overlap()
print(res["mont_vs_sha_overlap"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mul_ubench.json", "w"), indent=1)
