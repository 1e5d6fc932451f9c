#!/bin/bash
# One GPU session: parity of every chain-kernel variant, the lone-warp micro-benchmarks, and a short
# bench per variant (run under gpurun; outputs in gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest default: $?" | tee gpurun_out/sweep.log
for v in 1 2 3 10 11 12 13; do
  LGR_CHAIN_VARIANT=$v python -m pytest tests -m gpu -x -q -k "encode_commit_pipeline or sha_leaf or stage1_stage2" > gpurun_out/pytest_gpu_v$v.log 2>&1
  echo "pytest variant $v: $?" | tee -a gpurun_out/sweep.log
done
python tools/chain_ubench.py > gpurun_out/chain_ubench.log 2>&1
for v in 0 1 2 3 10 11 12 13; do
  LGR_CHAIN_VARIANT=$v python bench.py --log-rows 19 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY | tee -a gpurun_out/sweep.log
import json
try:
    d = json.loads(open("gpurun_out/bench_v$v.json").read().strip().splitlines()[-1])
    print("variant $v value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), {k: (round(x["ms_per_launch"], 4), round(x["share_of_step"], 3)) for k, x in d["kernels"].items()})
except Exception as e:
    print("variant $v failed", e)
PY
done
LGR_NO_SYSTEMATIC=1 python bench.py --log-rows 19 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_nosys.json 2>&1
python bench.py --log-rows 19 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --aux > gpurun_out/bench_aux.json 2>&1
LGR_NO_SYSTEMATIC=1 python bench.py --log-rows 19 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --aux > gpurun_out/bench_aux_nosys.json 2>&1
tail -c 1500 gpurun_out/pytest_gpu.log
cat gpurun_out/chain_ubench.log | tail -30
