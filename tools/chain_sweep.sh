#!/bin/bash
# One GPU session: parity + short bench of the chain-kernel knobs, and the lone-warp micro-benchmarks
# (run under gpurun; outputs in gpurun_out/).
mkdir -p gpurun_out
: > gpurun_out/sweep.log
run() {  # name, env...
  local name=$1; shift
  env "$@" python -m pytest tests -m gpu -x -q -k "encode_commit_pipeline or sha_leaf or stage1_stage2 or prove" > gpurun_out/pytest_$name.log 2>&1
  echo "pytest $name: $? $(tail -1 gpurun_out/pytest_$name.log)" | tee -a gpurun_out/sweep.log
  env "$@" python bench.py --log-rows 19 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY | tee -a gpurun_out/sweep.log
import json
try:
    d = json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$name value %.4g ms %.2f" % (d["value"], d["ms_per_step"]), {k: (round(x["ms_per_launch"], 4), round(x["share_of_step"], 3)) for k, x in d["kernels"].items()})
except Exception as e:
    print("$name failed", e)
PY
}
run split8 LGR_CHAIN_SPLIT=1 LGR_CHAIN_GROUP=8
run split4 LGR_CHAIN_SPLIT=1 LGR_CHAIN_GROUP=4
run split1 LGR_CHAIN_SPLIT=1 LGR_CHAIN_GROUP=1
run cols32 LGR_CHAIN_SPLIT=0
run cols32_textbook LGR_CHAIN_SPLIT=0 LGR_CHAIN_TEXTBOOK=1
