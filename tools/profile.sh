#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on one B200; outputs land in gpurun_out/prof/).
#   1. launch list (device time per launch, cold-cache + serialised: compare SHARES) of the bench command
#   2. one --set full capture each of the kernels DESIGN.md quotes: fused encoder, hash chain, 2^20 tile NTT, tile
#      combiners (code, quad), wide-matrix hash, latency NTT of the per-row path, and the Montgomery micro-benchmark
#      (the reference point for "x % of the multiplier ceiling")
#   3. the micro-benchmark / component timings the design numbers come from (not under ncu)
set -x
OUT=gpurun_out/prof
mkdir -p $OUT
NCU="ncu --clock-control none"
B="python bench.py --no-e2e --no-cpu-baseline --no-exact"
$NCU --metrics gpu__time_duration.sum -s 40 -c 400 --csv --log-file $OUT/launches.csv $B --steps 2 --warmup 3 > $OUT/bench_under_ncu.log 2>&1
$NCU --set full --import-source on -k regex:encode_rows_kernel -s 6 -c 1 -f -o $OUT/prof_encode $B --steps 1 --warmup 3 --log-rows 15 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:sha_chain -s 6 -c 1 -f -o $OUT/prof_sha_chain $B --steps 1 --warmup 3 --log-rows 15 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:sha_update_kernel -s 6 -c 1 -f -o $OUT/prof_sha_update $B --steps 1 --warmup 3 --k 8192 --log-rows 12 > /dev/null 2>&1
$NCU --set full --import-source on --kernel-name-base demangled -k "regex:ntt_tile_kernel<.int.10>" -s 12 -c 1 -f -o $OUT/prof_ntt python tools/ntt_bench.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:combine_partial -s 2 -c 1 -f -o $OUT/prof_combine_code python tools/combine_bench.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:combine_quad_kernel -s 1 -c 1 -f -o $OUT/prof_combine_quad python tools/combine_bench.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ntt_lat_kernel -s 40 -c 1 -f -o $OUT/prof_ntt_lat tests/cpp/per_row_bench 30 2 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ubench_mont_kernel -s 1 -c 1 -f -o $OUT/prof_ubench_mont python tools/mul_ubench.py > /dev/null 2>&1
python tools/ntt_bench.py > $OUT/ntt_bench.json 2> /dev/null
python tools/combine_bench.py > $OUT/combine_bench.json 2> /dev/null
python tools/chain_ubench.py > /dev/null 2>&1 && cp gpurun_out/chain_ubench.json $OUT/
python tools/sha_bench.py > $OUT/sha_bench.json 2> /dev/null
python tools/encode_bench.py 2> /dev/null | tail -1 > $OUT/encode_bench.json
python tools/mul_ubench.py > /dev/null 2>&1 && cp gpurun_out/mul_ubench.json $OUT/
tests/cpp/per_row_bench 4096 512 > $OUT/per_row.json 2> /dev/null
ls -la $OUT
