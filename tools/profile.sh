#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on one B200; outputs land in gpurun_out/).
#   1. launch list (device time per launch, cold-cache + serialised: compare SHARES) of the bench command
#   2. one --set full capture each of the encoder, the hash chain, the 2^20 tile NTT and the tile combiner
#   3. the micro-benchmark / component timings the design numbers come from (not under ncu)
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 40 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
$NCU --set full --import-source on -k regex:encode_rows_kernel -s 6 -c 2 -f -o gpurun_out/prof_encode \
    python bench.py --steps 1 --warmup 3 --log-rows 15 --no-e2e --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:sha_chain -s 6 -c 2 -f -o gpurun_out/prof_sha_chain \
    python bench.py --steps 1 --warmup 3 --log-rows 15 --no-e2e --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on --kernel-name-base demangled -k "regex:ntt_tile_kernel<.int.10>" -s 12 -c 2 -f -o gpurun_out/prof_ntt \
    python tools/ntt_bench.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:combine_partial -s 2 -c 1 -f -o gpurun_out/prof_combine \
    python tools/combine_bench.py > /dev/null 2>&1
python tools/ntt_bench.py > gpurun_out/ntt_bench.json 2> /dev/null
python tools/combine_bench.py > gpurun_out/combine_bench.json 2> /dev/null
python tools/chain_ubench.py > /dev/null 2>&1          # writes gpurun_out/chain_ubench.json
python tools/sha_bench.py > gpurun_out/sha_bench.json 2> /dev/null
python tools/encode_bench.py 2> /dev/null | tail -1 > gpurun_out/encode_bench.json
ls -la gpurun_out
