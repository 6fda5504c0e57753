#!/bin/bash
# usage (on the GPU box, from the repo root): bash profiles/tools/gpu_round_short.sh <tag>
# The end-of-round check: GPU test suite, bench line (both arms), ncu launch list of the bench command, full ncu
# captures of the headline solver kernels and of the fused-heads forward, the predict cases of bench_configs.py.
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
python bench_configs.py predict > $OUT/predict.jsonl 2> $OUT/predict.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch > $OUT/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_fwd_kernel -c 1 -o $OUT/prof_predict \
    python tests/prof_misc.py predict > $OUT/ncu_predict.log 2>&1
tail -3 $OUT/pytest_gpu.log; cut -c1-400 $OUT/bench.json; cut -c1-300 $OUT/bench_ref.json; cut -c1-420 $OUT/predict.jsonl
