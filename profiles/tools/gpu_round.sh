#!/bin/bash
# usage (on the GPU box, from the repo root): bash profiles/tools/gpu_round.sh <tag>
# One measurement round: GPU test suite, bench line (both arms), ncu launch list of the bench command, one full ncu
# capture of the two solver kernels of the headline configuration.  Everything lands in gpurun_out/<tag>/.
OUT=gpurun_out/$1; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py --steps 100 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch > $OUT/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fixed_ -c 2 -o $OUT/prof_rk4 \
    python tests/prof_one.py rk4 0 > $OUT/ncu_full.log 2>&1
for k in cvs heads dopri5; do
  ncu --set full --clock-control none -k regex:"cvs_|heads_|dopri5_" -c 2 -o $OUT/prof_$k \
      python tests/prof_misc.py $k > $OUT/ncu_$k.log 2>&1
  ncu -i $OUT/prof_$k.ncu-rep --page raw --csv > $OUT/prof_${k}_raw.csv 2>/dev/null && rm -f $OUT/prof_$k.ncu-rep  # 64 MiB pull limit
done
python bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err
tail -3 $OUT/pytest_gpu.log; cut -c1-400 $OUT/bench.json; cut -c1-300 $OUT/bench_ref.json
