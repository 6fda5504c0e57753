#!/bin/bash
# usage (GPU box, repo root): bash profiles/tools/ab_prefetch.sh
# A/B of reverse-sweep variants in ONE run: the default library against variant libraries built with
# profiles/tools/build_variant.sh <tag> 25_5 -DSLODE_FX_ROW_AHEAD=1 ... and copied to
# structured_latent_odes_b200/csrc/ab/ (build/ does not travel with gpurun snapshots).  Timing: autograd-level CUDA events
# (tests/gpu_perf_probe.py), 2^20 x 100.  Output of the round-2 run: profiles/r02/ab_row_prefetch_depth.txt.
for lib in "" base rows2; do
  if [ -n "$lib" ]; then export SLODE_B200_LIB=$PWD/structured_latent_odes_b200/csrc/ab/libslode_$lib.so; else unset SLODE_B200_LIB; fi
  echo "## ${lib:-default(2,1)}"
  python - <<'PY'
import sys; sys.path.insert(0,'tests'); import gpu_perf_probe as p
for m,a in (("rk4",False),("midpoint",True),("midpoint",False),("euler",True)):
    p.run(1<<20,100,15,25,5,m,a,reps=8)
PY
done
