for lib in "" base rows2; do
  if [ -n "$lib" ]; then export SLODE_B200_LIB=$PWD/structured_latent_odes_b200/csrc/ab/libslode_$lib.so; else unset SLODE_B200_LIB; fi
  echo "## ${lib:-default(2,1)}"
  python - <<'PY'
import sys; sys.path.insert(0,'tests'); import gpu_perf_probe as p
for m,a in (("rk4",False),("midpoint",True),("midpoint",False),("euler",True)):
    p.run(1<<20,100,15,25,5,m,a,reps=8)
PY
done
