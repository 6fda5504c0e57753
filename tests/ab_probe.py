"""A/B timing of library variants (diagnostic): python tests/ab_probe.py B1,B2 method [lib ...]"""
import os, subprocess, sys
Bs = sys.argv[1]; method = sys.argv[2]; libs = sys.argv[3:] or [""]
code = ("import sys; sys.path.insert(0,'tests'); import gpu_perf_probe as p\n"
        "for B in [int(b) for b in '%s'.split(',')]: p.run(B,100,15,25,5,'%s',False, reps=8)\n" % (Bs, method))
for lib in libs:
    env = dict(os.environ)
    if lib: env["SLODE_B200_LIB"] = os.path.abspath(lib)
    print("##", lib or "default", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
