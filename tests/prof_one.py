"""One configuration, few repetitions (for ncu): python tests/prof_one.py method adjoint(0/1) [B] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_perf_probe as p
method = sys.argv[1]; adjoint = bool(int(sys.argv[2]))
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
p.run(B, 100, 15, 25, 5, method, adjoint, reps=reps)
