"""configs[0] loop (sphere CMA-ME, 15 emitters x 36, n = 100), a few iterations for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_c1 import run_gpu
print(run_gpu(int(sys.argv[1]) if len(sys.argv) > 1 else 40))
