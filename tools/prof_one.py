"""ncu driver: Venice shape, one linearisation, then `reps` launches of selected kernel indices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
which = [int(x) for x in sys.argv[1].split(",")]
shape = sys.argv[2] if len(sys.argv) > 2 else "venice"
p = make_problem(shape, 1)
eng = binding.Engine(max_iterations=2, pcg_fp32=int(os.environ.get("PROF_FP32", "1")))      # the fp64 kernels are the same with the fp32 copies allocated; index 12 (cg_step) runs the fp32 sweeps unless PROF_FP32=0
eng.set_problem_from(p); eng.begin(); eng.debug_linearize()
names = eng.kernel_names()
for i in which:
    print(names[i], eng.time_kernel(i, 1))      # 3 warm-up launches + 1
