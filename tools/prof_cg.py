"""Tiny driver for ncu: Venice shape, one linearisation, a few launches of the two CG sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
shape = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = make_problem(shape, 1)
eng = binding.Engine(max_iterations=2)
eng.set_problem_from(p); eng.begin(); eng.debug_linearize()
for i, n in enumerate(eng.kernel_names()):
    print(n, eng.time_kernel(i, 20))
