"""Sweep of the bank-conflict-aware grouping window (options.lane_window): host cost vs by-point sweep time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
p = make_problem("venice", 1)
for W in (64, 0, 8, 16, 32, 64, 128):
    eng = binding.Engine(max_iterations=2, lane_window=W)
    t0 = time.perf_counter(); eng.set_problem_from(p); t1 = time.perf_counter()
    eng.begin(); eng.debug_linearize()
    names = eng.kernel_names()
    print(f"lane_window {W:4d}: set_problem {1e3*(t1-t0):6.1f} ms  cg_point_pass {1e3*eng.time_kernel(names.index('cg_point_pass'), 20):6.1f} us", flush=True)
    eng.close()
