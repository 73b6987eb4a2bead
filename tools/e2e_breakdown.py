import sys, time, os
sys.path.insert(0, '/root/repo')
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
p = make_problem('venice', 1)
for rep in range(2):
    t0=time.perf_counter()
    eng = binding.Engine(max_iterations=10, min_iterations=10**6, gradient_threshold=0.0, poll_every=10)
    t1=time.perf_counter()
    eng.set_problem_from(p)
    t2=time.perf_counter()
    eng.begin(); t3=time.perf_counter()
    eng.iterate(10); t4=time.perf_counter()
    r=eng.finish(); t5=time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} set_problem {1e3*(t2-t1):.1f} begin {1e3*(t3-t2):.1f} iterate {1e3*(t4-t3):.1f} finish+get {1e3*(t5-t4):.1f} | h2d {r.ms_h2d:.1f} solve {r.ms_solve:.1f} d2h {r.ms_d2h:.1f}", flush=True)
    eng.close()
