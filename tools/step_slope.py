import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, load_problem_npz
import os
for name in ("demo","ladybug","trafalgar","venice"):
    p = load_problem_npz('/root/repo/tests/golden/demo_kf49.npz') if name=="demo" else make_problem(name,1)
    out=[]
    for steps in (10,30,50):
        opt=binding.default_options(max_iterations=12,min_iterations=10**6,gradient_threshold=0.0,cg_max_steps=steps,poll_every=64)
        with binding.Engine(opt) as e:
            e.set_problem_from(p); e.begin(); e.iterate(3); e.restart(); e.iterate(8); ms=e.last_iterate_ms()/8
            out.append((steps,ms))
    slope=(out[2][1]-out[0][1])/40; icpt=out[0][1]-10*slope
    print(name, [(s,round(m,3)) for s,m in out], "per PCG step us", round(slope*1e3,2), "non-PCG us", round(icpt*1e3,1))
