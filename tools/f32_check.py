"""fp32-PCG mode at the Venice shape: time per LM iteration and cost trajectory against the fp64 path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
p = make_problem(sys.argv[1] if len(sys.argv) > 1 else "venice", 1)
out = {}
for mode in (0, 1):
    eng = binding.Engine(max_iterations=11, min_iterations=10**6, gradient_threshold=0.0, poll_every=64, pcg_fp32=mode)
    eng.set_problem_from(p); eng.begin(); eng.iterate(3); eng.restart()
    eng.iterate(8); ms = eng.last_iterate_ms() / 8
    r = eng.finish()
    if mode == 1:
        eng.begin(); eng.iterate(1)
        names = eng.kernel_names()
        print({n: round(1e3 * eng.time_kernel(names.index(n), 20), 1) for n in ('cg_point_pass_f32', 'cg_camera_pass_f32', 'cg_step', 'cg_camera_step_fused')})
    out[mode] = (ms, r.trace[:, 1].copy(), r.mean_px)
    eng.close()
n = min(len(out[0][1]), len(out[1][1]))
rel = np.abs(out[0][1][:n] - out[1][1][:n]) / out[0][1][:n]
print(f"fp64 {out[0][0]:.3f} ms/it   fp32-PCG {out[1][0]:.3f} ms/it   max rel cost diff {rel.max():.2e}  mean px {out[0][2]} {out[1][2]}")
