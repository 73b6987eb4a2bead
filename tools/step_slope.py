"""Time per LM iteration against the number of PCG steps (10 / 30 / 50): slope = cost of one PCG step, intercept = everything
else (linearisation, preconditioner / explicit reduced system, back-substitution, decision).

    python tools/step_slope.py [mode] [shape ...]     mode: 0 automatic plan, 1 matrix-free (options.reduced_system)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, load_problem_npz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for name in (sys.argv[2:] or ("demo", "ladybug", "trafalgar", "venice")):
    p = load_problem_npz(os.path.join(ROOT, "tests/golden/demo_kf49.npz")) if name == "demo" else make_problem(name, 1)
    out = []
    for steps in (10, 30, 50):
        opt = binding.default_options(max_iterations=12, min_iterations=10**6, gradient_threshold=0.0, cg_max_steps=steps, poll_every=64,
                                      reduced_system=mode)
        with binding.Engine(opt) as e:
            e.set_problem_from(p); e.begin(); e.iterate(3); e.restart(); e.iterate(8); ms = e.last_iterate_ms() / 8
            plan = e.plan()
            out.append((steps, ms))
    slope = (out[2][1] - out[0][1]) / 40; icpt = out[0][1] - 10 * slope
    print(name, "plan", plan["explicit_schur"], [(s, round(m, 3)) for s, m in out], "per PCG step us", round(slope * 1e3, 2), "non-PCG us", round(icpt * 1e3, 1),
          "launches", plan["launches_per_iteration"], flush=True)
