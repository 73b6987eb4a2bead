"""PBA's LM variant (options.lm_variant = 1) on the GPU against the reference PBA's own per-iteration report (oracle/_ref).
Usage: python tools/pba_lm_check.py [shape ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200 import binding
from sequencesfm_b200.synth import make_problem, pba_inputs
from oracle import refs
for shape in sys.argv[1:] or ["small", "ladybug"]:
    p = make_problem(shape, 1)
    ref, res, _ = refs.pba_trace(p)
    q = pba_inputs(p)
    with binding.Engine(binding.default_options("pba_lm")) as eng:
        eng.set_problem_from(q)
        r = eng.solve()
    mse = lambda c: c * q.K5[0] ** 2 / q.K
    print(f"== {shape}: PBA initial mse {res.initial_mse:.6g} final {res.final_mse:.6g} code {res.return_code} | ours initial {mse(r.trace[0, 1]):.6g} final {mse(r.final_cost):.6g} status {r.status}")
    for i in range(max(len(ref), len(r.trace))):
        a = ref[i] if i < len(ref) else None
        b = r.trace[i] if i < len(r.trace) else None
        sa = f"cg {int(a[1]):3d} e={a[2]:.6g} u={a[3]:.3g} r={a[4]:.3f} acc {int(a[5])}" if a is not None else "-"
        sb = f"cg {int(b[6]):3d} e={mse(b[3]):.6g} u={b[2]:.3g} r={b[4]:.3f} acc {int(b[5])}" if b is not None else "-"
        print(f"  #{i:2d}  PBA {sa:55s} | ours {sb}")
