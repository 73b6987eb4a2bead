"""Golden per-iteration reports of the UNMODIFIED reference PBA (oracle/_ref/libpba_ref.so, --double -schur, one thread so that
its sums are reproducible) for the PBA LM variant of the engine (options.lm_variant = 1).

    python tools/make_golden_pba.py        # -> tests/golden/pba_lm_<shape>_s<seed>.json   (build container only)
Rows: iteration, PCG steps, e = mean squared error in px^2 after the step (6 digits, as PBA prints it), u = damping used,
r = gain ratio (null when rejected), accepted."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import make_problem, input_checksum
from oracle import refs

for shape, seed in (("small", 1), ("ladybug", 1), ("trafalgar", 1), ("ladybug", 3)):
    p = make_problem(shape, seed)
    tr, res, text = refs.pba_trace(p)
    fx = {"engine": "reference PBA (Thirdparty/pba, -DPBA_NO_GPU, unmodified) --double -schur -tnum 1 via oracle/pba_ref_harness.cpp",
          "shape": shape, "seed": seed, "N": p.N, "M": p.M, "K": p.K, "input_sha256": input_checksum(p),
          "initial_mse": res.initial_mse, "final_mse": res.final_mse, "lm_iterations": res.lm_iterations, "cg_iterations": res.cg_iterations,
          "return_code": res.return_code, "trace_cols": ["iter", "cg_steps", "mse_after", "damping", "gain_ratio", "accepted"],
          "trace": [[None if np.isnan(v) else float(v) for v in row] for row in tr]}
    json.dump(fx, open(os.path.join(ROOT, "tests", "golden", f"pba_lm_{shape}_s{seed}.json"), "w"), indent=1)
    print(shape, seed, "rows", len(tr), "mse", res.initial_mse, "->", res.final_mse, res.return_code)
