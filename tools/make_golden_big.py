"""Golden per-iteration traces of the throughput mode (fixed 50-step PCG) at the benchmark's big shapes, computed by the CPU
restatement oracle/ba_oracle.c (itself pinned against the reference SSBA build, tests/test_oracle.py).

    python tools/make_golden_big.py venice final camheavy        # -> tests/golden/trace50_<shape>_s1.json

The fixtures are what `bench.py`'s `parity` block and tests/test_gpu_parity.py compare the CUDA path with at sizes where
running the oracle inside every bench call would take minutes (Venice: 11 s per LM iteration on one host core, Final: 70 s).
Each fixture stores the generator arguments, a checksum of the generated inputs and, per LM iteration, cost / lambda /
trial cost / gain ratio / accept flag / PCG steps.
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200.synth import make_problem, BIG_CASES, input_checksum
from oracle import oracle


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name in sys.argv[1:] or list(BIG_CASES):
        shape, seed, iters = BIG_CASES[name]
        t = time.time()
        p = make_problem(shape, seed)
        # the options of bench.py's timed loop: every iteration a full LM iteration, no early stop
        r = oracle.solve(p, max_iterations=iters, solver=0, cg_max_steps=50, cg_rel_tol=0.0, min_iterations=10 ** 6, gradient_threshold=0.0)
        g = oracle.linearize(p)
        fx = {"engine": "oracle/ba_oracle.c (CPU restatement, fixed 50-step PCG = bench.py's workload)", "case": name,
              "shape": list(shape) if not isinstance(shape, str) else shape, "seed": seed, "N": p.N, "M": p.M, "K": p.K,
              "input_sha256": input_checksum(p), "iterations": int(r.iterations), "mean_px": list(r.mean_px),
              "trace_cols": ["iter", "err", "lambda", "new_err", "rho", "accepted", "cg_steps", "cg_rel"],
              "trace": [[None if np.isnan(v) else float(v) for v in row] for row in r.trace],
              "lin": {"cost": g["cost"], "lambda0": g["lambda0"], "gc_sum": float(g["gc"].sum()), "gc_abs_sum": float(np.abs(g["gc"]).sum()),
                      "gp_abs_sum": float(np.abs(g["gp"]).sum()), "U_trace_sum": float(np.einsum("nii->", g["U"])),
                      "V_trace_sum": float(np.einsum("nii->", g["V"])), "w_sum": float(g["w"].sum()), "outliers": int((g["w"] != 1.0).sum())}}
        json.dump(fx, open(os.path.join(out_dir, f"trace50_{name}_s{seed}.json"), "w"), indent=1)
        print(name, p.N, p.M, p.K, "cost", [row[1] for row in fx["trace"]], f"{time.time() - t:.0f}s", flush=True)


if __name__ == "__main__":
    main()
