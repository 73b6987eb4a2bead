"""Profiling aid: cycles per phase of one PCG step inside k_dense_pcg (library built with -DB200BA_STAMPS).
Usage: B200BA_LIB=sequencesfm_b200/variants/lib_dstamps.so python tools/dense_stamps.py [shape]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem, load_problem_npz
from sequencesfm_b200 import binding
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for shape in (sys.argv[1:] or ["demo", "trafalgar"]):
    p = load_problem_npz(os.path.join(ROOT, "tests/golden/demo_kf49.npz")) if shape == "demo" else make_problem(shape, 1)
    eng = binding.Engine(max_iterations=3, min_iterations=10**6, gradient_threshold=0.0)
    eng.set_problem_from(p); eng.begin(); eng.iterate(3)
    buf = np.zeros(512 * 40, dtype=np.uint64)
    rc = eng.lib.b200ba_debug_stamps(eng.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int32(buf.size))
    assert rc == 0
    ph = buf[:10].astype(np.float64); steps = int(buf[10])
    names = ["S p + sync", "p.q: slot sum", "x, r, z = M r", "z, r.z: block sum + send + wait + slot sum", "beta, p + sync", "p.q: block sum", "p.q: send + wait", "", "", "loop top"]
    print(shape, "plan", eng.plan()["explicit_schur"], "steps", steps, "cycles per step", round(ph.sum() / max(steps, 1)))
    for n, v in zip(names, ph):
        if n: print(f"   {n:28s} {v / max(steps, 1):8.0f}")
    eng.close()
