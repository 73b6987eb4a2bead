"""Profiling aid: per-CTA / per-warp clock stamps of the two PCG sweeps (library built with -DB200BA_STAMPS, see
tools/build_variants.sh).  Usage: B200BA_LIB=.../lib_stamps.so python tools/stamps.py [shape]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sequencesfm_b200.synth import make_problem
from sequencesfm_b200 import binding
shape = sys.argv[1] if len(sys.argv) > 1 else "venice"
p = make_problem(shape, 1)
eng = binding.Engine(max_iterations=2)
eng.set_problem_from(p); eng.begin(); eng.debug_linearize()
names = eng.kernel_names()
print("cg_step ms", eng.time_kernel(names.index("cg_step"), 10))
buf = np.zeros(512 * 40, dtype=np.uint64)
rc = eng.lib.b200ba_debug_stamps(eng.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int32(buf.size))
assert rc == 0
st = buf.reshape(512, 40).astype(np.int64)
os.makedirs('gpurun_out', exist_ok=True)
np.save(os.path.join('gpurun_out', f'stamps_{shape}.npy'), st)
def q(a): a = np.asarray(a, dtype=np.float64); return f"min {a.min():.0f} med {np.median(a):.0f} p90 {np.percentile(a, 90):.0f} max {a.max():.0f}"
for name, base, nw in (("point_pass_tab", 0, int(os.environ.get("TABW", "19"))), ("camera_pass", 256, 16)):
    rows = st[base:base + 148]
    rows = rows[rows[:, 0] > 0]
    if not len(rows):
        continue
    print(f"== {name}: {len(rows)} CTAs")
    print("  CTA entry spread (globaltimer ns):", q(rows[:, 0] - rows[:, 0].min()))
    if base == 0:
        print("  table ready after entry (cycles):", q(rows[:, 3] - rows[:, 2]), "| (ns, globaltimer):", q(rows[:, 1] - rows[:, 0]))
    ends = rows[:, 8:8 + nw] - rows[:, 2:3]
    ends = np.where(rows[:, 8:8 + nw] > 0, ends, np.nan)
    print("  warp end after CTA entry (cycles):", q(ends[~np.isnan(ends)]))
    cta_end = np.nanmax(ends, axis=1)
    print("  CTA end (max over warps, cycles):", q(cta_end), "| warp spread inside a CTA (max-min):", q(np.nanmax(ends, axis=1) - np.nanmin(ends, axis=1)))
    # per-warp-slot finish time (mean over CTAs): which of a CTA's warps (= which band of track lengths) runs longest
    print("  warp slot mean end (cycles):", " ".join(f"{x:.0f}" for x in np.nanmean(ends, axis=0)))
