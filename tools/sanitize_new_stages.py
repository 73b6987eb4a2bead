"""Small, ragged-size run of every kernel of the post-BA stage and the triangulation (written as the driver of a
`compute-sanitizer --tool memcheck` pass; the sanitizer is closed on this GPU pool, so it only runs plainly here)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencesfm_b200 import binding, post, tri
from sequencesfm_b200.synth import make_ground_scene, make_problem, make_two_view

for M, H, cap in ((4097, 101, 0), (1000, 7, -1), (3000, 33, 2)):
    pts, cams = make_ground_scene(1, M)
    rng = np.random.default_rng(M)
    t = np.stack([rng.permutation(M)[:3] for _ in range(H)]).astype(np.int32)
    with post.PostStage(device=0, select_cap=cap) as st:
        fit = st.plane_fit(pts, t)
        st.transform(np.eye(3, 4), pts, cams)
        fit2, p2, c2, axes = st.ground_stage(pts, cams, t)
        print("post", M, H, cap, fit.n_kept, fit2.select_fallback, len(p2))
p = make_problem("small", 1)
with binding.Engine(device=0, max_iterations=3) as eng, post.PostStage(device=0) as st:
    eng.set_problem_from(p); eng.solve()
    t = np.stack([np.random.default_rng(i).permutation(p.M)[:3] for i in range(20)]).astype(np.int32)
    fit, pp, cc, ax = st.ground_stage_ba(eng, t)
    print("chained", fit.n_kept)
with tri.Triangulator(device=0) as e:
    for n in (1, 129, 5000):
        a, b, K, P, P1, _ = make_two_view(2, n, dist=[-0.1, 0.02, 1e-3, 1e-3, 0.01])
        r = e.triangulate(a, b, K, P, P1, dist=[-0.1, 0.02, 1e-3, 1e-3, 0.01])
        g = e.gate()
        print("tri", n, r.mean_reproj_error, g.status, g.n_kept)
print("done")
