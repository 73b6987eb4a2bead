"""CPU: the seeded scene generators behind the post-BA stage and triangulation tests / benches."""
import numpy as np

from sequencesfm_b200.synth import make_ground_scene, make_two_view


def test_ground_scene_is_a_noisy_plane_with_outliers_and_cameras_above_it():
    pts, cams = make_ground_scene(3, 20000, outlier_frac=0.1, noise=0.3)
    assert pts.shape == (20000, 3) and cams.shape == (8, 12) and pts.flags.c_contiguous
    c = pts.mean(0)
    _, s, vt = np.linalg.svd(pts - c, full_matrices=False)
    n = vt[2]                                            # least-variance direction = plane normal
    d = (pts - np.median(pts, 0)) @ n
    d -= np.median(d)
    assert 0.15 < np.median(np.abs(d)) < 0.35            # ~0.67 sigma of the height noise
    assert 0.05 < (np.abs(d) > 1.5).mean() < 0.15        # the outliers stand well clear of it
    for cam in cams:
        R, t = cam.reshape(3, 4)[:, :3], cam.reshape(3, 4)[:, 3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
        centre = -R.T @ t
        assert abs((centre - np.median(pts, 0)) @ n) > 20  # cameras fly 40-60 units above the plane
        z = (pts @ R.T + t)[:, 2]
        assert (z > 0).mean() > 0.99                     # and look at it
    again, _ = make_ground_scene(3, 20000, outlier_frac=0.1, noise=0.3)
    assert np.array_equal(pts, again)


def test_ground_scene_on_plane_fraction_gives_exact_ties():
    pts, _ = make_ground_scene(5, 3000, tilt=0.0, on_plane_frac=0.7)
    assert (pts[:, 2] == 0.0).sum() == 2100


def test_two_view_points_are_seen_by_both_cameras_with_the_injected_noise():
    pt1, pt2, K, P, P1, X = make_two_view(4, 5000, noise_px=0.5)
    assert pt1.dtype == np.float32 and pt2.dtype == np.float32
    for Pm, px in ((P, pt1), (P1, pt2)):
        xc = X @ Pm[:, :3].T + Pm[:, 3]
        assert (xc[:, 2] > 0).all()
        proj = (xc[:, :2] / xc[:, 2:3]) * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
        e = np.linalg.norm(proj - px, axis=1)
        assert 0.5 < e.mean() < 0.75                     # Rayleigh mean of sigma = 0.5 px is 0.63
    R = P1[:, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
