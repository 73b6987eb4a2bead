"""CPU, world_size 2 over gloo: the point partition used by the multi-GPU path.

Each rank owns a contiguous range of points and their observations plus a replica of the cameras; the only
exchange is a sum over ranks of camera-sized buffers (U | g_c | cost after linearisation, W v per PCG step).  The
per-rank numbers are produced here by the CPU oracle on the rank's shard (with the GLOBAL normalisation, exactly as
the engine does) and all-reduced with gloo; they must equal the single-rank linearisation and Schur product.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sequencesfm_b200.synth import Problem, make_problem, partition_points


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out):
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem("small", 3, outlier_frac=0.02)
    full = oracle.linearize(p)                       # also yields the globally normalised state
    j0, j1 = partition_points(p.obs_pt, p.M, world)[rank]
    k0, k1 = np.searchsorted(p.obs_pt, [j0, j1])
    # the shard in GLOBAL normalised coordinates: normalisation is a global reduction done before sharding
    shard = Problem(np.array([1.0, p.K5[1] / p.K5[0], p.K5[2] / p.K5[0], p.K5[3] / p.K5[0], p.K5[4] / p.K5[0]]),
                    full["RTn"].copy(), full["Xn"][j0:j1].copy(),
                    (np.rint(p.obs_xy[k0:k1].astype(np.float32)) / p.K5[0]).astype(np.float64),
                    p.obs_cam[k0:k1].copy(), (p.obs_pt[k0:k1] - j0).astype(np.int32))
    x = np.random.default_rng(0).normal(size=6 * p.N)
    lam = full["lambda0"]
    loc = oracle.linearize(shard, x=x, lam=lam, normalize=0, round_pixels=0, inlier_px=2.0 / p.K5[0])
    # camera-sized partial sums; S x = (U + lam I) x - W V^-1 W^t x : the W-part is what is all-reduced
    Ux = np.einsum("nab,nb->na", loc["U"], x.reshape(-1, 6)) + lam * x.reshape(-1, 6)
    wv_local = Ux - loc["Sx"]
    buf = torch.from_numpy(np.concatenate([loc["U"].reshape(-1), loc["gc"].reshape(-1), [loc["cost"]], wv_local.reshape(-1)]))
    dist.all_reduce(buf)
    if rank == 0:
        n = p.N
        U = buf[: 36 * n].numpy().reshape(n, 6, 6); gc = buf[36 * n: 42 * n].numpy().reshape(n, 6)
        cost = float(buf[42 * n]); wv = buf[42 * n + 1:].numpy().reshape(n, 6)
        Sx = np.einsum("nab,nb->na", U, x.reshape(-1, 6)) + lam * x.reshape(-1, 6) - wv
        ref = oracle.linearize(p, x=x, lam=lam)
        out["U"] = float(np.abs(U - ref["U"]).max() / np.abs(ref["U"]).max())
        out["gc"] = float(np.abs(gc - ref["gc"]).max() / np.abs(ref["gc"]).max())
        out["cost"] = abs(cost - ref["cost"]) / ref["cost"]
        out["Sx"] = float(np.abs(Sx - ref["Sx"]).max() / np.abs(ref["Sx"]).max())
    dist.barrier()
    dist.destroy_process_group()


def test_point_partition_allreduce_equals_single_rank():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        out = dict(out)
    assert out["U"] < 1e-12 and out["gc"] < 1e-11 and out["cost"] < 1e-12 and out["Sx"] < 1e-11, out
