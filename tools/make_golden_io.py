"""Generate tests/golden/io_ref_tiny.{txt,nvm,out,npz}: model files WRITTEN by the reference's own SaveBundlerModel /
SaveNVM / SaveBundlerOut (Thirdparty/pba/pba/pba_util.h via oracle/io_ref_harness.cpp) and what its own loaders read back.
Run in the build container only (needs oracle/_ref/libio_ref.so); the outputs are committed.
    python tools/make_golden_io.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refs
from sequencesfm_b200 import io as bio
from sequencesfm_b200.synth import make_problem


def main():
    out = os.path.join(ROOT, "tests", "golden")
    p = make_problem("tiny", 7)
    rng = np.random.default_rng(7)
    arrays = {}
    for kind, ext in ((0, "txt"), (1, "nvm"), (2, "out")):
        m = bio.from_problem(p)
        m.focal = rng.uniform(500, 700, m.N); m.radial = rng.uniform(-1e-2, 1e-2, m.N)
        m.pt_rgb = rng.integers(0, 256, (m.M, 3)).astype(np.int32)
        if kind == 1:
            m.distortion_type[:] = -1
        path = os.path.join(out, f"io_ref_tiny.{ext}")
        refs.io_save(path, kind, m)
        for junk in (path.replace(".out", "-list.txt"),):
            if kind == 2 and os.path.exists(junk):
                os.remove(junk)
        back = refs.io_load(path)
        for k, v in back.items():
            arrays[f"{ext}_{k}"] = v
        print(ext, os.path.getsize(path), "bytes")
    np.savez_compressed(os.path.join(out, "io_ref_tiny.npz"), **arrays)


if __name__ == "__main__":
    main()
