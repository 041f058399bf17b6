"""Child process that runs the UNMODIFIED REFERENCE's CUDA solver (oracle/_ref/libref_cuda.so, the reference
sources recompiled for sm_100) on one image pair: `python run_ref_cuda.py in.npz out.npz [runs]`.
in.npz holds img1, img2 (ny x nx float32); out.npz receives u, v of the first run, the spread over `runs`
runs (the reference's float atomics make it run-to-run non-reproducible) and the seconds of the best run.
A process of its own because the reference exit()s on errors and never frees its small allocations."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402


def main(src, dst, runs=1):
    d = np.load(src)
    img1, img2 = d["img1"], d["img2"]
    rp = O.ref_params()
    outs, secs = [], []
    for _ in range(runs):
        t = time.perf_counter()
        outs.append(O.ref_variational(img1, img2, rp))
        secs.append(time.perf_counter() - t)
    u, v = outs[0]
    spread = max([0.0] + [max(float(np.abs(o[0] - u).max()), float(np.abs(o[1] - v).max())) for o in outs[1:]])
    np.savez(dst, u=u, v=v, spread=np.float64(spread), seconds=np.float64(min(secs)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
