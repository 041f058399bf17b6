"""Generates the headline fixtures tests/golden/ref_*.npz: the UNMODIFIED REFERENCE (oracle/_ref/libref_cuda.so)
on the benchmarked configurations -- the 2000^2 mesoscale sector, CONUS 10000 x 6000, and three 2048^2 crops of
the tapered full-disk scene (cases.HEADLINE).  Needs a GPU and ~25 GB of device memory for CONUS:

    gpurun -- 'python tests/golden/make_golden_headline.py gpurun_out/golden_headline'
    cp gpurun_out/golden_headline/*.npz tests/golden/

Each case runs in a child process (run_ref_cuda.py).  A fixture keeps a strided sample of u, v, one
full-resolution block, float64 statistics, the reference's run-to-run spread and a hash of the inputs."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import cases  # noqa: E402


def run_reference(img1, img2, runs=1, timeout=1500):
    """(u, v, spread, seconds) of the reference's CUDA solver, in a child process"""
    with tempfile.TemporaryDirectory() as tmp:
        src, dst = os.path.join(tmp, "in.npz"), os.path.join(tmp, "out.npz")
        np.savez(src, img1=img1, img2=img2)
        subprocess.run([sys.executable, os.path.join(HERE, "run_ref_cuda.py"), src, dst, str(runs)], check=True, timeout=timeout)
        d = np.load(dst)
        return d["u"], d["v"], float(d["spread"]), float(d["seconds"])


def main(out, only=None):
    os.makedirs(out, exist_ok=True)
    for name, c in cases.HEADLINE.items():
        if only and name not in only:
            continue
        img1, img2 = cases.headline_inputs(c)
        try:
            u, v, spread, sec = run_reference(img1, img2, c.get("runs", 1))
        except Exception as e:       # keep going: one case failing (memory, time limit) must not lose the others
            print(name, "FAILED", repr(e)[:200], flush=True)
            continue
        d = cases.headline_digest(u, v, c["stride"])
        np.savez_compressed(os.path.join(out, name + ".npz"), stride=np.int32(c["stride"]), spread=np.float64(spread),
                            seconds=np.float64(sec), inputs_sha1=np.array(cases.input_hash(img1, img2)), **d)
        ny, nx = u.shape
        print(name, f"{nx}x{ny}", "ref seconds", round(sec, 2), "Mpix/s", round(nx * ny / sec / 1e6, 3), "spread", spread,
              "mean|u|", d["stats"][0], "max|u|", d["stats"][2], "finite", bool(np.isfinite(u).all() and np.isfinite(v).all()),
              flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0] if args else os.path.join(ROOT, "gpurun_out", "golden_headline"), set(args[1:]) or None)
