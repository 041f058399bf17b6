"""Times the two stages outside the benchmarked step on device buffers (CUDA events on the context's stream):
oct_zoom_out_float (a 0.5 km mesoscale channel, 2000 x 2000, down to 1 km and 2 km) and the -srsal smoother on a
2000 x 2000 flow.  Also dumps our srsal outputs for the golden cases so they can be compared offline with
the fixtures the reference produced in the same call.  Usage: python scripts/time_post.py [outdir]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402

import cases  # noqa: E402
import octane_b200 as ob  # noqa: E402


def timed(ctx, fn, reps=5):
    st = ctx._ext_stream()
    fn(); ctx.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        a.record(st)
        for _ in range(reps):
            fn()
        b.record(st)
    ctx.synchronize()
    return a.elapsed_time(b) / reps


def main(out):
    os.makedirs(out, exist_ok=True)
    ctx = ob.Context(0)
    n = 2000
    g = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn((n, n), device="cuda", generator=g) * 50 + 100
    lines = []
    for factor in (0.5, 0.25):
        ms = timed(ctx, lambda: ctx.oct_zoom_out_float(f, factor))
        lines.append(f"zoom_out_float {n}x{n} factor {factor}: {ms:.3f} ms ({n * n / ms / 1e3:.0f} Mpix/s of input)")
    u = torch.randn((n, n), device="cuda", generator=g)
    v = torch.randn((n, n), device="cuda", generator=g)
    cth = 6000 + 40 * torch.randn((n, n), device="cuda", generator=g)
    ms = timed(ctx, lambda: ctx.oct_srsal_cu(u, v, cth), reps=3)
    lines.append(f"srsal {n}x{n}: {ms:.2f} ms ({n * n / ms / 1e3:.1f} Mpix/s; {n * n * 1369 / ms / 1e6:.1f} G taps/s)")
    for name, c in cases.SRSAL.items():
        uu, vv, cc = cases.srsal_inputs(c)
        su, sv = ctx.oct_srsal_cu(uu.copy(), vv.copy(), cc)
        np.savez_compressed(os.path.join(out, "ours_" + name + ".npz"), u=su, v=sv)
    print("\n".join(lines))
    open(os.path.join(out, "time_post.txt"), "w").write("\n".join(lines) + "\n")
    ctx.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "post"))
