"""Quick GPU parity probe (developer tool): ours vs CPU oracle vs the reference CUDA build."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import octane_b200 as ob
from octane_b200 import synthetic as S
from oracle import oracle as O


def stat(name, a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"  {name}: mean {d.mean():.3e} max {d.max():.3e}", flush=True)


def main():
    ctx = ob.Context(0)
    dev = torch.device("cuda:0")
    sizes = [(96, 80), (200, 160), (500, 500)]
    if len(sys.argv) > 1:
        sizes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    for nx, ny in sizes:
        print(f"== {nx}x{ny}", flush=True)
        i1, i2, ut, vt = S.make_pair(nx, ny, seed=nx + ny)
        # stages
        for f in (0.5, 0.25, 0.125):
            exp_nx, exp_ny = int(nx * f + 0.5), int(ny * f + 0.5)
            if exp_nx < 4 or exp_ny < 4:
                continue
            o = np.zeros((exp_ny, exp_nx), np.float32)
            O.lib().oracle_blur_decimate(i1, nx, ny, 1, f, o)
            d_in = torch.from_numpy(i1).to(dev); d_out = torch.zeros((exp_ny, exp_nx), device=dev)
            ctx.stage_blur_decimate(d_in, nx, ny, 1, f, d_out)
            stat(f"blur_decimate f={f}", d_out.cpu().numpy(), o)
        gx = np.zeros_like(i1); gy = np.zeros_like(i1)
        O.lib().oracle_gradient(i1, gx, gy, nx, ny, 1)
        dgx = torch.zeros((ny, nx), device=dev); dgy = torch.zeros((ny, nx), device=dev)
        ctx.stage_gradient(torch.from_numpy(i1).to(dev), nx, ny, 1, dgx, dgy)
        stat("gradx", dgx.cpu().numpy(), gx); stat("grady", dgy.cpu().numpy(), gy)
        nxx, nyy = 2 * nx + 1, 2 * ny - 1
        zo = np.zeros((nyy, nxx), np.float32)
        O.lib().oracle_zoom_in(ut, zo, nx, ny, nxx, nyy, 0.5)
        dz = torch.zeros((nyy, nxx), device=dev)
        ctx.stage_zoom_in(torch.from_numpy(ut).to(dev), nx, ny, nxx, nyy, 0.5, dz)
        stat("zoom_in", dz.cpu().numpy(), zo)
        # full solve
        p = ob.default_params()
        t = time.time(); uo, vo, its = O.variational_flow(i1, i2); t_or = time.time() - t
        u = np.zeros((ny, nx), np.float32); v = np.zeros((ny, nx), np.float32)
        ctx.oct_variational_optical_flow(i1, i2, u, v, p)
        t = time.time(); ctx.oct_variational_optical_flow(i1, i2, u.copy() * 0, v.copy() * 0, p); t_me = time.time() - t
        st = ctx.stats()
        print("  its oracle", list(its), "\n  its ours  ", list(st.cg_iterations[:st.n_solves]), flush=True)
        stat("u ours-oracle", u, uo); stat("v ours-oracle", v, vo)
        t = time.time(); ur, vr = O.ref_variational(i1, i2); t_ref = time.time() - t
        stat("u ours-ref", u, ur); stat("v ours-ref", v, vr)
        stat("u oracle-ref", uo, ur); stat("v oracle-ref", vo, vr)
        ur2, vr2 = O.ref_variational(i1, i2)
        stat("u ref-ref(rerun)", ur, ur2)
        m = np.s_[16:-16, 16:-16]
        stat("u ours-truth(interior)", u[m], ut[m])
        print(f"  time: oracle {t_or:.3f}s  ours {t_me*1e3:.2f} ms  ref {t_ref:.3f}s  launches {st.kernel_launches}", flush=True)
        # navigation
        xs, ys, xo, yo, dt = S.SECTORS["meso_0.5km"]
        nav = ob.goes_nav(xs, ys, xo, yo)
        onav = O.goes_nav(xs, ys, xo, yo)
        outs = [np.zeros((ny, nx), np.int16) for _ in range(4)]
        ctx.oct_pix2uv_cuda(nav, 0.0, dt, u, v, *outs, p)
        oo = O.pix2uv(onav, 0.0, dt, u, v)
        rr = O.ref_pix2uv(onav, 0.0, dt, u, v)
        for k, nm in enumerate(("U", "V", "U_raw", "V_raw")):
            print(f"  nav {nm}: ours-oracle max {np.abs(outs[k].astype(int)-oo[k]).max()}  ours-ref max {np.abs(outs[k].astype(int)-rr[k]).max()}"
                  f"  n!=ref {(outs[k]!=rr[k]).sum()}  range [{outs[k].min()},{outs[k].max()}]", flush=True)


if __name__ == "__main__":
    main()
