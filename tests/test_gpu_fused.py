"""The merged-reduction PCG kernel (octane_b200/csrc/pcg_fused.cu) against the CPU oracle.

Stage level: the kernel on systems the oracle built, for shapes that exercise its tiling (several strips, a
narrow last strip, odd widths, one and several row segments), after 1, 2, 3, 4, 7 and 30 iterations -- against
oracle_pcg (the reference's recurrence, the yardstick) and oracle_pcg_merged (the model with the kernel's own
operation order).  Whole path: solver 1 (default) against solver 0 (the two-pass kernels that follow the
reference's loop literally) and against the reference's fixtures."""
import numpy as np
import pytest

import cases
import octane_b200 as ob
from conftest import load_golden, parity_report
from octane_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_system(oracle, nx, ny, gnc, seed):
    L = oracle.lib()
    a, b, ut, vt = S.make_pair(nx, ny, seed)
    g1 = a[None].copy(); g2 = b[None].copy()
    u = (0.7 * ut).astype(np.float32); v = (0.7 * vt).astype(np.float32)
    f = {k: np.zeros((1, ny, nx), np.float32) for k in ("g1x", "g1y", "g2x", "g2y", "g2xx", "g2xy", "g2yy")}
    L.oracle_gradient(g1, f["g1x"], f["g1y"], nx, ny, 1)
    L.oracle_gradient(g2, f["g2x"], f["g2y"], nx, ny, 1)
    L.oracle_gradient(f["g2x"], f["g2xx"], f["g2xy"], nx, ny, 1)
    L.oracle_gradient(f["g2y"], f["g2xy"], f["g2yy"], nx, ny, 1)
    coef = np.zeros((7, ny, nx), np.float32); bu = np.zeros((ny, nx), np.float32); bv = np.zeros((ny, nx), np.float32)
    L.oracle_build(u, v, None, None, g1, f["g1x"], f["g1y"], g2, f["g2x"], f["g2y"], f["g2xx"], f["g2xy"], f["g2yy"],
                   nx, ny, 1, 5.0, 0.2, 0.0, gnc, 1, coef, bu, bv)
    return coef, bu, bv


@pytest.mark.parametrize("shape,gnc", [((1100, 200), 2), ((1021, 96), 1), ((2043, 130), 2), ((777, 64), 0), ((512, 300), 2),
                                       ((3061, 70), 1)])
def test_fused_kernel_matches_both_oracle_recurrences(ctx, oracle, shape, gnc):
    import torch
    nx, ny = shape
    coef, bu, bv = oracle_system(oracle, nx, ny, gnc, 50 + gnc)
    L = oracle.lib()
    n = nx * ny
    for iters in (1, 2, 3, 4, 7, 30):
        want = {}
        for key, fn in (("reference", L.oracle_pcg), ("merged", L.oracle_pcg_merged)):
            b1, b2 = bu.copy(), bv.copy()
            xu, xv = np.zeros((ny, nx), np.float32), np.zeros((ny, nx), np.float32)
            its = fn(coef, b1, b2, xu, xv, nx, ny, iters, 1e-8, np.zeros(8 * n, np.float32))
            want[key] = (its, xu, xv)
        dxu, dxv = torch.zeros((ny, nx), device="cuda"), torch.zeros((ny, nx), device="cuda")
        got_its = ctx.stage_pcg(dev(coef), dev(bu), dev(bv), nx, ny, iters, 1e-8, dxu, dxv)
        assert ctx.stats().pcg_solver == 1                 # the merged-reduction kernel ran
        gu, gv = dxu.cpu().numpy(), dxv.cpu().numpy()
        assert np.isfinite(gu).all() and np.isfinite(gv).all()
        scale = max(1.0, float(np.abs(want["reference"][1]).max()))
        for key, tol in (("merged", 2e-5), ("reference", 5e-5)):
            its, xu, xv = want[key]
            assert got_its == its
            err = max(float(np.abs(gu - xu).max()), float(np.abs(gv - xv).max()))
            parity_report("fused_stage_pcg", shape=f"{nx}x{ny}", gnc=gnc, iters=iters, against=key, max_abs=err)
            assert err < tol * scale, (key, iters, err)


def test_fused_stop_rule_and_zero_rhs(ctx):
    import torch
    nx, ny = 640, 96
    coef = np.zeros((7, ny, nx), np.float32)
    coef[0] = 9; coef[2] = 9
    for k in (3, 4, 5, 6):
        coef[k] = -1
    coef[3][:, 0] = 0; coef[5][:, -1] = 0; coef[4][0, :] = 0; coef[6][-1, :] = 0
    coef[3][:, -1] = -2; coef[5][:, 0] = -2; coef[4][-1, :] = -2; coef[6][0, :] = -2
    z = np.zeros((ny, nx), np.float32)
    dxu, dxv = torch.ones((ny, nx), device="cuda"), torch.ones((ny, nx), device="cuda")
    assert ctx.stage_pcg(dev(coef), dev(z), dev(z), nx, ny, 30, 1e-8, dxu, dxv) == 0     # ||b||^2 <= tol: no iteration
    assert float(dxu.abs().max()) == 0 and float(dxv.abs().max()) == 0
    b = np.random.default_rng(1).standard_normal((ny, nx)).astype(np.float32)
    its = ctx.stage_pcg(dev(coef), dev(b), dev(b), nx, ny, 200, 1e-8, dxu, dxv)
    assert 0 < its < 200                                                               # converges before the cap
    # and solves the system: A x = b to the stop tolerance
    x = dxu.cpu().numpy().astype(np.float64)
    pad = np.pad(x, 1, mode="reflect")
    ax = 9 * x - pad[1:-1, :-2] - pad[1:-1, 2:] - pad[:-2, 1:-1] - pad[2:, 1:-1]
    assert np.square(ax - b).sum() * 2 < 1e-6


@pytest.mark.parametrize("name", ["ref_1024x768", "ref_fd_limb", "ref_meso_2000"])
def test_both_solvers_meet_the_reference_fixture(ctx, name):
    """solver 0 = the reference's loop literally (two launches per iteration), solver 1 = merged reduction"""
    g = load_golden(name)
    img1, img2 = cases.headline_inputs(cases.HEADLINE[name])
    ny, nx = img1.shape
    res = {}
    try:
        for solver in (0, 1):
            ctx.set_solver(solver)
            u, v = np.zeros((ny, nx), np.float32), np.zeros((ny, nx), np.float32)
            ctx.oct_variational_optical_flow(img1, img2, u, v, ob.default_params())
            st = ctx.stats()
            assert st.pcg_solver == solver
            res[solver] = (u, v, list(st.cg_iterations[:st.n_solves]), int(st.kernel_launches))
            d = cases.headline_digest(u, v, int(g["stride"]))
            mean = max(float(np.abs(d[k] - g[k]).mean()) for k in ("us", "vs"))
            mx = max(float(np.abs(d[k] - g[k]).max()) for k in ("us", "vs", "ub", "vb"))
            parity_report("solver_vs_reference_fixture", case=name, solver=solver, mean_abs=mean, max_abs=mx,
                          launches=int(st.kernel_launches))
            assert mean <= 1e-3 and mx <= 1e-2
    finally:
        ctx.set_solver(1)
    assert res[0][2] == res[1][2]                                     # same iteration counts
    assert res[1][3] < res[0][3]                                      # and fewer launches
    du = float(np.abs(res[0][0] - res[1][0]).max()); dv = float(np.abs(res[0][1] - res[1][1]).max())
    parity_report("solver0_vs_solver1", case=name, max_abs=max(du, dv))
    assert max(du, dv) < 5e-4


def test_fused_graph_and_plain_launch_paths_agree(ctx):
    a, b = S.make_pair(900, 260, 41)[:2]
    p = ob.default_params(kiters=2)
    outs = []
    for graphs, profile in ((True, False), (False, False), (True, True)):
        ctx.set_graphs(graphs); ctx.set_profile(profile)
        u, v = np.zeros((260, 900), np.float32), np.zeros((260, 900), np.float32)
        ctx.oct_variational_optical_flow(a, b, u, v, p)
        st = ctx.stats()
        outs.append((u, v))
    ctx.set_graphs(True); ctx.set_profile(False)
    assert st.pcg_solver == 1 and st.finest_pass1_ms > 0 and st.finest_pass2_ms == 0 and 45 < st.finest_pass1_bytes_per_px < 75
    for u, v in outs[1:]:
        assert np.array_equal(u, outs[0][0]) and np.array_equal(v, outs[0][1])
