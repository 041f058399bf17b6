"""The merged form of the PCG recurrence (what octane_b200/csrc/pcg_fused.cu computes), measured on the CPU.

oracle_pcg_merged is a model of the PRODUCT's kernel, not a restatement of the reference; the yardstick stays the
reference: the fixtures its own sm_100 build produced (tests/golden/var_*.npz) and oracle_pcg, which follows its
recurrence literally.  Two facts are pinned here:
  * with the symmetry-free expansion of p.Ap the merged form lands as close to the reference's outputs as the
    literal recurrence does (the reference's own run-to-run noise), with equal iteration counts;
  * the system is NOT symmetric at the mirrored edges, which is why the textbook single-reduction shortcut
    (p.Ap = z.w - beta (r.z) / alpha_prev) is not used."""
import ctypes as C

import numpy as np
import pytest

import cases
from conftest import load_golden


@pytest.mark.parametrize("name", sorted(cases.VARIATIONAL))
def test_merged_recurrence_matches_reference_fixture(oracle, name):
    c = cases.VARIATIONAL[name]
    g = load_golden(name)
    img1, img2, u0, v0 = cases.variational_inputs(c)
    p = oracle.params(**c.get("params", {}))
    L = oracle.lib()
    u, v, its = oracle.variational_flow(img1, img2, p, u0, v0, nc=c.get("nc", 1))
    L.oracle_set_solver(1)
    try:
        um, vm, itsm = oracle.variational_flow(img1, img2, p, u0, v0, nc=c.get("nc", 1))
    finally:
        L.oracle_set_solver(0)
    assert list(its) == list(itsm)
    for a, b in ((um, g["u"]), (vm, g["v"])):
        d = np.abs(a - b)
        assert d.mean() < 1e-3 and d.max() < 1e-2                     # the north-star gates
        assert d.max() < max(20 * float(g["spread"].max()), 2e-4)     # and the reference's own noise level
    assert np.abs(um - u).max() < 2e-4 and np.abs(vm - v).max() < 2e-4


def test_edge_merged_system_is_not_symmetric(oracle):
    """<A x, y> != <x, A y> for the boundary-merged operator: the first column couples to the second with
    2 W, the second to the first with W (reference :929-1077)."""
    nx, ny = 24, 16
    rng = np.random.default_rng(3)
    coef = np.zeros((7, ny, nx), np.float32)
    coef[0] = 9; coef[2] = 9
    for k in (3, 4, 5, 6):
        coef[k] = -1
    coef[3][:, 0] = 0; coef[5][:, -1] = 0; coef[4][0, :] = 0; coef[6][-1, :] = 0
    coef[3][:, -1] = -2; coef[5][:, 0] = -2; coef[4][-1, :] = -2; coef[6][0, :] = -2
    x = rng.standard_normal((2, ny, nx)).astype(np.float32)
    y = rng.standard_normal((2, ny, nx)).astype(np.float32)
    ax = np.zeros_like(x); ay = np.zeros_like(y)
    L = oracle.lib()
    L.oracle_apply(coef, x[0], x[1], nx, ny, ax[0], ax[1])
    L.oracle_apply(coef, y[0], y[1], nx, ny, ay[0], ay[1])
    lhs, rhs = float((ax.astype(np.float64) * y).sum()), float((x.astype(np.float64) * ay).sum())
    assert abs(lhs - rhs) > 1e-3 * abs(lhs)
    # interior-only vectors see a symmetric operator
    x[:, 0, :] = x[:, -1, :] = 0; x[:, :, 0] = x[:, :, -1] = 0; x[:, 1, :] = x[:, -2, :] = 0; x[:, :, 1] = x[:, :, -2] = 0
    y[:, 0, :] = y[:, -1, :] = 0; y[:, :, 0] = y[:, :, -1] = 0; y[:, 1, :] = y[:, -2, :] = 0; y[:, :, 1] = y[:, :, -2] = 0
    L.oracle_apply(coef, x[0], x[1], nx, ny, ax[0], ax[1])
    L.oracle_apply(coef, y[0], y[1], nx, ny, ay[0], ay[1])
    lhs, rhs = float((ax.astype(np.float64) * y).sum()), float((x.astype(np.float64) * ay).sum())
    assert abs(lhs - rhs) < 1e-5 * max(abs(lhs), 1.0)
