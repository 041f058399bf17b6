"""Developer check: SHA-256 of the outputs of the build stage and of whole solves on seeded inputs.
Run once per library build (OCTANE_B200_LIB=...) and diff the printed lines: a kernel revision that
claims bit-identical results must print the same hashes."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import octane_b200 as ob  # noqa: E402
from octane_b200 import synthetic as S  # noqa: E402


def h(*arrs):
    m = hashlib.sha256()
    for a in arrs:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()[:16]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    ctx = ob.Context(0)
    p = ob.default_params()
    for nc in (1, 2):
        nx, ny = 333, 257
        c1, c2 = [], []
        for ch in range(nc):
            a, b, ut, vt = S.make_pair(nx, ny, 30 + ch)
            c1.append(a); c2.append(b)
        g1 = np.stack(c1); g2 = np.stack(c2)
        u = (0.7 * ut).astype(np.float32); v = (0.7 * vt).astype(np.float32)
        u[5:40, 5:40] = 0.0; v[5:40, 5:40] = 0.0           # flat patch: psi at its 1/sqrt(1e-6) ceiling
        u[:, -3:] += 9.0; v[:3, :] -= 9.0                  # warps that leave the image (clamp + derivative zeroing)
        uh = (0.5 * ut).astype(np.float32); vh = (0.5 * vt).astype(np.float32)
        for gnc in (0, 1, 2):
            for lc in (0.0, 0.25):
                coef = torch.zeros((7, ny, nx), device="cuda"); bu = torch.zeros((ny, nx), device="cuda"); bv = torch.zeros_like(bu)
                ctx.stage_build(dev(u), dev(v), dev(uh) if lc else None, dev(vh) if lc else None, dev(g1), dev(g2), nx, ny, nc, p,
                                lc, gnc, coef, bu, bv)
                print(f"build nc={nc} gnc={gnc} lambdac={lc}: {h(coef.cpu().numpy(), bu.cpu().numpy(), bv.cpu().numpy())}")
    for (nx, ny, seed) in ((500, 500, 1), (1500, 1100, 2), (2000, 2000, 3)):
        a, b = S.make_pair_torch(nx, ny, seed, "cuda")
        u = torch.zeros((ny, nx), device="cuda"); v = torch.zeros_like(u)
        ctx.oct_variational_optical_flow(a, b, u, v, p)
        ctx.synchronize()
        st = ctx.stats()
        print(f"flow {nx}x{ny}: {h(u.cpu().numpy(), v.cpu().numpy())} its={sum(st.cg_iterations[:st.n_solves])}")
    ctx.close()


if __name__ == "__main__":
    main()
