"""Multi-GPU worker (launched by torch.distributed.run, one rank per GPU): solves one
pair as row bands and compares with the single-GPU solve of the same pair on rank 0.
Prints one line `BAND_RESULT {...json...}` on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import octane_b200 as ob  # noqa: E402
from octane_b200 import synthetic as S  # noqa: E402


def main():
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    max_disp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    with_fg = len(sys.argv) > 4 and sys.argv[4] == "fg"        # first guess + hinting (-firstguess -lambdac 0.5)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p = ob.default_params(max_disp=max_disp, first_guess=int(with_fg), lambdac=0.5 if with_fg else 0.0)
    ctx = ob.Context(local)

    def first_guess(r0, r1):
        """smooth displacement field close to the true drift, rows [r0, r1)"""
        y = torch.arange(r0, r1, device=dev, dtype=torch.float32)[:, None]
        x = torch.arange(nx, device=dev, dtype=torch.float32)[None, :]
        return ((0.6 + 0.3 * torch.sin(x / 23.0) * torch.cos(y / 31.0)).contiguous(),
                (-0.3 + 0.2 * torch.cos(x / 19.0 + y / 29.0)).contiguous())

    ids = [ob.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    own0, own1, in0, in1 = ob.band_plan(nx, ny, p, rank, world)
    img1, img2 = S.make_pair_torch(nx, ny, 9, dev, rows=(in0, in1))
    u = torch.zeros((own1 - own0, nx), device=dev); v = torch.zeros_like(u)
    fgu, fgv = first_guess(in0, in1) if with_fg else (None, None)
    ctx.oct_variational_optical_flow_band(img1, img2, u, v, nx, ny, p, fg_u_band=fgu, fg_v_band=fgv)
    ctx.synchronize()
    st = ctx.stats()
    # second run: bit-reproducible for a fixed world size
    u2 = torch.zeros_like(u); v2 = torch.zeros_like(v)
    ctx.oct_variational_optical_flow_band(img1, img2, u2, v2, nx, ny, p, fg_u_band=fgu, fg_v_band=fgv)
    ctx.synchronize()
    repro = bool(torch.equal(u, u2) and torch.equal(v, v2))
    # navigation of the band
    xs, ys, xo, yo, dt = S.SECTORS["conus_0.5km"]
    nav = ob.goes_nav(xs, ys, xo, yo)
    sh = [torch.zeros((own1 - own0, nx), dtype=torch.int16, device=dev) for _ in range(4)]
    ctx.oct_pix2uv_band(nav, 0.0, dt, u, v, nx, own0, own1 - own0, *sh, p)
    ctx.synchronize()
    # the pipelined host-buffer dispatcher on the same band: two pairs in flight, outputs equal to the device path's
    stream_equal = True
    if not with_fg:
        h1 = torch.empty(img1.shape, pin_memory=True).copy_(img1).numpy()
        h2 = torch.empty(img2.shape, pin_memory=True).copy_(img2).numpy()
        outs = [{k: torch.zeros((own1 - own0, nx), dtype=(torch.float32 if k.endswith("Pix") else torch.int16), pin_memory=True).numpy()
                 for k in ("uPix", "vPix", "uVal", "vVal", "uVal2", "vVal2")} for _ in range(2)]
        for k in range(3):
            ctx.stream_submit(k % 2, h1, h2, nav, 0.0, dt, p, outs[k % 2], nx, ny)
            if k > 0:
                ctx.stream_wait((k - 1) % 2)
        ctx.stream_wait(0)
        for o in outs:
            stream_equal = stream_equal and np.array_equal(o["uPix"], u.cpu().numpy()) and np.array_equal(o["vPix"], v.cpu().numpy()) \
                and np.array_equal(o["uVal"], sh[0].cpu().numpy()) and np.array_equal(o["vVal2"], sh[3].cpu().numpy())
    # OCTANE_EHALO is not sticky: a halo sized for a displacement bound that is too small fails on the ranks whose warp
    # leaves the band, and the SAME context solves again after the caller raised max_disp (a collective re-plan)
    halo_recovers = True
    if not with_fg:
        tight = ob.default_params(max_disp=0)
        o0, o1, i0, i1 = ob.band_plan(nx, ny, tight, rank, world)
        t1_, t2_ = S.make_pair_torch(nx, ny, 9, dev, rows=(i0, i1), drift=(0.8, -9.0))     # 9 px of vertical motion
        ut = torch.zeros((o1 - o0, nx), device=dev); vt = torch.zeros_like(ut)
        failed = 0
        try:
            ctx.oct_variational_optical_flow_band(t1_, t2_, ut, vt, nx, ny, tight)
        except ob.OctaneError as e:
            failed = int(e.code == -5)
        flags = [None] * world
        dist.all_gather_object(flags, failed)
        u3 = torch.zeros_like(u); v3 = torch.zeros_like(v)
        ctx.oct_variational_optical_flow_band(img1, img2, u3, v3, nx, ny, p)               # the original plan again: must succeed
        ctx.synchronize()
        halo_recovers = any(flags) and bool(torch.equal(u3, u) and torch.equal(v3, v))
    # a re-plan that fails on ONE rank (no device memory for the larger workspace) must come back as an error on every
    # rank -- the failing rank votes in the workspace-mapping exchange instead of leaving it -- and the context must
    # solve again afterwards
    replan_recovers = True
    if not with_fg and world > 1:
        ny2 = 3 * ny                           # a workspace three times the current one ...
        o0, o1, i0, i1 = ob.band_plan(nx, ny2, p, rank, world)
        b1, b2 = S.make_pair_torch(nx, ny2, 9, dev, rows=(i0, i1))
        ub = torch.zeros((o1 - o0, nx), device=dev); vb = torch.zeros_like(ub)
        hog = None
        if rank == world - 1:                  # ... cannot be allocated on the last rank: 8 MiB + the old one are free
            torch.cuda.synchronize(dev)
            torch.cuda.empty_cache()
            free, _ = torch.cuda.mem_get_info(dev)
            hog = torch.empty(max(free - (8 << 20), 0), dtype=torch.uint8, device=dev)
        code = 0
        try:
            ctx.oct_variational_optical_flow_band(b1, b2, ub, vb, nx, ny2, p)
            ctx.synchronize()
        except ob.OctaneError as e:
            code = e.code
        codes = [None] * world
        dist.all_gather_object(codes, code)
        del hog, b1, b2, ub, vb
        torch.cuda.empty_cache()
        u4 = torch.zeros_like(u); v4 = torch.zeros_like(v)
        ctx.oct_variational_optical_flow_band(img1, img2, u4, v4, nx, ny, p)
        ctx.synchronize()
        replan_recovers = (codes[world - 1] == -3 and all(c_ != 0 for c_ in codes)
                           and bool(torch.equal(u4, u) and torch.equal(v4, v)))
        if not replan_recovers:
            print(f"REPLAN rank {rank}: codes {codes}", flush=True)
    parts = [None] * world
    dist.all_gather_object(parts, (own0, own1, u.cpu().numpy(), v.cpu().numpy(), sh[0].cpu().numpy(),
                                   repro and stream_equal and halo_recovers and replan_recovers, list(st.cg_iterations[:st.n_solves])))
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        U = np.concatenate([t[2] for t in parts]); V = np.concatenate([t[3] for t in parts])
        SU = np.concatenate([t[4] for t in parts])
        # single-GPU reference on rank 0 with a fresh, un-banded context
        c1 = ob.Context(local)
        a, b = S.make_pair_torch(nx, ny, 9, dev)
        u1 = torch.zeros((ny, nx), device=dev); v1 = torch.zeros_like(u1)
        if with_fg:
            u1, v1 = first_guess(0, ny)
        c1.oct_variational_optical_flow(a, b, u1, v1, p)
        s1 = [torch.zeros((ny, nx), dtype=torch.int16, device=dev) for _ in range(4)]
        c1.oct_pix2uv_cuda(nav, 0.0, dt, u1, v1, *s1, p)
        c1.synchronize()
        st1 = c1.stats()
        du = np.abs(U - u1.cpu().numpy()); dv = np.abs(V - v1.cpu().numpy())
        res = dict(world=world, nx=nx, ny=ny, du_mean=float(du.mean()), du_max=float(du.max()), dv_mean=float(dv.mean()),
                   dv_max=float(dv.max()), repro=all(t[5] for t in parts),
                   its_equal=all(t[6] == list(st1.cg_iterations[:st1.n_solves]) for t in parts),
                   nav_max=int(np.abs(SU.astype(int) - s1[0].cpu().numpy()).max()),
                   bands=[(t[0], t[1]) for t in parts])
        print("BAND_RESULT " + json.dumps(res), flush=True)
        c1.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
