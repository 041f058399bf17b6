"""CPU tests of the boundary and the host logic (no compute calls without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import octane_b200 as ob
from octane_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "octane_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(octane_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/octane_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "python binding list and header disagree"
    assert L.octane_abi_version() == _lib.EXPECTED_ABI == 3


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "octane_b200.h")).read()
    assert "torch" not in src.replace("no C++ / torch types", "") and "std::" not in src and "extern \"C\"" in src


def test_struct_layout_matches_header():
    # sizes computed by the C compiler vs ctypes
    code = r'''
#include <stdio.h>
#include "octane_b200.h"
int main(){printf("%zu %zu %zu %zu\n", sizeof(octane_params), sizeof(octane_nav), sizeof(octane_stats), sizeof(octane_cal));return 0;}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        a, b, c, e = map(int, subprocess.check_output([os.path.join(d, "t")]).split())
    assert (a, b, c, e) == (C.sizeof(ob.Params), C.sizeof(ob.Nav), C.sizeof(ob.Stats), C.sizeof(ob.Cal))


def test_defaults_are_the_reference_defaults():
    p = ob.default_params()      # src/main.cc:77-98
    assert (p.alpha, p.lambda_, p.lambdac, p.scaleF) == (5.0, 1.0, 0.0, 0.5)
    assert (p.kiters, p.liters, p.cgiters, p.dozim) == (4, 3, 30, 1)
    assert (p.pixuv, p.dopolar, p.domerc, p.setdevice) == (0, 0, 0, 0)


def test_level_dims_bit_exact(oracle):
    for nx, ny in [(500, 500), (2000, 2000), (10000, 6000), (21696, 21696), (63, 70), (1001, 333), (4097, 257)]:
        for k in (1, 2, 3, 4, 5):
            p = ob.default_params(kiters=k)
            if min(nx, ny) * 0.5 ** (k - 1) < 4:
                continue
            assert ob.level_dims(nx, ny, p) == oracle.level_dims(nx, ny, kiters=k)
    assert ob.level_dims(21696, 21696) == [(2712, 2712), (5424, 5424), (10848, 10848), (21696, 21696)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ob.OctaneError) as e:
        ob.Context(0)
    assert e.value.code == -1          # ENODEV: the product path fails loudly without CUDA


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "octane_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("no python or cpu fallback", ""), f"{f} mentions the oracle"


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(21696, 21696), (10000, 6000), (2000, 2000)])
def test_band_plan_tiles_the_scene(shape, world):
    nx, ny = shape
    p = ob.default_params()
    prev = 0
    for r in range(world):
        own0, own1, in0, in1 = ob.band_plan(nx, ny, p, r, world)
        assert own0 == prev and own1 > own0
        assert 0 <= in0 <= own0 and own1 <= in1 <= ny
        if world > 1:
            # blur radius 5 at full res for the coarsest level's halo: (max_disp/8 + 3 + 4 rows) * 8 + 5
            if r > 0:
                assert own0 - in0 >= p.max_disp + 5
            if r < world - 1:
                assert in1 - own1 >= p.max_disp + 5
        else:
            assert (in0, in1) == (0, ny)
        prev = own1
    assert prev == ny


def test_band_plan_rejects_thin_bands():
    with pytest.raises(ob.OctaneError):
        ob.band_plan(256, 128, ob.default_params(), 0, 8)     # coarsest level: 16 rows / 8 ranks


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # bootstrap pattern of bench.py: rank 0 makes the 128-byte id, everyone receives it
        ids = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nx, ny = 640, 512
        p = ob.default_params(max_disp=8)
        own0, own1, in0, in1 = ob.band_plan(nx, ny, p, rank, world)
        # each rank "solves" its band: here the band is filled with its global row index
        import torch
        band = torch.arange(own0, own1, dtype=torch.float32)[:, None].expand(own1 - own0, nx).contiguous()
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([own1 - own0]))
        parts = [torch.zeros((int(s.item()), nx)) for s in sizes]
        # uneven bands: gather through object lists
        objs = [None] * world
        dist.all_gather_object(objs, band.numpy())
        full = np.concatenate(objs, axis=0)
        ok = full.shape == (ny, nx) and np.array_equal(full[:, 0], np.arange(ny, dtype=np.float32))
        # time reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ids[0] == bytes(range(128)), ok, float(t.item()), (own0, own1, in0, in1)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_band_bootstrap_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    assert all(r[1] and r[2] for r in res)
    assert all(r[3] == 2.0 for r in res)
    assert res[0][4][1] == res[1][4][0] == 256           # bands meet at the middle row
    assert res[0][4][3] > 256 and res[1][4][2] < 256     # and overlap in their inputs


def test_exact_division_shortcuts(oracle):
    """build.cu replaces the fp64 divisions the reference's promotion rules force with cheaper
    sequences that must give the SAME float/double: float(1./double(s)) == 1.0f/s for every float s
    (sampled with a stride over all bit patterns), and x*RN(1/a) + one Markstein step == double(x)/a."""
    import ctypes as C
    L = oracle.lib()
    L.oracle_check_recip_float.argtypes = [C.c_uint, C.c_uint]
    L.oracle_check_recip_float.restype = C.c_long
    L.oracle_check_div_const.argtypes = [C.c_double, C.c_ulonglong, C.c_long]
    L.oracle_check_div_const.restype = C.c_long
    assert L.oracle_check_recip_float(61, 7) == 0
    for a in (5.0, 3.0, 7.3, 0.1, 1e-3, 15.0, 1.0, 2.5, 123.456):
        assert L.oracle_check_div_const(a, 12345, 2_000_000) == 0
    # pyramid.cu div12: the gradient's double numerator / 12.0 without a division
    L.oracle_check_div12.argtypes = [C.c_ulonglong, C.c_long, C.c_int]
    L.oracle_check_div12.restype = C.c_long
    assert L.oracle_check_div12(777, 3_000_000, 0) == 0
    assert L.oracle_check_div12(778, 3_000_000, 1) == 0


def test_band_table_matches_reference_values():
    """octane_band_minmax == oct_bandminmax (reference src/oct_normalize_geo.cc:9-88); needs no GPU."""
    assert ob.band_minmax(2) == pytest.approx((628.98723908, -20.28991094), rel=1e-7)
    assert ob.band_minmax(13) == pytest.approx((185.5699, -1.6443), rel=1e-7)
    assert ob.band_minmax(7) == (2.0, 0.0) and ob.band_minmax(8) == (6.0, 3.0)    # the "meteorological" ranges
    with pytest.raises(ob.OctaneError):
        ob.band_minmax(17)


def test_host_mirror_refuses_arrays_the_abi_would_misread():
    """the C ABI takes raw pointers: the Python mirror rejects wrong element types, strided views and host / device
    mix-ups before anything reaches the library (octane_b200.api._require)"""
    import torch
    from octane_b200.api import _require
    a = np.zeros((4, 6), np.float32)
    _require("a", a, "float32"); _require("a", a, "float32", cuda=False); _require("none", None, "int16", cuda=True)
    _require("s", np.zeros((2, 2), np.int16), "int16")
    t = torch.zeros((4, 6))
    _require("t", t, "float32", cuda=False)
    bad = [(a.astype(np.float64), "float32", None), (a[:, ::2], "float32", None), (a.T, "float32", None),
           ([1.0, 2.0], "float32", None), (a, "float32", True), (a, "int16", None),
           (t, "float32", True), (t.double(), "float32", None), (t.t(), "float32", None), (t, "int16", None)]
    for x, dt, cuda in bad:
        with pytest.raises(TypeError):
            _require("x", x, dt, cuda=cuda)


def test_plans_that_cannot_run_are_refused_on_the_host():
    """octane_workspace_bytes makes the same plan a solve would: 0 + octane_last_error for parameter sets the
    kernels cannot run (reference: no checks at all, src/oct_variational_optical_flow.cu:1229-1266)"""
    import ctypes as C
    L = _lib.load()

    def ws(nx, ny, nc=1, **kw):
        p = ob.default_params(**kw)
        return L.octane_workspace_bytes(nx, ny, nc, C.byref(p))

    assert ws(21696, 21696) > 40 * 2**30 and ws(21696, 21696, 3) < 90 * 2**30      # 108 B/px and the 3-channel case fit 180 GB
    assert ws(21696, 21696, kiters=8) > 0
    assert ws(21696, 21696, kiters=9) == 0 and b"pyramid too deep" in L.octane_last_error()
    assert ws(20, 20) == 0 and b"smaller than 4 pixels" in L.octane_last_error()      # 20 * 0.125 = 3 (rounded)
    assert ws(64, 64, 4) == 0 and ws(64, 64, alpha=0.0) == 0 and ws(64, 64, scaleF=1.0) == 0
    assert ws(64, 64, kiters=4, liters=30) == 0                                        # more than OCTANE_MAX_SOLVES solves


def test_band_plan_properties_over_random_scenes():
    """randomised scenes, world sizes, pyramid depths and displacement bounds: either the plan is refused (a band
    thinner than its halo, a level smaller than 4 pixels) or the bands tile the scene and every rank's input rows
    cover what its coarsest level reads -- blur radius R_k at full resolution around (own rows +- warp halo) / factor"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(nx=st.integers(64, 4000), ny=st.integers(64, 30000), world=st.integers(1, 8), kiters=st.integers(1, 5),
           max_disp=st.integers(0, 96))
    def check(nx, ny, world, kiters, max_disp):
        p = ob.default_params(kiters=kiters, max_disp=max_disp)
        plans = []
        try:
            for r in range(world):
                plans.append(ob.band_plan(nx, ny, p, r, world))
        except ob.OctaneError as e:
            assert e.code == -2                      # refused for every rank alike or not at all is checked below
            return
        prev = 0
        for r, (own0, own1, in0, in1) in enumerate(plans):
            assert own0 == prev and own1 > own0 and 0 <= in0 <= own0 and own1 <= in1 <= ny
            prev = own1
            if world == 1:
                assert (in0, in1) == (0, ny)
                continue
            # finest level: warp halo ceil(max_disp) + 3 rows and 4 more for the second derivatives
            need = max_disp + 7
            assert own0 - in0 >= min(need, own0) and in1 - own1 >= min(need, ny - own1)
            for k in range(kiters - 1):              # coarser levels read full-resolution rows through the blur
                f = 0.5 ** (kiters - 1 - k)
                yk = int(ny * f + 0.5)
                o0, o1 = r * yk // world, (r + 1) * yk // world
                H = max(int(np.ceil(max_disp * f)) + 7, 4)
                lo, hi = max(o0 - H, 0), min(o1 + H, yk)
                R = max(5, int(2 * np.float32(1.0 / np.sqrt(2.0 * f))))
                assert in0 <= max(int(lo / f) - R, 0) and in1 >= min(int((hi - 1) / f) + R, ny)
        assert prev == ny

    check()
