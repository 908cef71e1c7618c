"""CPU tests: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from odefilters_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "pnde.h")).read()
    declared = set(re.findall(r"\b(pnde_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pnde_config", "pnde_handle"}
    lib = _lib.load()
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_layout_matches_header():
    from odefilters_b200 import _lib

    cfg = _lib.PndeConfig()
    lib = _lib.load()
    assert lib.pnde_default_config(C.byref(cfg), _lib.ALG_EK1, 3, 0) == 0
    assert (cfg.abi_version, cfg.alg, cfg.order, cfg.adaptive) == (2, 1, 3, 1)
    assert (cfg.abstol, cfg.reltol, cfg.qmax, cfg.gamma, cfg.qoldinit) == (1e-6, 1e-3, 10.0, 0.9, 1e-4)
    assert cfg.maxiters == 100000 and C.sizeof(cfg) == 12 * 4 + 15 * 8 + 2 * 8 + (2 + 16) * 4
    assert (cfg.n_devices, cfg.flags) == (0, 0)


def test_create_argument_errors_are_reported_without_a_gpu():
    from odefilters_b200 import _lib

    lib = _lib.load()
    cfg = _lib.PndeConfig()
    lib.pnde_default_config(C.byref(cfg), _lib.ALG_EK1, 3, 0)
    h = C.c_void_p()
    cfg.diffusion = _lib.DIFFUSIONS["dynamicMV"]  # EK1 + MV: the reference asserts (src/diffusions.jl:97)
    assert lib.pnde_create(C.byref(cfg), C.byref(h)) == -1 and b"EK0-only" in lib.pnde_last_error(None)
    cfg.diffusion = 0
    cfg.adaptive, cfg.dt = 0, 0.0  # test/errors.jl:16-20
    assert lib.pnde_create(C.byref(cfg), C.byref(h)) == -1 and b"require a choice of dt" in lib.pnde_last_error(None)
    cfg.adaptive = 1
    cfg.order = 9
    assert lib.pnde_create(C.byref(cfg), C.byref(h)) == -3


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import odefilters_b200 as B

    with pytest.raises(RuntimeError, match="no CUDA device"):
        B.solve(B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 1.0), (0.2, 0.2, 3.0)), B.EK1())


def test_host_argument_checks():
    import odefilters_b200 as B

    with pytest.raises(ValueError):  # test/errors.jl:11-14 (non-vector u0)
        B.ODEProblem("fhn_readme", [[-1.0, 1.0]], (0.0, 1.0), (0.2, 0.2, 3.0))
    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 1.0), (0.2, 0.2, 3.0))
    with pytest.raises(ValueError, match="choice of dt"):
        B.solve(prob, B.EK0(), adaptive=False)
    with pytest.raises(ValueError):
        B.EK1(diffusionmodel="nope")
    with pytest.raises(ValueError, match="dense"):
        B.solve(prob, B.EK1(smooth=True), dense=False)


def test_shard_range_partitions():
    import odefilters_b200 as B

    for n in (1, 7, 1000, 10 ** 6 + 3):
        for w in (1, 2, 4, 8):
            parts = [B.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_unpack_lower():
    from odefilters_b200.api import _unpack_lower

    D = 4
    A = np.arange(16.0).reshape(4, 4)
    A = A + A.T
    il = np.tril_indices(D)
    assert np.array_equal(_unpack_lower(A[il][None], D)[0], A)


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist

    import odefilters_b200 as B

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = B.shard_range(n, rank, world)
    import torch

    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt)
    mx = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    q.put((rank, int(cnt.item()), float(mx.item()), lo, hi))
    dist.destroy_process_group()


def test_world_size_2_sharding_over_gloo():
    """The N > 1 path of bench.py: shard by rank, no data-path collective, max-over-ranks timing."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    n = 1001
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    assert [r[1] for r in res] == [n, n] and [r[2] for r in res] == [2.0, 2.0]
    assert res[0][3:] == (0, 500) and res[1][3:] == (500, 1001)


def test_custom_vector_field_source_is_validated_without_gpu():
    """pnde_check_custom: NVRTC compile-only validation (needs libnvrtc, no device)."""
    import odefilters_b200 as B

    ok = B.CustomVectorField(2, 1, "du[0] = u[1]; du[1] = -p[0]*sin(u[0]);",
                             "J[0][0] = 0.0; J[0][1] = 1.0; J[1][0] = -p[0]*cos(u[0]); J[1][1] = 0.0;")
    assert ok.check(B.EK1(order=2)) == ""
    assert ok.check(B.EK0(order=3, diffusionmodel="dynamicMV")) == ""
    bad = B.CustomVectorField(2, 1, "du[0] = u[1] du[1] = 0.0;")
    with pytest.raises(ValueError, match="expected a"):
        bad.check(B.EK0(order=2))


def test_ieks_kernels_compile_without_gpu():
    """The IEKS flavour of the filter kernel (ieks_kernel.cuh: linearisation at the previous iterate's dense output) is
    compiled on demand; the compile-only check runs here."""
    import odefilters_b200 as B

    vf = B.CustomVectorField(2, 1, "du[0] = u[1]; du[1] = -p[0]*sin(u[0]);",
                             "J[0][0] = 0.0; J[0][1] = 1.0; J[1][0] = -p[0]*cos(u[0]); J[1][1] = 0.0;")
    assert vf.check(B.IEKS(order=2)) == ""
    with pytest.raises(ValueError):
        B.solve_ieks(B.ODEProblem("fhn_lib", [1.0, 1.0], (0.0, 1.0), (0.7, 0.8, 0.08, 0.5)), B.EK1(order=2))


def test_gaussian_list_moments():
    """mean / var / std of SRGaussian and SRGaussianList (src/ProbNumDiffEq.jl:61-66)."""
    import odefilters_b200 as B
    from odefilters_b200.api import _GaussianList

    rng = np.random.default_rng(0)
    A = rng.standard_normal((5, 3, 3))
    cov = A @ A.transpose(0, 2, 1)
    gl = _GaussianList(rng.standard_normal((5, 3)), cov)
    assert np.array_equal(B.mean(gl), gl.mu)
    assert np.allclose(B.var(gl), np.stack([np.diag(c) for c in cov]))
    assert np.allclose(B.std(gl[2]), np.sqrt(np.diag(cov[2])))


def test_strided_sharding_covers_the_ensemble_once():
    import odefilters_b200 as B

    n, world = 1003, 4
    parts = [B.shard_strided(n, r, world) for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))
    assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    with pytest.raises(ValueError):
        B.shard_strided(n, 4, 4)


def test_multi_device_config_errors_without_a_gpu():
    """cfg.n_devices / device_list (ABI version 2) are validated before any device is touched."""
    import odefilters_b200 as B
    from odefilters_b200 import _lib

    prob = B.ODEProblem("fhn_readme", [-1.0, 1.0], (0.0, 1.0), (0.2, 0.2, 3.0))
    with pytest.raises(RuntimeError, match="twice"):
        B.FilterSolver(prob, B.EK1(order=3, smooth=False), save_everystep=False, devices=[0, 0])
    with pytest.raises(ValueError, match="devices"):
        B.FilterSolver(prob, B.EK1(order=3, smooth=False), save_everystep=False, devices=list(range(_lib.MAX_DEVICES + 1)))
    assert isinstance(B.api.device_count(), int)
    assert B.EnsembleB200(devices=[0, 1]).devices == [0, 1] and B.EnsembleB200().devices is None


def test_header_constants_match_the_python_binding():
    from odefilters_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "pnde.h")).read()
    get = lambda name: int(re.search(rf"#define {name} (-?\d+)", hdr).group(1))  # noqa: E731
    assert get("PNDE_ABI_VERSION") == _lib.ABI_VERSION and get("PNDE_MAX_DEVICES") == _lib.MAX_DEVICES
    assert get("PNDE_FLAG_REFERENCE_QUIRKS") == _lib.FLAG_REFERENCE_QUIRKS and get("PNDE_FLAG_ONE_THREAD") == _lib.FLAG_ONE_THREAD
    assert get("PNDE_RET_ZERO_RESIDUAL") == 6 and _lib.RETCODES[6] == "ZeroResidual"
    assert get("PNDE_ALG_IEKS") == _lib.ALG_IEKS and get("PNDE_SAVE_STRIDE") == _lib.SAVE_STRIDE
