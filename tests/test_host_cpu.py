"""CPU: host-side logic of the drop-in modules, the C-ABI surface, the synthetic generator and the roofline formulas."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import load_golden, module_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from molkgnn_b200 import build, _lib
    build.build()
    hdr = open(os.path.join(ROOT, "include", "molkgnn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(molkgnn_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 14
    so = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), name
    assert sorted(_lib.EXPORTS) == declared          # the ctypes binding covers the whole header
    l = _lib.lib()
    assert l.molkgnn_version() >= 100
    assert l.molkgnn_packed_floats(4, 50, 112) > 0 and l.molkgnn_packed_floats(5, 1, 4) == -1
    # the ctypes mirrors of the header's structs have the compiled sizes
    for which, cls in enumerate([_lib.Plan, _lib.Layer, _lib.LayerGrads, _lib.StackLayout]):
        assert l.molkgnn_struct_bytes(which) == ctypes.sizeof(cls), cls.__name__


def test_state_dict_keys_and_init_order_match_reference():
    """Parameter names are pinned by model.py:373-381,426-428 and kernel_reader.py:86; init RNG order by
    kernels.py:50-53,762-774 (same seed => same kernels as the reference, so seeds/checkpoints carry over)."""
    import molkgnn_b200 as mk
    g = load_golden("molgcn_small")
    net = module_from_golden(g, "cpu")
    ref_keys = sorted(k[len("param_"):] for k in g if k.startswith("param_"))
    assert sorted(net.state_dict().keys()) == ref_keys
    torch.manual_seed(int(g["seed"]))
    L1, LN = [int(v) for v in g["L1"]], [int(v) for v in g["LN"]]
    fresh = mk.MolGCN(int(g["num_layers"]), *L1, *LN, x_dim=28, p_dim=3, edge_attr_dim=7)
    for k, v in fresh.state_dict().items():
        if k.endswith("_weight"):
            continue   # make_golden.py perturbs the scalar mixing weights after construction
        assert np.array_equal(v.numpy(), g["param_" + k]), k
    assert fresh.layers[0].get_num_kernel() == sum(L1)
    assert fresh.num_kernels(1) == sum(LN)


def test_reference_error_conventions():
    import molkgnn_b200 as mk
    from molkgnn_b200._lib import MolKGNNError
    with pytest.raises(Exception):
        mk.KernelConv(L=3, D=3)                       # kernels.py:43-48
    with pytest.raises(Exception):
        mk.MolGCN(num_layers=0)                       # KernelLayer.py:16-17
    net = mk.MolGCN(1, 2, 2, 2, 2, 2, 2, 2, 2, x_dim=28, p_dim=3, edge_attr_dim=7)
    with pytest.raises(Exception):
        net(torch.zeros(3, 28))                       # KernelLayer.py:54-57
    with pytest.raises(Exception):
        net.layers[0](False, torch.zeros(3, 28))      # kernels.py:617-620
    with pytest.raises(MolKGNNError):                 # no CPU fallback: CPU tensors are refused loudly
        net(x=torch.zeros(2, 28), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.ones(2, 7),
            p=torch.zeros(2, 3), save_score=False)
    kc = mk.KernelConv(L=2, D=3, num_supports=2, node_attr_dim=4, edge_attr_dim=7)
    with pytest.raises(Exception):                    # kernels.py:443-445 coordinate dimension mismatch
        kc(False, x_focal=torch.zeros(1, 4), p_focal=torch.zeros(1, 2), x_neighbor=torch.zeros(1, 2, 4),
           p_neighbor=torch.zeros(1, 2, 2), edge_attr_neighbor=torch.zeros(1, 2, 7))


def test_synth_generator_contract():
    from molkgnn_b200 import synth
    b = synth.make_batch(64, seed=0)
    ei = b["edge_index"]
    N = b["x"].shape[0]
    assert b["x"].shape[1] == 28 and b["edge_attr"].shape[1] == 7 and b["p"].shape[1] == 3
    deg = np.bincount(ei[0], minlength=N)
    assert deg.min() >= 1 and deg.max() <= 4
    # bond b on rows 2b:(i,j), 2b+1:(j,i) with identical attributes (wrapper.py:152-156)
    assert np.array_equal(ei[0, 0::2], ei[1, 1::2]) and np.array_equal(ei[1, 0::2], ei[0, 1::2])
    assert np.array_equal(b["edge_attr"][0::2], b["edge_attr"][1::2])
    assert (np.linalg.norm(b["edge_attr"], axis=1) > 0).all()
    # no edge crosses a molecule
    assert np.array_equal(b["batch"][ei[0]], b["batch"][ei[1]])
    sizes = np.bincount(b["batch"])
    assert sizes.min() >= 18 and sizes.max() <= 32
    b2 = synth.make_batch(64, seed=0)
    assert all(np.array_equal(b[k], b2[k]) for k in b)
    hist = np.bincount(deg, minlength=5)[1:] / N
    assert 0.1 < hist[0] < 0.35 and 0.3 < hist[1] < 0.6 and hist[3] < 0.15
    b3 = synth.make_batch(200, seed=3, pool=16)
    assert len(np.bincount(b3["batch"])) == 200


def test_roofline_formula_matches_survey_worked_numbers():
    from molkgnn_b200 import roofline
    # canonical molecule N=25, E=54, n=(6,11,6,2): base 3-layer fwd 65 264 B, bwd 89 764 B
    fwd, bwd = roofline.stack_bytes(N=25, E=54, n=(6, 11, 6, 2), x_dim=28, L1=(10, 20, 30, 50), LN=(10, 20, 30, 50),
                                    num_layers=3)
    assert (fwd, bwd) == (65264, 89764)
    fwd, bwd = roofline.stack_bytes(N=25, E=54, n=(6, 11, 6, 2), x_dim=28, L1=(40, 80, 120, 200),
                                    LN=(40, 80, 120, 200), num_layers=5)
    assert (fwd, bwd) == (419440, 597940)


def test_tile_schedule_model_properties():
    """The balanced tile schedule (model of k_tile_order; the GPU test compares the kernel with this model entry by entry):
    a permutation; and when a few CTAs walk one tile more than the rest, those CTAs get the small tiles, so that the largest
    per-CTA node total stays close to the mean instead of one full tile above it."""
    from tests.helpers import tile_schedule_model
    rng = np.random.default_rng(0)
    for n, G in [(1, 1), (5, 5), (148, 148), (149, 148), (887, 148), (892, 148), (890, 148), (3000, 148), (37, 8)]:
        nn = rng.integers(97, 129, size=n)
        order = tile_schedule_model(nn, G)
        assert np.array_equal(np.sort(order), np.arange(n))
        q, r = divmod(n, G)
        load, pos = np.zeros(G), 0
        for c in range(G):
            cnt = q + (1 if c < r else 0)
            load[c] = nn[order[pos:pos + cnt]].sum()
            pos += cnt
        rr = np.zeros(G)
        for t in range(n):
            rr[t % G] += nn[t]
        if q >= 4:
            assert load.max() <= rr.max()
            if 0 < r <= G // 2:
                assert load.max() / load.mean() < 1.03 < rr.max() / rr.mean()


def test_tile_image_sizes_for_base_wide_and_oversized_layers():
    """Host logic of the layer classification (no GPU): the base model takes the resident-operand tile kernels, the configs[2]
    model the streamed-operand wide kernels (stage-major forward + K-step-major backward images + bond table), a kernel set
    beyond 16 blocks neither (the bucket-order kernels take it)."""
    import ctypes as C
    from molkgnn_b200 import _lib

    def layer(F, L):
        ly = _lib.Layer()
        ly.F, ly.Fp, ly.Fe, ly.K = F, (F + 3) // 4 * 4, 7, sum(L)
        ko = 0
        for d in range(4):
            ly.L[d] = L[d]
            ly.koff[d] = ko
            ko += L[d]
        return ly

    lib = _lib.lib()
    base = lib.molkgnn_tile_img_bytes(C.byref(layer(110, (10, 20, 30, 50))))
    assert base > 0
    for F in (28, 440):
        Fk = (F + 31) // 32 * 32
        got = lib.molkgnn_tile_img_bytes(C.byref(layer(F, (40, 80, 120, 200))))
        nb = 8 + 4 + 2 + 1                                         # 25 / 30 / 40 / 40 kernels per 128-row block
        es = 2 * (40 + 2 * 80 + 3 * 120 + 4 * 200) * 16
        want = nb * (Fk // 32) * 16384 + (es + 127) // 128 * 128 + nb * 8 * Fk * 64
        assert got == want, (F, got, want)
    assert lib.molkgnn_tile_img_bytes(C.byref(layer(440, (400, 800, 1200, 2000)))) == 0


def test_bench_clock_sampler_window():
    """bench.ClockSampler.stop: the lines inside the timed region count; if the region was shorter than nvidia-smi's first period
    (no line inside), the lines next to it are used and the record says so -- never an empty `clocks` record while lines exist."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _P(object):
        def terminate(self):
            pass

    def sampler(rows):
        s = bench.ClockSampler.__new__(bench.ClockSampler)
        s.rows, s.p = rows, _P()
        return s

    line = "1965, 1965, 400.0, Not Active, Not Active, Not Active, Not Active"
    hot = "1200, 1965, 700.0, Not Active, Active, Not Active, Active"
    r = sampler([(9.0, line), (10.01, line), (10.1, hot), (10.19, line), (11.0, line)]).stop(10.0, 10.2)
    assert r["samples"] == 3 and r["window_s"] == 0.05 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    r = sampler([(9.8, line), (10.4, line), (12.0, hot)]).stop(10.0, 10.02)        # nothing inside: the neighbours, not the far line
    assert r["samples"] == 2 and r["window_s"] == 0.5 and r["sm_mhz"] == 1965.0 and r["reasons"] == []
    r = sampler([]).stop(10.0, 10.2)
    assert r["samples"] == 0 and r["sm_mhz"] is None
