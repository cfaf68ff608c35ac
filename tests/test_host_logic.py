"""Host-side logic that needs no GPU: drop-in module surface (names, state_dict keys and shapes,
loud failure without CUDA), flat parameter re-homing, synthetic generator."""
import numpy as np
import pytest
import torch

from common import SCALES
from lshm_b200 import synthetic as S
from oracle import lofar_oracle as O


@pytest.mark.parametrize("ndim", [2, 1])
def test_module_surface_matches_reference_layout(ndim):
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
    hs = torch.tensor(SCALES)
    cls = AutoEncoderCNN2 if ndim == 2 else AutoEncoder1DCNN
    for C, L, rica in ((8, 32, True), (4, 224, True), (4, 16, False)):
        net = cls(latent_dim=L, channels=C, harmonic_scales=hs, rica=rica)
        shapes = O.ae_param_shapes(L, C, 16, rica, ndim)   # = reference state_dict (see test_oracle_golden)
        sd = net.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
        assert net.rica == rica and net.latent_dim == L and net.harmonic_dim == 16
        assert net.harmonic_scales is hs
    km = Kmeans(latent_dim=64, K=10, p=4)
    assert list(km.state_dict().keys()) == ["M"] and km.M.shape == (10, 64)
    assert (km.K, km.p, km.EPS, km.latent_dim) == (10, 4, 1e-9, 64)
    assert float(km.M.min()) >= 0 and float(km.M.max()) <= 1


def test_cpu_tensors_fail_loudly():
    from lshm_b200.lofar_models import AutoEncoderCNN2, Kmeans
    net = AutoEncoderCNN2(32, 8, torch.tensor(SCALES), True)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 8, 128, 128), torch.zeros(1, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        Kmeans(64, 10, 4)(torch.zeros(3, 64))


def test_flat_params_keep_parameters_ordinary_leaves():
    from lshm_b200.kharmonic_lofar import FlatParams
    from lshm_b200.lofar_models import AutoEncoder1DCNN, Kmeans
    net = AutoEncoder1DCNN(16, 8, torch.tensor(SCALES), True)
    km = Kmeans(48, 5, 4)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams([net, km], torch.device("cpu"))
    assert all(p.is_leaf and p.requires_grad for p in flat.params)
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    assert all(o % FlatParams.ALIGN == 0 for o in flat.offsets)
    # in-place updates through the views reach the flat buffer (LBFGSNew: p.data.add_, src/lbfgsnew.py:101)
    flat.params[0].data.add_(1.0)
    o = flat.offsets[0]
    assert torch.equal(flat.flat[o:o + flat.params[0].numel()].view_as(flat.params[0]), flat.params[0].data)
    # optimizer.zero_grad(set_to_none=True) detaches .grad; attach_grads restores the views
    for p in flat.params:
        p.grad = None
    flat.attach_grads()
    assert all(p.grad is gv for p, gv in zip(flat.params, flat.grad_views))
    flat.grad.fill_(2.0)
    assert float(flat.params[-1].grad.sum()) == 2.0 * flat.params[-1].numel()
    assert flat.loss_tail.numel() == 16


def test_synthetic_measurement_layout():
    m = S.make_measurement(5, 140, 130, seed=1)
    sap = m["measurement"]["saps"]["0"]
    assert sap["visibilities"].shape == (5, 140, 130, 4, 2) and sap["visibilities"].dtype == np.int8
    assert sap["visibility_scale_factors"].shape == (5, 130, 4) and sap["visibility_scale_factors"].min() > 0
    assert sap["baselines"].shape == (5, 2)
    assert all(nm in sap["antenna_locations"]["XYZ"] for nm in sap["baselines"].reshape(-1))
    assert np.array_equal(S.make_measurement(5, 140, 130, seed=1)["measurement"]["saps"]["0"]["visibilities"],
                          sap["visibilities"])


def test_fastdiv_constants_match_integer_division():
    """The multiply-shift division of the conv kernels' index math (conv_geom.cuh), checked on the host through
    the library: divisors that occur ((h+1)(w+1), w+1, w, tiles per row, slots) and random ones, dividends up to
    the 2^31 limit the launchers enforce."""
    import ctypes
    from lshm_b200 import _lib
    # self-test hook declared in include/lshm_selftest.h (not part of the product ABI, include/lshm.h)
    check = ctypes.CDLL(_lib.LIBRARY).lshm_fastdiv_check
    check.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    check.restype = ctypes.c_int
    assert "lshm_fastdiv_check" not in _lib.parse_header()
    rng = np.random.default_rng(7)
    divisors = [1, 2, 3, 4, 5, 7, 9, 17, 33, 65, 129, 200, 328, 25, 81, 289, 1089, 4225, 16641, 4096, 16384,
                (1 << 31) - 1] + [int(v) for v in rng.integers(1, 1 << 20, 200)]
    for d in divisors:
        n = np.concatenate([np.array([0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - 4097], np.int64),
                            rng.integers(0, 1 << 31, 2000).astype(np.int64),
                            (np.arange(1, 50, dtype=np.int64) * d - 1), np.arange(1, 50, dtype=np.int64) * d])
        n = np.ascontiguousarray(n[(n >= 0) & (n < (1 << 31))])
        bad = ctypes.c_int64(-1)
        assert check(d, n.ctypes.data, len(n), ctypes.addressof(bad)) == 0
        assert bad.value == 0, (d, bad.value)


def test_workspace_rows_are_contiguous_views():
    """A micro-batch workspace is the same memory restricted to rows [n0, n1) of every buffer (engine.Workspace.rows)."""
    from lshm_b200.engine import Workspace
    w = Workspace(8, 8, 32, 16, "cpu", True, True)
    v = w.rows(2, 6)
    assert v.N == 4 and v.sizes == w.sizes and hasattr(v, "g_enc") and v.enc[0] is None
    for name in ("xhat", "cat1", "mu", "zcat", "g_cat1", "dx", "uvh"):
        a, b = getattr(v, name), getattr(w, name)
        assert a.shape[0] == 4 and a.is_contiguous() and a.data_ptr() == b[2:].data_ptr(), name
    for lst in ("enc", "dec", "g_enc", "g_dec"):
        for a, b in zip(getattr(v, lst), getattr(w, lst)):
            assert (a is None and b is None) or a.data_ptr() == b[2:].data_ptr()
    w.xhat.zero_()
    v.xhat.fill_(3.0)
    assert float(w.xhat[2:6].min()) == 3.0 and float(w.xhat[:2].abs().max()) == 0.0 and float(w.xhat[6:].abs().max()) == 0.0
    no_grad = Workspace(4, 4, 16, 16, "cpu", False, False).rows(0, 2)
    assert not hasattr(no_grad, "g_enc")


def test_fft_features_argument_checks():
    from lshm_b200 import lofar_tools as T
    with pytest.raises(RuntimeError, match="CUDA"):
        T.fft_features(torch.zeros(1, 2, 128, 128))
