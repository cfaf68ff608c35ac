"""Pins the CPU oracle against the golden fixtures produced from the live reference
(oracle/gen_golden.py), and - when /root/reference is present - against the reference itself."""
import os
import sys

import numpy as np
import pytest
import torch

from common import REFERENCE_SRC, SCALES, closure_case, golden, max_abs, oracle_closure, rel_err
from lshm_b200 import synthetic as S
from oracle import lofar_oracle as O

HS = torch.tensor(SCALES)


@pytest.mark.parametrize("tag", ["c8", "c4"])
def test_ae_forward_matches_golden(tag):
    g = golden(f"ae_forward_{tag}.npz")
    C, L, Lt = int(g["C"]), int(g["L"]), int(g["Lt"])
    pn = O.make_ae_params(L, C, ndim=2, seed=11)
    pT = O.make_ae_params(Lt, C, ndim=1, seed=12)
    x = torch.from_numpy(S.make_patches(2, C, seed=21))
    uv = torch.from_numpy(S.make_uv(2, seed=21))
    xh, mu = O.ae_forward(pn, x, uv, HS, 2, True)
    yT, muT = O.ae_forward(pT, torch.flatten(x, 2, 3), uv, HS, 1, True)
    assert max_abs(mu, g["mu"]) == 0 and max_abs(muT, g["muT"]) == 0
    assert max_abs(xh[:, :, ::16, ::16], g["xhat_sub"]) == 0
    assert max_abs(yT.view(2, C, 128, 128)[:, :, ::16, ::16], g["yT_sub"]) == 0


def test_kmeans_matches_golden():
    g = golden("kmeans.npz")
    X, M = torch.from_numpy(g["X"]), torch.from_numpy(g["M"])
    for p in (2, 4):
        assert abs(float(O.khm_loss(X, M, p)) - float(g[f"loss_p{p}"])) <= 2e-6 * abs(float(g[f"loss_p{p}"]))
        assert abs(float(O.khm_loss_loops(X, M, p)) - float(g[f"loss_p{p}"])) <= 1e-6 * abs(float(g[f"loss_p{p}"]))
        gx, gm = O.khm_grads_analytic(X, M, p)
        assert rel_err(gx, g[f"gX_p{p}"]) < 1e-5
        assert rel_err(gm, g[f"gM_p{p}"]) < 1e-5
    assert abs(float(O.cluster_similarity(M)) - float(g["sim"])) < 1e-6 * float(g["sim"])
    assert abs(float(O.cluster_similarity_loops(M)) - float(g["sim"])) < 1e-6 * float(g["sim"])
    Mg = M.clone().requires_grad_()
    O.cluster_similarity(Mg).backward()
    assert rel_err(Mg.grad, g["gsim"]) < 1e-5


def test_closure_matches_golden():
    g = golden("closure_cfg1.npz")
    out = oracle_closure(closure_case())
    for k in ("total", "loss0", "loss1", "loss2", "loss3", "kdist", "aug", "sim", "rica"):
        assert abs(out[k] - float(g[k])) <= 2e-6 * abs(float(g[k])), k
    assert rel_err(out["Mu"], g["Mu"]) < 1e-6
    for key, gr in out["grads"].items():
        tag, name = key.split(".", 1)
        gk = {"0": "n", "1": "T", "2": "F", "3": "k"}[tag] + "." + name
        ref_norm = float(g[gk + ":norm"])
        assert abs(gr.double().norm().item() - ref_norm) <= 1e-4 * ref_norm, key
        assert np.allclose(gr.reshape(-1)[:4].numpy(), g[gk + ":head"], rtol=1e-3, atol=1e-7 * ref_norm), key


def test_multiplier_and_eval_match_golden():
    g = golden("closure_cfg1.npz")
    c = closure_case()
    y1, y2, y3 = O.multiplier_update(c["pn"], c["pT"], c["pF"], c["x"], c["uv"], HS, *c["ys"])
    assert np.allclose(y1[:8].numpy(), g["y1_head"], rtol=1e-5, atol=1e-6)
    assert abs(y2.double().sum().item() - float(g["y2_sum"])) < 1e-3
    assert abs(y3.double().sum().item() - float(g["y3_sum"])) < 1e-3
    Mu = torch.from_numpy(g["Mu"])
    dist, idx, _ = O.eval_distances(Mu[:c["bpb"]], c["M"], 4)
    assert np.allclose(dist.numpy(), g["eval_dist"], rtol=1e-5)
    assert idx == int(g["eval_id"])


def test_loader_matches_golden():
    g = golden("loader.npz")
    meas = S.make_measurement(5, 256, 192, seed=0)
    sap = meas["measurement"]["saps"]["0"]
    for tag, C, norm in (("c8n", 8, True), ("c8", 8, False), ("c4n", 4, True)):
        np.random.seed(3)
        np.random.randint(0, 1)
        bl = np.random.randint(0, 5, 3)
        px, py, y = O.load_minibatch(sap["visibilities"], sap["visibility_scale_factors"], bl, num_channels=C,
                                     normalise=norm)
        assert [px, py] == list(g[f"{tag}:pxpy"])
        assert np.array_equal(y[:, :, ::16, ::16], g[f"{tag}:sub"])
        uv = O.uv_coordinates(sap["antenna_locations"]["XYZ"], sap["baselines"], bl, 12.5,
                              sap["central_frequencies"][96], px * py)
        assert np.array_equal(uv, g[f"{tag}:uv"])
    px, py, y = O.load_minibatch(sap["visibilities"], sap["visibility_scale_factors"], [2], clamp=1e6)
    assert np.array_equal(y[:, :, ::16, ::16], g["base2:sub"])
    meas2 = S.make_measurement(3, 100, 150, seed=1)
    sap2 = meas2["measurement"]["saps"]["0"]
    np.random.seed(4)
    np.random.randint(0, 1)
    bl = np.random.randint(0, 3, 2)
    px, py, y = O.load_minibatch(sap2["visibilities"], sap2["visibility_scale_factors"], bl, normalise=False)
    assert [px, py] == list(g["pad:pxpy"]) == [1, 1]
    assert np.array_equal(y[:, :, ::8, ::8], g["pad:sub"])


def test_fft_matches_golden():
    g = golden("fft.npz")
    x = torch.from_numpy(S.make_patches(2, 4, seed=31))
    xhat = 0.3 * torch.from_numpy(S.make_patches(2, 4, seed=32))
    y = O.fft_features(x, xhat)
    assert np.array_equal(y[:, :, ::8, ::8].numpy(), g["sub"])
    assert np.array_equal(y[:, :, 60:68, 60:68].numpy(), g["centre"])


def test_offline_update_is_a_fixed_point_improvement():
    """No reference oracle exists for offline_update (it cannot run); check the GKHM property:
    one update does not increase the K-harmonic objective."""
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((200, 16)).astype(np.float32))
    M = O.make_centres(5, 16, seed=1)
    before = float(O.khm_loss(X, M, 4))
    Mn, num, den = O.offline_update(X, M, 4)
    assert float(O.khm_loss(X, Mn, 4)) <= before
    assert torch.allclose(Mn, num / den[:, None], rtol=1e-6)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="live reference not present")
def test_oracle_against_live_reference_modules():
    sys.path.insert(0, REFERENCE_SRC)
    import lofar_models as R
    C, L = 4, 224  # the reference's own defaults, src/kharmonic_lofar.py:37,53
    pn = O.make_ae_params(L, C, ndim=2, seed=7)
    x = torch.from_numpy(S.make_patches(1, C, seed=8))
    uv = torch.from_numpy(S.make_uv(1, seed=8))
    net = R.AutoEncoderCNN2(L, C, HS, True)
    net.load_state_dict(pn)
    with torch.no_grad():
        a, b = net(x, uv)
    c, d = O.ae_forward(pn, x, uv, HS, 2, True)
    assert max_abs(a, c) == 0 and max_abs(b, d) == 0
    km = R.Kmeans(40, 7, 3)
    X = torch.randn(9, 40)
    assert abs(float(km(X)) - float(O.khm_loss(X, km.M.detach(), 3))) < 1e-5 * float(km(X))
