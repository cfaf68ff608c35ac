"""Generates tests/golden/*.npz by running the LIVE reference (/root/reference) on seeded
synthetic inputs.  TEST INFRASTRUCTURE ONLY; run in the authoring container:

    python -m oracle.gen_golden

The reference ships no golden vectors, so these fixtures are what pins the
oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py).  Inputs are
regenerated from numpy seeds (lshm_b200.synthetic, oracle.make_ae_params), only outputs are
stored: latents, loss terms, per-parameter gradient summaries, strided sub-samples.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SCALES = [1e-4, 1e-3, 1e-2, 1e-1]  # src/kharmonic_lofar.py:57


def sub2(t):  # strided sub-sample of a [N,C,128,128] tensor
    return t[:, :, ::16, ::16].contiguous().numpy()


def grad_summary(params):
    out = {}
    for k, v in params.items():
        g = v.grad.detach()
        out[k + ":norm"] = np.float64(g.double().norm().item())
        out[k + ":head"] = g.reshape(-1)[:4].numpy().copy()
    return out


def main():
    sys.path.insert(0, REF)
    from oracle import fake_h5
    fake_h5.install()
    import lofar_models as R
    import lofar_tools as RT
    from oracle import lofar_oracle as O
    from lshm_b200 import synthetic as S
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    hs = torch.tensor(SCALES)

    # ---------------- autoencoders forward (C=8 L=32 rica ; C=4 L=24 no-rica for the 2-D net)
    for tag, C, L, Lt, rica in (("c8", 8, 32, 16, True), ("c4", 4, 24, 8, True)):
        pn = O.make_ae_params(L, C, ndim=2, seed=11)
        pT = O.make_ae_params(Lt, C, ndim=1, seed=12)
        x = torch.from_numpy(S.make_patches(2, C, seed=21))
        uv = torch.from_numpy(S.make_uv(2, seed=21))
        net = R.AutoEncoderCNN2(L, C, hs, rica); net.load_state_dict(pn)
        nT = R.AutoEncoder1DCNN(Lt, C, hs, rica); nT.load_state_dict(pT)
        with torch.no_grad():
            xh, mu = net(x, uv)
            yT, muT = nT(torch.flatten(x, 2, 3), uv)
        np.savez(os.path.join(OUT, f"ae_forward_{tag}.npz"), C=C, L=L, Lt=Lt, rica=rica,
                 xhat_sub=sub2(xh), xhat_sum=np.float64(xh.double().sum().item()), mu=mu.numpy(),
                 yT_sub=sub2(yT.view(2, C, 128, 128)), yT_sum=np.float64(yT.double().sum().item()), muT=muT.numpy())

    # ---------------- Kmeans forward / similarity (literal loops of the reference)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((12, 64)).astype(np.float32)
    M = O.make_centres(10, 64, seed=6)
    X[3] = M[2].numpy()  # a point sitting exactly on a centre (d=0 edge case)
    rec = dict(X=X, M=M.numpy())
    for p in (2, 4):
        km = R.Kmeans(64, 10, p); km.load_state_dict({"M": M})
        Xt = torch.from_numpy(X).requires_grad_()
        loss = km(Xt)
        loss.backward()
        rec[f"loss_p{p}"] = np.float64(loss.item())
        rec[f"gX_p{p}"] = Xt.grad.numpy().copy()
        rec[f"gM_p{p}"] = km.M.grad.numpy().copy()
        km.zero_grad()
        sim = km.cluster_similarity()
        sim.backward()
        rec["sim"] = np.float64(sim.item())
        rec["gsim"] = km.M.grad.numpy().copy()
    np.savez(os.path.join(OUT, "kmeans.npz"), **rec)

    # ---------------- full closure, cfg1-like (reference modules + restated closure lines)
    C, L, Lt, K, Khp, N, bpb = 8, 32, 16, 10, 4, 8, 4
    alpha = beta = gamma = 0.01; rho = 1.0; lam = 0.01
    pn = O.make_ae_params(L, C, ndim=2, seed=1); pT = O.make_ae_params(Lt, C, ndim=1, seed=2)
    pF = O.make_ae_params(Lt, C, ndim=1, seed=3); M = O.make_centres(K, L + 2 * Lt, seed=4)
    net = R.AutoEncoderCNN2(L, C, hs, True); net.load_state_dict(pn)
    netT = R.AutoEncoder1DCNN(Lt, C, hs, True); netT.load_state_dict(pT)
    netF = R.AutoEncoder1DCNN(Lt, C, hs, True); netF.load_state_dict(pF)
    mod = R.Kmeans(L + 2 * Lt, K, Khp); mod.load_state_dict({"M": M})
    x = torch.from_numpy(S.make_patches(N, C, seed=5)); uv = torch.from_numpy(S.make_uv(N, seed=5, per_group=bpb))
    g = torch.Generator().manual_seed(1)
    y1, y2, y3 = (0.1 * torch.randn(x.numel(), generator=g) for _ in range(3))
    crit = torch.nn.MSELoss(reduction="sum")

    def ref_aug(mu, b, bs):  # literal src/kharmonic_lofar.py:97-110
        loss = torch.zeros(1)
        for ck in range(bs):
            Z = mu[ck * b:(ck + 1) * b, :]
            prod = torch.zeros(1)
            for ci in range(b):
                zi = Z[ci, :] / (torch.norm(Z[ci, :]) + 1e-6)
                for cj in range(ci + 1, b):
                    zj = Z[cj, :] / (torch.norm(Z[cj, :]) + 1e-6)
                    prod = prod + torch.exp(-torch.dot(zi, zj))
            loss = loss + prod / b
        return loss / (bs * b)

    # src/kharmonic_lofar.py:135-172
    x1, mu = net(x, uv)
    x11 = (x - x1) / 2
    yyT, yyTmu = netT(torch.flatten(x11, 2, 3), uv)
    x2 = yyT.view_as(x11)
    yyF, yyFmu = netF(torch.flatten(torch.transpose(x11, 2, 3), 2, 3), uv)
    x3 = torch.transpose(yyF.view_as(x11), 2, 3)
    xrecon = x1 + x2 + x3
    n = x.numel()
    loss0 = crit(xrecon, x) / n
    loss1 = (torch.dot(y1, (x - x1).reshape(-1)) + rho / 2 * crit(x, x1)) / n
    loss2 = (torch.dot(y2, (x11 - x2).reshape(-1)) + rho / 2 * crit(x11, x2)) / n
    loss3 = (torch.dot(y3, (x11 - x3).reshape(-1)) + rho / 2 * crit(x11, x3)) / n
    Mu = torch.cat((mu, yyTmu, yyFmu), 1)
    kdist = alpha * mod.clustering_error(Mu)
    sim = beta * mod.cluster_similarity()
    aug = gamma * ref_aug(Mu, bpb, N // bpb)
    rl = lam * (torch.sum(torch.log(torch.cosh(mu))) / mu.numel() + torch.sum(torch.log(torch.cosh(yyTmu))) / yyTmu.numel()
                + torch.sum(torch.log(torch.cosh(yyFmu))) / yyFmu.numel())
    loss = loss0 + loss1 + loss2 + loss3 + kdist + aug + sim + rl
    loss.backward()
    rec = dict(total=loss.item(), loss0=loss0.item(), loss1=loss1.item(), loss2=loss2.item(), loss3=loss3.item(),
               kdist=kdist.item(), aug=aug.item(), sim=sim.item(), rica=rl.item(), Mu=Mu.detach().numpy(),
               xrecon_sub=sub2(xrecon.detach()))
    for tag, m in (("n", net), ("T", netT), ("F", netF), ("k", mod)):
        for k, v in grad_summary(dict(m.named_parameters())).items():
            rec[f"{tag}.{k}"] = v
    # multiplier update, src/kharmonic_lofar.py:200-202
    with torch.no_grad():
        rec["y1_head"] = (y1 + rho * (x - x1).reshape(-1))[:8].numpy()
        rec["y2_sum"] = np.float64((y2 + rho * (x11 - x2).reshape(-1)).double().sum().item())
        rec["y3_sum"] = np.float64((y3 + rho * (x11 - x3).reshape(-1)).double().sum().item())
        # eval distances, src/evaluate_clustering.py:110-117 (first baseline group)
        Mg = Mu[:bpb]
        dist = torch.zeros(K)
        for ck in range(K):
            for cn in range(bpb):
                dist[ck] = dist[ck] + torch.sum(torch.pow(torch.linalg.norm(Mg[cn, :] - mod.M[ck, :], 2), Khp))
        dist = dist / bpb
        rec["eval_dist"] = dist.numpy()
        rec["eval_id"] = int(torch.min(dist.view(K, 1), 0)[1][0])
    np.savez(os.path.join(OUT, "closure_cfg1.npz"), **rec)

    # ---------------- loader (unmodified reference get_data_minibatch on the fake h5)
    meas = S.make_measurement(5, 256, 192, seed=0)
    fake_h5.register("golden0", meas)
    rec = {}
    for tag, C, norm in (("c8n", 8, True), ("c8", 8, False), ("c4n", 4, True)):
        np.random.seed(3)
        px, py, y, uv1 = RT.get_data_minibatch(["golden0"], ["0"], batch_size=3, patch_size=128, normalize_data=norm,
                                               num_channels=C, uvdist=True)
        rec[f"{tag}:pxpy"] = np.array([px, py])
        rec[f"{tag}:sub"] = y[:, :, ::16, ::16].contiguous().numpy()
        rec[f"{tag}:sum"] = np.float64(y.double().sum().item())
        rec[f"{tag}:abs"] = np.float64(y.double().abs().sum().item())
        rec[f"{tag}:uv"] = uv1.numpy()
    px, py, y, uv1 = RT.get_data_for_baseline("golden0", "0", baseline_id=2, patch_size=128, num_channels=8, uvdist=True)
    rec["base2:sub"] = y[:, :, ::16, ::16].contiguous().numpy()
    rec["base2:uv"] = uv1.numpy()
    # small observation: zero padding to 128 (src/lofar_tools.py:86)
    meas2 = S.make_measurement(3, 100, 150, seed=1)
    fake_h5.register("golden1", meas2)
    np.random.seed(4)
    px, py, y = RT.get_data_minibatch(["golden1"], ["0"], batch_size=2, patch_size=128, normalize_data=False, num_channels=8)
    rec["pad:pxpy"] = np.array([px, py])
    rec["pad:sub"] = y[:, :, ::8, ::8].contiguous().numpy()
    rec["pad:sum"] = np.float64(y.double().sum().item())
    np.savez(os.path.join(OUT, "loader.npz"), **rec)

    # ---------------- Fourier features (Demo.ipynb:169-174 with the reference torch_fftshift)
    x = torch.from_numpy(S.make_patches(2, 4, seed=31))
    xhat = 0.3 * torch.from_numpy(S.make_patches(2, 4, seed=32))
    fftx = torch.fft.fftn(x - xhat, dim=(2, 3), norm="ortho")
    fr, fi = RT.torch_fftshift(fftx.real, fftx.imag)
    yf = torch.cat((fr, fi), 1)
    yf.clamp_(min=-10, max=10)
    np.savez(os.path.join(OUT, "fft.npz"), sub=yf[:, :, ::8, ::8].contiguous().numpy(),
             centre=yf[:, :, 60:68, 60:68].contiguous().numpy(), sum=np.float64(yf.double().sum().item()),
             abs=np.float64(yf.double().abs().sum().item()))
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(" ", fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
