"""Kernel-level parity of liblshm_sm100 against PyTorch CPU fp32 (the arithmetic the reference
itself calls): convs / transposed convs and their gradients, linear layers, K-harmonic family,
small losses.  All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import max_abs, rel_err
from lshm_b200._lib import lib
from oracle import lofar_oracle as O

pytestmark = pytest.mark.gpu
CH = (8, 12, 24, 48, 96, 192)
# tensor-core convs: fp32 operands split into bf16 hi+lo, 3 MMAs, fp32 accumulate:
# ~3e-6 relative per output; the parity contract (north_star) is 1e-3 on losses / gradients
TC_TOL = 2e-5


def st():
    return torch.cuda.current_stream().cuda_stream


def dp(t):
    return None if t is None else t.data_ptr()


def image(w, dim, which=0):
    from lshm_b200.engine import conv_image
    return conv_image(w, dim, which, st())


def elu_grad_from_out(a):
    return torch.where(a > 0, torch.ones_like(a), a + 1)


# ---------------------------------------------------------------------------- convs 2-D
@pytest.mark.parametrize("lvl", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("C", [8, 4])
def test_conv2d_family(cuda, lvl, C):
    torch.manual_seed(lvl)
    ch = (C,) + CH
    A, Bc, s = ch[lvl], ch[lvl - 1], 128 >> lvl
    N = 3
    big = torch.randn(N, Bc, 2 * s, 2 * s)
    w = torch.randn(A, Bc, 4, 4) * 0.1
    bias = torch.randn(A)
    # down = Conv2d forward (+ELU)
    ref = F.elu(F.conv2d(big, w, bias, stride=2, padding=1))
    out = torch.empty(N, A, s, s, device=cuda)
    bg, wg, bsg = big.to(cuda), w.to(cuda), bias.to(cuda)
    wdn = image(wg, 2)
    lib().down2d(dp(bg), Bc * 4 * s * s, dp(wdn), dp(bsg), None, 0, dp(out), A * s * s, N, A, Bc, s, s, 1, st())
    assert rel_err(out, ref) < TC_TOL
    # strided destination (conv5 writes into the concat buffer)
    pad = torch.zeros(N, A * s * s + 16, device=cuda)
    lib().down2d(dp(bg), Bc * 4 * s * s, dp(wdn), dp(bsg), None, 0, dp(pad), A * s * s + 16, N, A, Bc, s, s, 1, st())
    assert rel_err(pad[:, :A * s * s].reshape(N, A, s, s), ref) < TC_TOL and float(pad[:, A * s * s:].abs().max()) == 0
    # up = ConvTranspose2d forward with weight [A,Bc,4,4] (+ELU) and = Conv2d dgrad
    small = torch.randn(N, A, s, s)
    bias_b = torch.randn(Bc)
    ref_up = F.elu(F.conv_transpose2d(small, w, bias_b, stride=2, padding=1))
    sg = small.to(cuda)
    out_up = torch.empty(N, Bc, 2 * s, 2 * s, device=cuda)
    bbg = bias_b.to(cuda)
    wup = image(wg, 2, 1)
    lib().up2d(dp(sg), A * s * s, dp(wup), dp(bbg), None, 0, dp(out_up), Bc * 4 * s * s, N, A, Bc, s, s, 1, st())
    assert rel_err(out_up, ref_up) < TC_TOL
    # up with DELU epilogue = dgrad through the ELU of the layer below
    act = F.elu(torch.randn(N, Bc, 2 * s, 2 * s))
    ref_d = F.conv_transpose2d(small, w, None, stride=2, padding=1) * elu_grad_from_out(act)
    ag = act.to(cuda)
    lib().up2d(dp(sg), A * s * s, dp(wup), None, dp(ag), Bc * 4 * s * s, dp(out_up), Bc * 4 * s * s, N, A, Bc, s, s, 2, st())
    assert rel_err(out_up, ref_d) < TC_TOL
    # down with DELU and no bias = ConvTranspose2d dgrad
    act_s = F.elu(torch.randn(N, A, s, s))
    ref_dd = F.conv2d(big, w, None, stride=2, padding=1) * elu_grad_from_out(act_s)
    asg = act_s.to(cuda)
    lib().down2d(dp(bg), Bc * 4 * s * s, dp(wdn), None, dp(asg), A * s * s, dp(out), A * s * s, N, A, Bc, s, s, 2, st())
    assert rel_err(out, ref_dd) < TC_TOL
    # wgrad
    bigr = big.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    F.conv2d(bigr, wr, None, stride=2, padding=1).backward(small)
    dw = torch.empty(A, Bc, 4, 4, device=cuda)
    lib().wgrad2d(dp(sg), A * s * s, dp(bg), Bc * 4 * s * s, dp(dw), N, A, Bc, s, s, st())
    assert rel_err(dw, wr.grad) < 1e-5
    db = torch.empty(A, device=cuda)
    lib().channel_sum(dp(sg), A * s * s, dp(db), N, A, s * s, st())
    assert rel_err(db, small.sum(dim=(0, 2, 3))) < 1e-5


# ---------------------------------------------------------------------------- convs 1-D
@pytest.mark.parametrize("lvl", [1, 2, 3, 4, 5, 6])
def test_conv1d_family(cuda, lvl):
    torch.manual_seed(10 + lvl)
    ch = (8,) + CH
    A, Bc, l = ch[lvl], ch[lvl - 1], 16384 >> (2 * lvl)
    N = 2
    wg_ = torch.randn(A, Bc, 4) * 0.1
    big = torch.randn(N, Bc, 4 * l)
    small = torch.randn(N, A, l)
    bg, sg, wg = big.to(cuda), small.to(cuda), wg_.to(cuda)
    bias_a, bias_b = torch.randn(A), torch.randn(Bc)
    out_s = torch.empty(N, A, l, device=cuda)
    out_b = torch.empty(N, Bc, 4 * l, device=cuda)
    # Conv1d(k4,s4,p1) forward
    bag, bbg = bias_a.to(cuda), bias_b.to(cuda)
    wdn = image(wg, 1)
    lib().down1d(dp(bg), Bc * 4 * l, dp(wdn), dp(bag), None, 0, dp(out_s), A * l, N, A, Bc, l, 1, 1, st())
    assert rel_err(out_s, F.elu(F.conv1d(big, wg_, bias_a, stride=4, padding=1))) < TC_TOL
    # ConvTranspose1d(k4,s4,p0) forward
    wup = image(wg, 1, 1)
    lib().up1d(dp(sg), A * l, dp(wup), dp(bbg), None, 0, dp(out_b), Bc * 4 * l, N, A, Bc, l, 0, 1, st())
    assert rel_err(out_b, F.elu(F.conv_transpose1d(small, wg_, bias_b, stride=4, padding=0))) < TC_TOL
    # Conv1d dgrad (pad 1) with DELU
    act = F.elu(torch.randn(N, Bc, 4 * l))
    bigr = big.clone().requires_grad_()
    wr = wg_.clone().requires_grad_()
    F.conv1d(bigr, wr, None, stride=4, padding=1).backward(small)
    actg = act.to(cuda)
    lib().up1d(dp(sg), A * l, dp(wup), None, dp(actg), Bc * 4 * l, dp(out_b), Bc * 4 * l, N, A, Bc, l, 1, 2, st())
    assert rel_err(out_b, bigr.grad * elu_grad_from_out(act)) < TC_TOL
    dw = torch.empty(A, Bc, 4, device=cuda)
    lib().wgrad1d(dp(sg), A * l, dp(bg), Bc * 4 * l, dp(dw), N, A, Bc, l, 1, st())
    assert rel_err(dw, wr.grad) < 1e-5
    # ConvTranspose1d dgrad (pad 0) and wgrad
    smr = small.clone().requires_grad_()
    wr2 = wg_.clone().requires_grad_()
    F.conv_transpose1d(smr, wr2, None, stride=4, padding=0).backward(big)
    lib().down1d(dp(bg), Bc * 4 * l, dp(wdn), None, None, 0, dp(out_s), A * l, N, A, Bc, l, 0, 0, st())
    assert rel_err(out_s, smr.grad) < TC_TOL
    lib().wgrad1d(dp(sg), A * l, dp(bg), Bc * 4 * l, dp(dw), N, A, Bc, l, 0, st())
    assert rel_err(dw, wr2.grad) < 1e-5


# ---------------------------------------------------------------------------- linear
@pytest.mark.parametrize("N,K,J", [(5, 784, 32), (70, 16, 16), (33, 48, 768), (1, 32, 32)])
def test_linear(cuda, N, K, J):
    torch.manual_seed(N)
    x, w, b = torch.randn(N, K + 3)[:, :K], torch.randn(J, K) * 0.1, torch.randn(J)
    xg = torch.zeros(N, K + 3, device=cuda)
    xg[:, :K] = x.to(cuda)
    wg, bgp = w.to(cuda), b.to(cuda)
    y = torch.empty(N, J + 5, device=cuda)
    lib().linear_fwd(dp(xg), K + 3, dp(wg), dp(bgp), dp(y), J + 5, N, K, J, 1, st())
    ref = F.elu(F.linear(x, w, b))
    assert rel_err(y[:, :J], ref) < 2e-6
    dz = torch.randn(N, J)
    add = torch.randn(N, K)
    dzg = dz.to(cuda)
    dx = torch.empty(N, K, device=cuda)
    aux = F.elu(torch.randn(N, K))
    addg, auxg = add.to(cuda), aux.to(cuda)
    lib().linear_bwd_data(dp(dzg), J, dp(wg), dp(addg), K, dp(auxg), K, dp(dx), K, N, K, J, st())
    assert rel_err(dx, (dz @ w + add) * elu_grad_from_out(aux)) < 2e-6
    dw, db = torch.empty(J, K, device=cuda), torch.empty(J, device=cuda)
    lib().linear_bwd_weight(dp(xg), K + 3, dp(dzg), J, dp(dw), dp(db), N, K, J, st())
    assert rel_err(dw, dz.t() @ x) < 1e-5 and rel_err(db, dz.sum(0)) < 1e-5


def test_uv_harmonics(cuda):
    uv = torch.randn(9, 2) * 300
    sc = torch.tensor([1e-4, 1e-3, 1e-2, 1e-1])
    out = torch.empty(9, 16, device=cuda)
    uvg, scg = uv.to(cuda), sc.to(cuda)
    lib().uv_harmonics(dp(uvg), dp(scg), 9, 4, dp(out), st())
    assert max_abs(out, O.uv_harmonics(uv, sc)) < 5e-6


# ---------------------------------------------------------------------------- K-harmonic
KHM_CASES = [(37, 10, 64, 4.0), (1000, 10, 64, 4.0), (129, 7, 32, 2.0), (200, 64, 128, 4.0),
             (64, 300, 256, 4.0), (50, 10, 256, 3.0), (90, 1024, 32, 4.0), (31, 5, 48, 4.0)]


@pytest.mark.parametrize("N,K,L,p", KHM_CASES)
def test_khm_family(cuda, N, K, L, p):
    rng = np.random.default_rng(N + K)
    X = torch.from_numpy(rng.standard_normal((N, L)).astype(np.float32))
    M = O.make_centres(K, L, seed=K)
    if N > 5:
        X[3] = M[min(2, K - 1)]  # exactly on a centre
    Xg, Mg = X.to(cuda), M.to(cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    e = torch.empty(N, device=cuda)
    lib().khm_fwd(dp(Xg), L, dp(Mg), N, K, L, p, dp(acc), dp(e), st())
    ref = float(O.khm_loss(X, M, p))
    assert abs(float(acc) / (N * K * L) - ref) <= 2e-5 * abs(ref)
    # gradients vs float64 analytic oracle (which matches autograd, test_oracle_golden)
    gx_ref, gm_ref = O.khm_grads_analytic(X, M, p)
    gX = torch.empty(N, L, device=cuda)
    gM = torch.zeros(K, L, device=cuda)
    acc2 = torch.zeros(1, dtype=torch.float64, device=cuda)
    lib().khm_fwd_bwd(dp(Xg), L, dp(Mg), N, K, L, p, 1.0 / (N * K * L), dp(acc2), dp(gX), L, 0, dp(gM), st())
    assert abs(float(acc2) - float(acc)) <= 1e-6 * abs(float(acc))   # two kernels, two fp32 summation orders
    assert rel_err(gX, gx_ref) < 1e-4 and rel_err(gM, gm_ref) < 1e-4
    gM2 = torch.zeros(K, L, device=cuda)
    gX2 = torch.ones(N, L, device=cuda)
    lib().khm_bwd(dp(Xg), L, dp(Mg), N, K, L, p, 1.0 / (N * K * L), dp(gX2), L, 1, dp(gM2), st())
    assert rel_err(gX2 - 1, gx_ref) < 1e-3 and rel_err(gM2, gm_ref) < 1e-4
    # assignment
    ids = torch.empty(N, dtype=torch.int32, device=cuda)
    lib().khm_assign(dp(Xg), L, dp(Mg), N, K, L, dp(ids), st())
    d = torch.cdist(X.double(), M.double())
    ref_ids = d.argmin(dim=1)
    agree = (ids.cpu().long() == ref_ids).double().mean().item()
    assert agree >= 0.999 or (ids.cpu().long() != ref_ids).sum() <= 1
    # centre update sums
    Mn, num_ref, den_ref = O.offline_update(X, M, p)
    num = torch.zeros(K, L, device=cuda)
    den = torch.zeros(K, device=cuda)
    lib().khm_center_sums(dp(Xg), L, dp(Mg), N, K, L, p, dp(num), dp(den), st())
    assert rel_err(num, num_ref) < 1e-3 and rel_err(den, den_ref) < 1e-3
    Mnew = torch.empty(K, L, device=cuda)
    lib().khm_center_apply(dp(num), dp(den), dp(Mnew), K, L, st())
    assert rel_err(Mnew, Mn) < 1e-3


def test_khm_group_dist_and_strides(cuda):
    N, K, L, p, grp = 48, 10, 64, 4.0, 6
    rng = np.random.default_rng(1)
    buf = torch.from_numpy(rng.standard_normal((N, L + 8)).astype(np.float32))
    X = buf[:, 4:4 + L].contiguous()
    M = O.make_centres(K, L, seed=2)
    bg = buf.to(cuda)
    dist = torch.empty(N // grp, K, device=cuda)
    gid = torch.empty(N // grp, dtype=torch.int32, device=cuda)
    Mg = M.to(cuda)
    lib().khm_group_dist(bg.data_ptr() + 16, L + 8, dp(Mg), N, K, L, p, grp, dp(dist), dp(gid), st())
    for g in range(N // grp):
        d, idx, _ = O.eval_distances(X[g * grp:(g + 1) * grp], M, p)
        assert rel_err(dist[g], d) < 1e-5 and int(gid[g]) == idx


def test_khm_rejects_bad_arguments(cuda):
    from lshm_b200._lib import LshmError
    X = torch.zeros(4, 30, device=cuda)
    M = torch.zeros(2, 30, device=cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    with pytest.raises(LshmError):
        lib().khm_fwd(dp(X), 30, dp(M), 4, 2, 30, 2.0, dp(acc), None, st())   # L % 4 != 0
    # empty input is a no-op, not an error
    lib().khm_fwd(dp(X), 32, dp(M), 0, 2, 32, 2.0, dp(acc), None, st())
    assert float(acc) == 0.0


# ---------------------------------------------------------------------------- small losses
@pytest.mark.parametrize("K,L", [(10, 64), (3, 256), (40, 32)])
def test_similarity(cuda, K, L):
    M = O.make_centres(K, L, seed=K)
    Mr = M.clone().requires_grad_()
    ref = O.cluster_similarity(Mr)
    ref.backward()
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    gM = torch.zeros(K, L, device=cuda)
    work = torch.empty(2 * K * K, device=cuda)
    Mg = M.to(cuda)
    lib().similarity(dp(Mg), K, L, 0.5, dp(acc), dp(gM), dp(work), st())
    assert abs(float(acc) - 0.5 * float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(gM, 0.5 * Mr.grad) < 1e-4


@pytest.mark.parametrize("bpb,groups,L", [(4, 8, 64), (6, 3, 64), (16, 2, 256), (1, 5, 32)])
def test_augment(cuda, bpb, groups, L):
    torch.manual_seed(bpb)
    mu = torch.randn(bpb * groups, L)
    mur = mu.clone().requires_grad_()
    ref = O.augmented_loss(mur, bpb, groups).sum()
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    g = torch.zeros(bpb * groups, L, device=cuda)
    scale = 1.0 / (bpb * groups * bpb)
    mug = mu.to(cuda)
    lib().augment(dp(mug), L, bpb * groups, L, bpb, scale, dp(acc), dp(g), L, st())
    assert abs(float(acc) - float(ref)) <= 1e-5 * max(abs(float(ref)), 1e-12)
    if bpb > 1:
        ref.backward()
        assert rel_err(g, mur.grad) < 1e-4


def test_logcosh_and_adam(cuda):
    torch.manual_seed(0)
    mu = torch.randn(33, 48) * 2
    mur = mu.clone().requires_grad_()
    ref = 0.01 * torch.log(torch.cosh(mur)).sum() / mur.numel()
    ref.backward()
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    g = torch.zeros(33, 48, device=cuda)
    mug = mu.to(cuda)
    lib().logcosh(dp(mug), 48, 33, 48, 0.01, dp(acc), dp(g), 48, st())
    assert abs(float(acc) - float(ref)) < 1e-5 * abs(float(ref)) and rel_err(g, mur.grad) < 1e-5
    # Adam: three steps against torch.optim.Adam
    p = torch.randn(1000)
    pr = p.clone().requires_grad_()
    opt = torch.optim.Adam([pr], lr=1e-3)
    pg, m, v = p.to(cuda), torch.zeros(1000, device=cuda), torch.zeros(1000, device=cuda)
    for t in range(1, 4):
        gr = torch.randn(1000)
        pr.grad = gr.clone()
        opt.step()
        grg = gr.to(cuda)
        lib().adam_step(dp(pg), dp(grg), dp(m), dp(v), 1000, 1e-3, 0.9, 0.999, 1e-8, t, st())
    assert max_abs(pg, pr.detach()) < 1e-6


# ---------------------------------------------------------------------------- cascade glue
def test_cascade_kernels(cuda):
    torch.manual_seed(3)
    N, C, P, rho = 3, 4, 128, 0.7
    x, x1, x2, x3 = (torch.randn(N, C, P, P) for _ in range(4))
    y1, y2, y3 = (torch.randn(N * C * P * P) for _ in range(3))
    x3f = x3.transpose(2, 3).contiguous()
    g = [t.to(cuda) for t in (x, x1, x2, x3f, y1, y2, y3)]
    iyT, iyF = torch.empty_like(g[0]), torch.empty_like(g[0])
    lib().residual_split(dp(g[0]), dp(g[1]), dp(iyT), dp(iyF), N, C, P, st())
    x11 = (x - x1) / 2
    assert max_abs(iyT, x11) == 0 and max_abs(iyF, x11.transpose(2, 3)) == 0
    # losses + gradients against autograd
    a1, a2, a3 = (t.clone().requires_grad_() for t in (x1, x2, x3))
    n = x.numel()
    x11r = (x - a1) / 2
    sse = lambda a, b: ((a - b) ** 2).sum()
    l0 = sse(a1 + a2 + a3, x) / n
    l1 = (torch.dot(y1, (x - a1).reshape(-1)) + rho / 2 * sse(x, a1)) / n
    l2 = (torch.dot(y2, (x11r - a2).reshape(-1)) + rho / 2 * sse(x11r, a2)) / n
    l3 = (torch.dot(y3, (x11r - a3).reshape(-1)) + rho / 2 * sse(x11r, a3)) / n
    (l0 + l1 + l2 + l3).backward()
    sums = torch.zeros(8, dtype=torch.float64, device=cuda)
    g1p, g2, g3f = (torch.empty_like(g[0]) for _ in range(3))
    lib().cascade_losses(*(dp(t) for t in g), rho, N, C, P, 1.0 / n, dp(sums), dp(g1p), dp(g2), dp(g3f), None, None, st())
    s = sums.cpu()
    assert abs(float(s[0]) / n - float(l0)) < 1e-6 * float(l0)
    for (i, l) in ((1, l1), (3, l2), (5, l3)):
        assert abs((float(s[i]) + rho / 2 * float(s[i + 1])) / n - float(l)) < 1e-5 * abs(float(l)) + 1e-9
    assert rel_err(g2, a2.grad) < 1e-5 and rel_err(g3f, a3.grad.transpose(2, 3)) < 1e-5
    assert rel_err(g1p, a1.grad) < 1e-5  # with no 1-D input gradients g1p is the whole d/dx1
    # forward-only variant gives the same sums
    sums2 = torch.zeros(8, dtype=torch.float64, device=cuda)
    lib().cascade_losses(*(dp(t) for t in g), rho, N, C, P, 1.0 / n, dp(sums2), None, None, None, None, None, st())
    assert torch.allclose(sums2, sums, rtol=1e-12)
    # fused bias-gradient sums (per-channel sums of g2 / g3f), same sums and gradients otherwise
    sums3 = torch.zeros(8, dtype=torch.float64, device=cuda)
    h1, h2, h3 = (torch.empty_like(g[0]) for _ in range(3))
    db2, db3 = torch.full((C,), 9.0, device=cuda), torch.full((C,), 9.0, device=cuda)
    lib().cascade_losses(*(dp(t) for t in g), rho, N, C, P, 1.0 / n, dp(sums3), dp(h1), dp(h2), dp(h3), dp(db2), dp(db3), st())
    assert torch.allclose(sums3, sums, rtol=1e-12) and torch.equal(h2, g2) and torch.equal(h3, g3f) and torch.equal(h1, g1p)
    assert rel_err(db2, g2.sum(dim=(0, 2, 3))) < 1e-5 and rel_err(db3, g3f.sum(dim=(0, 2, 3))) < 1e-5
    gT, gF = torch.randn(N, C, P, P), torch.randn(N, C, P, P)
    gx1 = torch.empty_like(g[0])
    gTg, gFg = gT.to(cuda), gF.to(cuda)
    lib().cascade_combine(dp(g1p), dp(gTg), dp(gFg), dp(gx1), N, C, P, None, st())
    assert rel_err(gx1, g1p.cpu() - 0.5 * (gT + gF.transpose(2, 3))) < 1e-6
    gx1b, db1 = torch.empty_like(g[0]), torch.full((C,), 9.0, device=cuda)
    lib().cascade_combine(dp(g1p), dp(gTg), dp(gFg), dp(gx1b), N, C, P, dp(db1), st())
    assert torch.equal(gx1b, gx1) and rel_err(db1, gx1.sum(dim=(0, 2, 3))) < 1e-5
    lib().multiplier_update(dp(g[0]), dp(g[1]), dp(g[2]), dp(g[3]), rho, dp(g[4]), dp(g[5]), dp(g[6]), N, C, P, st())
    assert rel_err(g[4], y1 + rho * (x - x1).reshape(-1)) < 1e-6
    assert rel_err(g[5], y2 + rho * (x11 - x2).reshape(-1)) < 1e-6
    assert rel_err(g[6], y3 + rho * (x11 - x3).reshape(-1)) < 1e-6


@pytest.mark.parametrize("Cn,ln,pad", [(192, 4, 0), (96, 16, 0), (24, 256, 0), (12, 1024, 0), (8, 16384, 0),
                                        (48, 64, 8), (5, 37, 0), (8, 4100, 4), (3, 6, 2)])
def test_channel_sum_all_row_lengths_and_strides(cuda, Cn, ln, pad):
    """Bias-gradient reduction: long rows (persistent chunks), short contiguous rows (flat kernel), and the
    strided / unaligned / odd-length fallbacks."""
    torch.manual_seed(Cn * 7 + ln)
    N = 37
    buf = torch.randn(N, Cn * ln + pad, device=cuda)
    g = buf[:, :Cn * ln]
    db = torch.full((Cn,), 7.0, device=cuda)          # written, not accumulated
    lib().channel_sum(g.data_ptr(), buf.stride(0), dp(db), N, Cn, ln, st())
    ref = g.reshape(N, Cn, ln).double().sum(dim=(0, 2)).float()
    assert rel_err(db, ref) < 1e-5
