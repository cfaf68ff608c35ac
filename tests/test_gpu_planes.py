"""Operand planes (include/lshm.h "operand planes", csrc/tma.cuh): the staging kernels against a CPU restatement of
the layout, the tensor-TMA instances of the first-layer conv kernels against the fp32-input instances (same bf16
hi/lo values, same MMA order: identical results) and against PyTorch, and the fused writers (residual split,
gradient combine) against staging their fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

from common import rel_err
from lshm_b200._lib import lib

pytestmark = pytest.mark.gpu
TC_TOL = 2e-5


def st():
    return torch.cuda.current_stream().cuda_stream


def dp(t):
    return None if t is None else t.data_ptr()


def image(w, dim, which=0):
    from lshm_b200.engine import conv_image
    return conv_image(w, dim, which, st())


def planes_buffer(dim, N, Bc, h, w, dev):
    import ctypes
    n = ctypes.c_int64()
    assert lib().cdll.lshm_planes_bytes(dim, N, Bc, h, w, ctypes.byref(n)) == 0
    # zero-filled: the <= 31 padding positions per chunk (chunk stride = positions rounded up to 32) are never written
    return torch.zeros(n.value // 2, dtype=torch.bfloat16, device=dev)


def pad_positions(z):
    """[cc, Q, 8] -> [cc, Qs, 8] with Qs = Q rounded up to 32 (zero padding)."""
    cc, Q, _ = z.shape
    Qs = (Q + 31) // 32 * 32
    out = torch.zeros(cc, Qs, 8, dtype=z.dtype)
    out[:, :Q] = z
    return out


def split_ref(v):
    hi = v.bfloat16()
    lo = (v - hi.float()).bfloat16()
    return hi, lo


def planes_ref_2d(big):
    """[half][chunk][q][8] from big [N,Bc,2h,2w]: q over the (h+1)x(w+1) block grid, block (by,bx) = rows 2by-1,2by x cols
    2bx-1,2bx (zero outside), chunk = channel pair, element = (b&1)*4 + sy*2 + sx."""
    N, Bc, H, W = big.shape
    pad = F.pad(big, (1, 1, 1, 1))                       # pixel (r, c) at [r+1, c+1]
    blocks = pad.unfold(2, 2, 2).unfold(3, 2, 2)         # [N,Bc,h+1,w+1,2(sy),2(sx)]
    z = blocks.reshape(N, Bc // 2, 2, H // 2 + 1, W // 2 + 1, 4)      # [N,cc,bb,by,bx,sub]
    z = z.permute(1, 0, 3, 4, 2, 5).reshape(Bc // 2, -1, 8)           # [cc, q, 8]
    hi, lo = split_ref(pad_positions(z.contiguous()))
    return torch.stack((hi, lo))


def planes_ref_1d(big, pad):
    N, Bc, Lb = big.shape
    src = F.pad(big, (pad, 0))[:, :, :Lb]                # sample s at index s + pad  -> window j = [4j, 4j+3] of src
    z = src.reshape(N, Bc // 2, 2, Lb // 4, 4).permute(1, 0, 3, 2, 4).reshape(Bc // 2, -1, 8)
    hi, lo = split_ref(pad_positions(z.contiguous()))
    return torch.stack((hi, lo))


@pytest.mark.parametrize("N,Bc,s", [(3, 8, 64), (2, 4, 16), (2, 12, 8)])
def test_stage_planes2d_layout(cuda, N, Bc, s):
    torch.manual_seed(N + Bc)
    big = torch.randn(N, Bc, 2 * s, 2 * s)
    bg = big.to(cuda)
    pl = planes_buffer(2, N, Bc, s, s, cuda)
    lib().stage_planes2d(dp(bg), Bc * 4 * s * s, dp(pl), N, Bc, s, s, st())
    ref = planes_ref_2d(big)
    assert torch.equal(pl.cpu().view(torch.int16), ref.reshape(-1).view(torch.int16))


@pytest.mark.parametrize("N,Bc,l,pad", [(3, 8, 4096, 1), (2, 8, 4096, 0), (2, 4, 64, 1)])
def test_stage_planes1d_layout(cuda, N, Bc, l, pad):
    torch.manual_seed(N + l)
    big = torch.randn(N, Bc, 4 * l)
    bg = big.to(cuda)
    pl = planes_buffer(1, N, Bc, 1, l, cuda)
    lib().stage_planes1d(dp(bg), Bc * 4 * l, dp(pl), N, Bc, l, pad, st())
    ref = planes_ref_1d(big, pad)
    assert torch.equal(pl.cpu().view(torch.int16), ref.reshape(-1).view(torch.int16))


@pytest.mark.parametrize("N,A", [(2, 8), (64, 8), (5, 12)])
def test_down2d_and_wgrad2d_from_planes(cuda, N, A):
    torch.manual_seed(N)
    Bc, s = 8, 64
    big = torch.randn(N, Bc, 2 * s, 2 * s)
    w = torch.randn(A, Bc, 4, 4) * 0.1
    bias = torch.randn(A)
    bg, wg, bsg = big.to(cuda), w.to(cuda), bias.to(cuda)
    wdn = image(wg, 2)
    pl = planes_buffer(2, N, Bc, s, s, cuda)
    lib().stage_planes2d(dp(bg), Bc * 4 * s * s, dp(pl), N, Bc, s, s, st())
    out_f = torch.empty(N, A, s, s, device=cuda)
    out_p = torch.empty(N, A, s, s, device=cuda)
    lib().down2d(dp(bg), Bc * 4 * s * s, dp(wdn), dp(bsg), None, 0, dp(out_f), A * s * s, N, A, Bc, s, s, 1, st())
    lib().down2d_planes(dp(pl), dp(wdn), dp(bsg), None, 0, dp(out_p), A * s * s, N, A, Bc, s, s, 1, st())
    assert torch.equal(out_p, out_f)
    assert rel_err(out_p, F.elu(F.conv2d(big, w, bias, stride=2, padding=1))) < TC_TOL
    # dgrad of the transposed conv with ELU'
    act = F.elu(torch.randn(N, A, s, s))
    ag = act.to(cuda)
    lib().down2d_planes(dp(pl), dp(wdn), None, dp(ag), A * s * s, dp(out_p), A * s * s, N, A, Bc, s, s, 2, st())
    ref = F.conv2d(big, w, None, stride=2, padding=1) * torch.where(act > 0, torch.ones_like(act), act + 1)
    assert rel_err(out_p, ref) < TC_TOL
    # weight gradient
    small = torch.randn(N, A, s, s)
    sg = small.to(cuda)
    wr = w.clone().requires_grad_()
    F.conv2d(big, wr, None, stride=2, padding=1).backward(small)
    dw = torch.empty(A, Bc, 4, 4, device=cuda)
    lib().wgrad2d_planes(dp(sg), A * s * s, dp(pl), dp(dw), N, A, Bc, s, s, st())
    assert rel_err(dw, wr.grad) < 2e-5


@pytest.mark.parametrize("N,A,pad", [(2, 8, 1), (64, 8, 0), (3, 12, 1)])
def test_down1d_and_wgrad1d_from_planes(cuda, N, A, pad):
    torch.manual_seed(N + 7)
    Bc, l = 8, 4096
    big = torch.randn(N, Bc, 4 * l)
    w = torch.randn(A, Bc, 4) * 0.1
    bias = torch.randn(A)
    bg, wg, bsg = big.to(cuda), w.to(cuda), bias.to(cuda)
    wdn = image(wg, 1)
    pl = planes_buffer(1, N, Bc, 1, l, cuda)
    lib().stage_planes1d(dp(bg), Bc * 4 * l, dp(pl), N, Bc, l, pad, st())
    out_f = torch.empty(N, A, l, device=cuda)
    out_p = torch.empty(N, A, l, device=cuda)
    lib().down1d(dp(bg), Bc * 4 * l, dp(wdn), dp(bsg), None, 0, dp(out_f), A * l, N, A, Bc, l, pad, 1, st())
    lib().down1d_planes(dp(pl), dp(wdn), dp(bsg), None, 0, dp(out_p), A * l, N, A, Bc, l, 1, st())
    assert torch.equal(out_p, out_f)
    assert rel_err(out_p, F.elu(F.conv1d(big, w, bias, stride=4, padding=pad))) < TC_TOL
    small = torch.randn(N, A, l)
    sg = small.to(cuda)
    wr = w.clone().requires_grad_()
    F.conv1d(big, wr, None, stride=4, padding=pad).backward(small)
    dw = torch.empty(A, Bc, 4, device=cuda)
    lib().wgrad1d_planes(dp(sg), A * l, dp(pl), dp(dw), N, A, Bc, l, st())
    assert rel_err(dw, wr.grad) < 2e-5


@pytest.mark.parametrize("N,A,Bc,l", [(2, 8, 8, 4096), (64, 8, 8, 4096), (5, 8, 4, 4096), (3, 12, 8, 256), (1024, 8, 8, 256)])
def test_fused_last_layer_backward_1d(cuda, N, A, Bc, l):
    """lshm_tconv_bwd1d_planes = lshm_wgrad1d_planes + lshm_down1d_planes(ELU') with the gradient planes read once:
    the data gradient is bit-identical to the separate kernel's (same MMA sequence per tile), the weight gradient
    agrees to split-K summation order; both against torch."""
    torch.manual_seed(N + A + l)
    big = torch.randn(N, Bc, 4 * l)                       # gradient w.r.t. the layer's output
    w = torch.randn(A, Bc, 4) * 0.1                       # ConvTranspose1d weight [in = A, out = Bc, 4]
    act = F.elu(torch.randn(N, A, l))                     # the layer's input (post-ELU activation of the layer before)
    bg, wg, ag = big.to(cuda), w.to(cuda), act.to(cuda)
    wdn = image(wg, 1)
    pl = planes_buffer(1, N, Bc, 1, l, cuda)
    lib().stage_planes1d(dp(bg), Bc * 4 * l, dp(pl), N, Bc, l, 0, st())
    dz_s, dw_s = torch.empty(N, A, l, device=cuda), torch.empty(A, Bc, 4, device=cuda)
    lib().down1d_planes(dp(pl), dp(wdn), None, dp(ag), A * l, dp(dz_s), A * l, N, A, Bc, l, 2, st())
    lib().wgrad1d_planes(dp(ag), A * l, dp(pl), dp(dw_s), N, A, Bc, l, st())
    dz_f, dw_f = torch.full((N, A, l), 7.0, device=cuda), torch.full((A, Bc, 4), 7.0, device=cuda)
    lib().tconv_bwd1d_planes(dp(ag), A * l, dp(pl), dp(wdn), dp(dz_f), A * l, dp(dw_f), N, A, Bc, l, st())
    assert torch.equal(dz_f, dz_s)
    assert rel_err(dw_f, dw_s) < 1e-5
    ar = act.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    F.conv_transpose1d(ar, wr, None, stride=4, padding=0).backward(big)
    assert rel_err(dz_f, ar.grad * torch.where(act > 0, torch.ones_like(act), act + 1)) < TC_TOL
    assert rel_err(dw_f, wr.grad) < 2e-5


@pytest.mark.parametrize("N,A,Bc,s", [(2, 8, 8, 64), (40, 8, 8, 64), (3, 8, 4, 64), (5, 4, 8, 16), (300, 8, 8, 16)])
def test_fused_last_layer_backward_2d(cuda, N, A, Bc, s):
    """lshm_tconv_bwd2d_planes = lshm_wgrad2d_planes + lshm_down2d_planes(ELU') from one read of the gradient planes."""
    torch.manual_seed(N + A + s)
    big = torch.randn(N, Bc, 2 * s, 2 * s)
    w = torch.randn(A, Bc, 4, 4) * 0.1                    # ConvTranspose2d weight [in = A, out = Bc, 4, 4]
    act = F.elu(torch.randn(N, A, s, s))
    bg, wg, ag = big.to(cuda), w.to(cuda), act.to(cuda)
    wdn = image(wg, 2)
    pl = planes_buffer(2, N, Bc, s, s, cuda)
    lib().stage_planes2d(dp(bg), Bc * 4 * s * s, dp(pl), N, Bc, s, s, st())
    dz_s, dw_s = torch.empty(N, A, s, s, device=cuda), torch.empty(A, Bc, 4, 4, device=cuda)
    lib().down2d_planes(dp(pl), dp(wdn), None, dp(ag), A * s * s, dp(dz_s), A * s * s, N, A, Bc, s, s, 2, st())
    lib().wgrad2d_planes(dp(ag), A * s * s, dp(pl), dp(dw_s), N, A, Bc, s, s, st())
    dz_f, dw_f = torch.full((N, A, s, s), 7.0, device=cuda), torch.full((A, Bc, 4, 4), 7.0, device=cuda)
    lib().tconv_bwd2d_planes(dp(ag), A * s * s, dp(pl), dp(wdn), dp(dz_f), A * s * s, dp(dw_f), N, A, Bc, s, s, st())
    assert torch.equal(dz_f, dz_s)
    assert rel_err(dw_f, dw_s) < 1e-5
    ar = act.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    F.conv_transpose2d(ar, wr, None, stride=2, padding=1).backward(big)
    assert rel_err(dz_f, ar.grad * torch.where(act > 0, torch.ones_like(act), act + 1)) < TC_TOL
    assert rel_err(dw_f, wr.grad) < 2e-5


@pytest.mark.parametrize("dim,N,A,small", [(1, 3, 12, 1024), (1, 256, 12, 1024), (1, 5, 16, 256), (2, 3, 12, 32), (2, 130, 12, 32), (2, 7, 10, 16)])
def test_fused_second_layer_backward(cuda, dim, N, A, small):
    """lshm_tconv_bwd1d / 2d = lshm_wgrad*d + lshm_down*d(ELU') on an fp32 output gradient gathered once."""
    torch.manual_seed(N + A + small)
    Bc = 8
    if dim == 2:
        big = torch.randn(N, Bc, 2 * small, 2 * small); act = F.elu(torch.randn(N, A, small, small)); w = torch.randn(A, Bc, 4, 4) * 0.1
    else:
        big = torch.randn(N, Bc, 4 * small); act = F.elu(torch.randn(N, A, small)); w = torch.randn(A, Bc, 4) * 0.1
    bg, wg, ag = big.to(cuda), w.to(cuda), act.to(cuda)
    wdn = image(wg, dim)
    bns, sns = big[0].numel(), act[0].numel()
    dz_s, dw_s = torch.empty_like(ag), torch.empty_like(wg)
    dz_f, dw_f = torch.full_like(ag, 7.0), torch.full_like(wg, 7.0)
    if dim == 2:
        lib().down2d(dp(bg), bns, dp(wdn), None, dp(ag), sns, dp(dz_s), sns, N, A, Bc, small, small, 2, st())
        lib().wgrad2d(dp(ag), sns, dp(bg), bns, dp(dw_s), N, A, Bc, small, small, st())
        lib().tconv_bwd2d(dp(ag), sns, dp(bg), bns, dp(wdn), dp(dz_f), sns, dp(dw_f), N, A, Bc, small, small, st())
    else:
        lib().down1d(dp(bg), bns, dp(wdn), None, dp(ag), sns, dp(dz_s), sns, N, A, Bc, small, 0, 2, st())
        lib().wgrad1d(dp(ag), sns, dp(bg), bns, dp(dw_s), N, A, Bc, small, 0, st())
        lib().tconv_bwd1d(dp(ag), sns, dp(bg), bns, dp(wdn), dp(dz_f), sns, dp(dw_f), N, A, Bc, small, st())
    assert torch.equal(dz_f, dz_s)
    assert rel_err(dw_f, dw_s) < 1e-5
    ar = act.clone().requires_grad_()
    wr = w.clone().requires_grad_()
    (F.conv_transpose2d(ar, wr, None, stride=2, padding=1) if dim == 2 else F.conv_transpose1d(ar, wr, None, stride=4)).backward(big)
    assert rel_err(dz_f, ar.grad * torch.where(act > 0, torch.ones_like(act), act + 1)) < TC_TOL
    assert rel_err(dw_f, wr.grad) < 2e-5


@pytest.mark.parametrize("N,C", [(3, 8), (2, 4)])
def test_fused_plane_writers(cuda, N, C):
    torch.manual_seed(N * C)
    P = 128
    x, x1 = torch.randn(N, C, P, P, device=cuda), torch.randn(N, C, P, P, device=cuda)
    iyT, iyF = torch.empty_like(x), torch.empty_like(x)
    lib().residual_split(dp(x), dp(x1), dp(iyT), dp(iyF), N, C, P, st())
    l = P * P // 4
    refT, refF = planes_buffer(1, N, C, 1, l, cuda), planes_buffer(1, N, C, 1, l, cuda)
    lib().stage_planes1d(dp(iyT), C * P * P, dp(refT), N, C, l, 1, st())
    lib().stage_planes1d(dp(iyF), C * P * P, dp(refF), N, C, l, 1, st())
    pT, pF = torch.zeros_like(refT), torch.zeros_like(refF)
    lib().residual_split_planes(dp(x), dp(x1), dp(pT), dp(pF), N, C, P, st())
    assert torch.equal(pT.view(torch.int16), refT.view(torch.int16))
    assert torch.equal(pF.view(torch.int16), refF.view(torch.int16))
    # gradient combine -> 2-D planes (+ bias-gradient sums)
    g1p, gT, gF = (torch.randn(N, C, P, P, device=cuda) for _ in range(3))
    gx1 = torch.empty_like(g1p)
    db_ref = torch.empty(C, device=cuda)
    lib().cascade_combine(dp(g1p), dp(gT), dp(gF), dp(gx1), N, C, P, dp(db_ref), st())
    ref = planes_buffer(2, N, C, P // 2, P // 2, cuda)
    lib().stage_planes2d(dp(gx1), C * P * P, dp(ref), N, C, P // 2, P // 2, st())
    got = torch.zeros_like(ref)
    db = torch.full((C,), 3.0, device=cuda)
    lib().cascade_combine_planes(dp(g1p), dp(gT), dp(gF), dp(got), N, C, P, dp(db), st())
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    assert rel_err(db, db_ref) < 1e-5


@pytest.mark.parametrize("C,N", [(8, 8), (4, 6)])
def test_training_closure_with_and_without_operand_planes(cuda, C, N):
    """The fused closure with the input-sized tensors as operand planes (default) against the same closure on fp32
    tensors: identical loss columns and latents (the first-layer forward results are bit-identical), gradients equal
    up to the order of the split-K atomics."""
    from common import SCALES, closure_case
    from lshm_b200.kharmonic_lofar import DeepKHarmonicStep
    from lshm_b200.lofar_models import AutoEncoder1DCNN, AutoEncoderCNN2, Kmeans
    case = closure_case(C=C, L=32, Lt=16, N=N, bpb=2, seed=31)

    def run(use_planes):
        hs = torch.tensor(SCALES).to(cuda)
        net = AutoEncoderCNN2(32, C, hs, True); netT = AutoEncoder1DCNN(16, C, hs, True); netF = AutoEncoder1DCNN(16, C, hs, True)
        mod = Kmeans(64, case["K"], 4)
        net.load_state_dict(case["pn"]); netT.load_state_dict(case["pT"]); netF.load_state_dict(case["pF"])
        mod.load_state_dict({"M": case["M"]})
        step = DeepKHarmonicStep(net.to(cuda), netT.to(cuda), netF.to(cuda), mod.to(cuda), use_planes=use_planes)
        step.set_batch(case["x"].to(cuda), case["uv"].to(cuda), 2)
        for dst, src in zip((step.y1, step.y2, step.y3), case["ys"]):
            dst.copy_(src.to(cuda))
        step.closure()
        return step

    a, b = run(True), run(False)
    assert a.mb[0].xp is not None and b.mb[0].xp is None
    ta, tb = a.loss_terms(), b.loss_terms()
    for k in ta:
        assert abs(ta[k] - tb[k]) <= 1e-6 * abs(tb[k]) + 1e-12, (k, ta[k], tb[k])
    assert torch.equal(a.latents(), b.latents())
    for nm, pa, pb in zip(a.flat.names, a.flat.params, b.flat.params):
        assert rel_err(pa.grad, pb.grad) < 2e-5, nm


@pytest.mark.parametrize("N,C,upd", [(3, 8, 1), (2, 4, 0)])
def test_loss_pass_with_plane_outputs(cuda, N, C, upd):
    """lshm_cascade_losses_planes against lshm_cascade_losses_upd (fp32 gradients) + staging: same sums, same g1p, same
    multipliers, same bias-gradient sums, and the two gradient planes equal the staged fp32 gradients."""
    torch.manual_seed(N * 10 + C)
    P, rho = 128, 0.7
    x, x1, x2, x3f = (torch.randn(N, C, P, P, device=cuda) for _ in range(4))
    ys = [torch.randn(N * C * P * P, device=cuda) for _ in range(3)]
    n = x.numel()
    ya = [t.clone() for t in ys]
    sums_a = torch.zeros(8, dtype=torch.float64, device=cuda)
    g1a, g2a, g3a = (torch.empty_like(x) for _ in range(3))
    db2a, db3a = torch.empty(C, device=cuda), torch.empty(C, device=cuda)
    lib().cascade_losses_upd(dp(x), dp(x1), dp(x2), dp(x3f), dp(ya[0]), dp(ya[1]), dp(ya[2]), rho, upd, N, C, P, 1.0 / n,
                             dp(sums_a), dp(g1a), dp(g2a), dp(g3a), dp(db2a), dp(db3a), st())
    l = P * P // 4
    ref2, ref3 = planes_buffer(1, N, C, 1, l, cuda), planes_buffer(1, N, C, 1, l, cuda)
    lib().stage_planes1d(dp(g2a), C * P * P, dp(ref2), N, C, l, 0, st())
    lib().stage_planes1d(dp(g3a), C * P * P, dp(ref3), N, C, l, 0, st())
    yb = [t.clone() for t in ys]
    sums_b = torch.zeros(8, dtype=torch.float64, device=cuda)
    g1b = torch.empty_like(x)
    p2, p3 = planes_buffer(1, N, C, 1, l, cuda), planes_buffer(1, N, C, 1, l, cuda)
    db2b, db3b = torch.full((C,), 5.0, device=cuda), torch.full((C,), 5.0, device=cuda)
    lib().cascade_losses_planes(dp(x), dp(x1), dp(x2), dp(x3f), dp(yb[0]), dp(yb[1]), dp(yb[2]), rho, upd, N, C, P, 1.0 / n,
                                dp(sums_b), dp(g1b), dp(p2), dp(p3), dp(db2b), dp(db3b), st())
    assert torch.allclose(sums_b, sums_a, rtol=1e-6)
    for a, b in zip(ya, yb):
        assert torch.equal(a, b)
    assert rel_err(g1b, g1a) < 1e-7
    # the planes hold bf16 hi/lo of the same fp32 values: decode and compare (a last-bit fp32 difference from a
    # different multiply-add contraction would flip low bits of lo, so compare values, not bit patterns)
    def decode(pl):
        h = pl.view(2, -1).float()
        return h[0] + h[1]
    assert rel_err(decode(p2), decode(ref2)) < 1e-6 and rel_err(decode(p3), decode(ref3)) < 1e-6
    assert rel_err(db2b, db2a) < 1e-5 and rel_err(db3b, db3a) < 1e-5
