"""Per-op parity on the GPU: each autograd op of ic_gan_b200.ops against the same op written in plain fp32 PyTorch
(the reference's own ATen calls), forward and backward, in parity (fp32) and throughput (bf16) dtypes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# the PyTorch side is the fp32 reference: no TF32 in cuDNN/cuBLAS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

TOL = {torch.float32: 2e-4, torch.bfloat16: 3e-2}


def _close(a, b, tol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    scale = max(1.0, b.abs().max().item())
    assert err <= tol * scale, f"{what}: max-abs err {err:.3e} (scale {scale:.3g}, tol {tol})"


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 32, 48, 16, 3), (2, 64, 96, 8, 3), (4, 16, 32, 4, 1), (2, 3, 16, 16, 3),
                                   (2, 16, 3, 16, 3)])
def test_snconv2d_forward_backward(cuda_device, cdt, shape):
    from ic_gan_b200.biggan import layers
    B, cin, cout, H, k = shape
    torch.manual_seed(0)
    m = layers.SNConv2d(cin, cout, k, padding=k // 2, eps=1e-6).to(cuda_device)
    m.compute_dtype = cdt
    m.train()
    x = torch.randn(B, cin, H, H, device=cuda_device)
    u0 = m.u0.clone()
    xq = x.to(cdt).float() if cdt != torch.float32 else x
    x1 = xq.clone().requires_grad_(True)
    y = m(x1)
    gy = torch.randn_like(y.float())
    y.float().backward(gy)
    # plain-PyTorch restatement (layers.py:98-112, :144-153)
    W = m.weight.detach().clone().requires_grad_(True)
    Wm = W.reshape(cout, -1)
    with torch.no_grad():
        v = F.normalize(u0 @ Wm, eps=1e-6)
        un = F.normalize(v @ Wm.t(), eps=1e-6)
    sigma = (v @ Wm.t() @ un.t()).squeeze()
    x2 = xq.clone().requires_grad_(True)
    bias = m.bias.detach().clone().requires_grad_(True)
    yr = F.conv2d(x2, W / sigma, bias, 1, k // 2)
    yr.backward(gy.to(cdt).float() if cdt != torch.float32 else gy)
    tol = TOL[cdt]
    _close(y, yr, tol, "conv fwd")
    _close(m.u0, un, 1e-5, "u update")
    _close(m.sv0, sigma.reshape(1), 1e-5, "sv0")
    _close(x1.grad, x2.grad, tol, "conv dgrad")
    _close(m.weight.grad, W.grad, tol * 2, "conv wgrad (through sigma)")
    _close(m.bias.grad, bias.grad, tol * 2, "conv bias grad")


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("up", [False, True])
def test_ccbn_relu_up(cuda_device, cdt, up):
    from ic_gan_b200 import ops
    torch.manual_seed(1)
    B, C, H = 4, 32, 8
    x = torch.randn(B, C, H, H, device=cuda_device) * 2 + 0.5
    gain = (1 + 0.3 * torch.randn(B, C, device=cuda_device)).requires_grad_(True)
    bias = (0.3 * torch.randn(B, C, device=cuda_device)).requires_grad_(True)
    rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    xq = x.to(cdt).float()
    x1 = _nhwc(xq).to(cdt).requires_grad_(True)
    y = ops.BNActFn.apply(x1, gain, bias, rm, rv, True, 1e-5, 0.1, True, up, cdt)
    gy = torch.randn_like(y.float())
    y.float().backward(gy)
    x2 = xq.clone().requires_grad_(True)
    g2, b2 = gain.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
    yr = F.batch_norm(x2, rm2, rv2, None, None, True, 0.1, 1e-5) * g2[:, :, None, None] + b2[:, :, None, None]
    yr = F.relu(yr)
    if up:
        yr = F.interpolate(yr, scale_factor=2)
    gyr = gy.to(cdt).float().permute(0, 3, 1, 2) if cdt != torch.float32 else gy.permute(0, 3, 1, 2)
    yr.backward(gyr)
    tol = TOL[cdt]
    _close(y.permute(0, 3, 1, 2), yr, tol, "bn fwd")
    _close(rm, rm2, 1e-5, "running mean")
    _close(rv, rv2, 1e-5, "running var")
    _close(x1.grad.permute(0, 3, 1, 2), x2.grad, tol * 2, "bn dx")
    _close(gain.grad, g2.grad, tol * 4, "bn dgain")
    _close(bias.grad, b2.grad, tol * 4, "bn dbias")


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_attention_block(cuda_device, cdt):
    from ic_gan_b200.biggan import layers
    import functools
    torch.manual_seed(2)
    C, B, H = 64, 2, 16
    att = layers.Attention(C, functools.partial(layers.SNConv2d, eps=1e-6)).to(cuda_device)
    for m in att.modules():
        if isinstance(m, layers.SN):
            m.compute_dtype = cdt
    with torch.no_grad():
        att.gamma.fill_(0.7)
    att.train()
    u = {n: getattr(att, n).u0.clone() for n in ("theta", "phi", "g", "o")}
    x = torch.randn(B, C, H, H, device=cuda_device)
    xq = x.to(cdt).float()
    x1 = xq.to(cdt).clone().requires_grad_(True)
    y = att(x1)
    gy = torch.randn_like(y.float())
    y.float().backward(gy)

    def snw(name):
        W = getattr(att, name).weight.detach()
        Wm = W.reshape(W.shape[0], -1)
        v = F.normalize(u[name] @ Wm, eps=1e-6)
        un = F.normalize(v @ Wm.t(), eps=1e-6)
        return W / (v @ Wm.t() @ un.t()).squeeze()
    x2 = xq.clone().requires_grad_(True)
    theta = F.conv2d(x2, snw("theta"))
    phi = F.max_pool2d(F.conv2d(x2, snw("phi")), 2)
    g = F.max_pool2d(F.conv2d(x2, snw("g")), 2)
    theta = theta.view(B, C // 8, H * H)
    phi = phi.view(B, C // 8, H * H // 4)
    g = g.view(B, C // 2, H * H // 4)
    beta = F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1)
    o = F.conv2d(torch.bmm(g, beta.transpose(1, 2)).view(B, C // 2, H, H), snw("o"))
    yr = 0.7 * o + x2
    yr.backward(gy.to(cdt).float() if cdt != torch.float32 else gy)
    tol = TOL[cdt] * 2
    _close(y, yr, tol, "attention fwd")
    _close(x1.grad, x2.grad, tol * 2, "attention dx")


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_pool_relu_sumpool_linear(cuda_device, cdt):
    from ic_gan_b200 import ops
    from ic_gan_b200.biggan import layers
    torch.manual_seed(3)
    x = torch.randn(3, 24, 8, 8, device=cuda_device).to(cdt).float()
    add = torch.randn(3, 24, 4, 4, device=cuda_device).to(cdt).float()
    x1 = _nhwc(x).to(cdt).requires_grad_(True)
    a1 = _nhwc(add).to(cdt).requires_grad_(True)
    y = ops.Pool2Fn.apply(ops.ReluFn.apply(x1), a1, 0.25, 0)
    h = ops.ReluSumPoolFn.apply(y)
    mx = ops.Pool2Fn.apply(x1, None, 1.0, 1)
    (h.sum() * 0.5 + (mx.float() ** 2).sum()).backward()
    x2, a2 = x.clone().requires_grad_(True), add.clone().requires_grad_(True)
    yr = F.avg_pool2d(F.relu(x2), 2) + a2
    hr = F.relu(yr).sum(dim=(2, 3))
    mr = F.max_pool2d(x2, 2)
    (hr.sum() * 0.5 + (mr ** 2).sum()).backward()
    tol = TOL[cdt]
    _close(y.permute(0, 3, 1, 2), yr, tol, "avgpool+add")
    _close(h, hr, tol * 4, "relu sumpool")
    _close(mx.permute(0, 3, 1, 2), mr, tol, "maxpool")
    _close(x1.grad.permute(0, 3, 1, 2), x2.grad, tol * 4, "dx")
    _close(a1.grad.permute(0, 3, 1, 2), a2.grad, tol * 4, "dadd")
    lin = layers.SNLinear(37, 19, eps=1e-6).to(cuda_device)
    lin.train()
    u0 = lin.u0.clone()
    xi = torch.randn(5, 37, device=cuda_device, requires_grad=True)
    out = lin(xi)
    out.pow(2).sum().backward()
    W = lin.weight.detach().clone().requires_grad_(True)
    v = F.normalize(u0 @ W.detach(), eps=1e-6)
    un = F.normalize(v @ W.detach().t(), eps=1e-6)
    sigma = (v @ W.t() @ un.t()).squeeze()
    xr = xi.detach().clone().requires_grad_(True)
    outr = F.linear(xr, W / sigma, lin.bias.detach())
    outr.pow(2).sum().backward()
    _close(out, outr, 1e-4, "linear fwd")
    _close(xi.grad, xr.grad, 1e-4, "linear dx")
    _close(lin.weight.grad, W.grad, 1e-4, "linear dW through sigma")


@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_operand_copies_follow_weight_updates(cuda_device, cdt):
    """The bf16/fp32 operand copies are rebuilt lazily (tensor version counter): an optimizer step must be seen."""
    from ic_gan_b200.biggan import layers
    torch.manual_seed(5)
    m = layers.SNConv2d(32, 16, 3, padding=1, eps=1e-6).to(cuda_device)
    m.compute_dtype = cdt
    m.eval()
    x = torch.randn(2, 32, 8, 8, device=cuda_device).to(cdt).float()
    opt = torch.optim.Adam(m.parameters(), lr=0.05, fused=True)
    for _ in range(2):
        y = m(x.to(cdt))
        W = m.weight.detach()
        Wm = W.reshape(16, -1)
        v = F.normalize(m.u0 @ Wm, eps=1e-6)
        un = F.normalize(v @ Wm.t(), eps=1e-6)
        ref = F.conv2d(x, W / (v @ Wm.t() @ un.t()).squeeze(), m.bias.detach(), 1, 1)
        _close(y, ref, TOL[cdt], "forward after weight update")
        m.train()
        m(x.to(cdt)).float().pow(2).mean().backward()
        opt.step()
        m.eval()


@pytest.mark.parametrize("shape", [(2, 24, 40, 96, 3, 3), (1, 16, 16, 8, 3, 1), (3, 20, 12, 200, 2, 3)])
@pytest.mark.parametrize("act", [0, 1])
def test_rgb_input_conv_c_abi(cuda_device, shape, act):
    """icgan_conv2d_rgb_tc (im2col fused into the tensor-core kernel) against F.conv2d on the same bf16-rounded inputs;
    pixel counts that are not multiples of the 128-row tile, a Cout tail, ReLU epilogue."""
    from ic_gan_b200 import _lib as L
    from ic_gan_b200._lib import call, dt, ptr
    B, H, W, cout, cs, k = shape
    torch.manual_seed(1)
    x = torch.randn(B, cs, H, W, device=cuda_device).to(torch.bfloat16)
    w = (torch.randn(cout, cs, k, k, device=cuda_device) * 0.3).to(torch.bfloat16)
    bias = torch.randn(cout, device=cuda_device)
    alpha = torch.tensor([0.7], device=cuda_device)
    wcol = torch.zeros(cout, 32, device=cuda_device, dtype=torch.bfloat16)
    wcol[:, :k * k * cs] = w.permute(0, 2, 3, 1).reshape(cout, -1)
    xn = _nhwc(x)
    y = torch.empty(B, H, W, cout, device=cuda_device, dtype=torch.bfloat16)
    call("icgan_conv2d_rgb_tc", ptr(xn), ptr(wcol), ptr(alpha), ptr(bias), ptr(y), B, H, W, cs, cout, k, dt(y),
         L.ACT_RELU if act else L.ACT_NONE, torch.cuda.current_stream().cuda_stream)
    ref = 0.7 * F.conv2d(x.float(), w.float(), None, 1, k // 2) + bias.view(1, -1, 1, 1)
    if act:
        ref = torch.relu(ref)
    _close(y.permute(0, 3, 1, 2), ref, 1e-2, "rgb conv")


def test_subpixel_upconv_matches_upsample_then_conv(cuda_device):
    """ops.UpConvFn (conv3x3(nearest_up2(x)) as four 2x2-tap parity classes on the tap-table tensor-core kernels, 16 instead
    of 36 MACs per low-resolution pixel) against the plain path (BN writes the upsampled tensor, 3x3 conv at high resolution):
    GBlock output and every gradient, bf16."""
    import copy
    import functools
    from ic_gan_b200 import ops
    from ic_gan_b200.biggan import layers
    conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=1, num_itrs=1, eps=1e-8)
    lin = functools.partial(layers.SNLinear, num_svs=1, num_itrs=1, eps=1e-8, bias=False)
    bn = functools.partial(layers.ccbn, which_linear=lin, input_size=20, norm_style="bn", eps=1e-5)
    torch.manual_seed(3)
    ref_blk = layers.GBlock(64, 32, which_conv=conv, which_bn=bn, activation=torch.nn.ReLU(), upsample=True).to(cuda_device)
    sub_blk = copy.deepcopy(ref_blk)
    for blk in (ref_blk, sub_blk):
        for m in blk.modules():
            if isinstance(m, layers.SN):
                m.compute_dtype = torch.bfloat16
        blk.train()
    g = torch.Generator(device=cuda_device).manual_seed(4)
    x0 = torch.randn(3, 64, 16, 16, device=cuda_device, generator=g).bfloat16()
    y0 = torch.randn(3, 20, device=cuda_device, generator=g)
    outs = []
    old = ops.SUBPIXEL_UP
    try:
        for flag, blk in ((False, ref_blk), (True, sub_blk)):
            ops.SUBPIXEL_UP = flag
            x, y = x0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
            out = blk(x, y)
            gy = torch.randn(out.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(9))
            out.float().backward(gy)
            outs.append((out.float(), x.grad.float(), {k: p.grad.clone() for k, p in blk.named_parameters()}))
    finally:
        ops.SUBPIXEL_UP = old
    assert getattr(sub_blk.conv1, "sub_pixel_up", False) and not getattr(ref_blk.conv1, "sub_pixel_up", False)
    (o0, dx0, g0), (o1, dx1, g1) = outs
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))
    assert rel(o1, o0) <= 2e-2 and rel(dx1, dx0) <= 3e-2, (rel(o1, o0), rel(dx1, dx0))
    for k in g0:
        if g0[k].abs().max() > 1e-4 and k != "conv1.bias":  # conv1.bias feeds bn2: analytically zero, rounding noise only
            assert rel(g1[k], g0[k]) <= 5e-2, (k, rel(g1[k], g0[k]))


def test_pooled_downconv_matches_conv_then_pool(cuda_device):
    """ops.DownConvFn (avgpool2(conv3x3(h)) + shortcut as one stride-2 4x4 convolution, 16 instead of 36 MACs per output pixel;
    transposed parity classes with the ReLU gate for dgrad; tap-table wgrad) against the plain DBlock tail, bf16."""
    import copy
    import functools
    from ic_gan_b200 import ops
    from ic_gan_b200.biggan import layers
    conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=1, num_itrs=1, eps=1e-8)
    torch.manual_seed(5)
    ref_blk = layers.DBlock(32, 64, which_conv=conv, wide=True, preactivation=True, activation=torch.nn.ReLU(),
                            downsample=torch.nn.AvgPool2d(2)).to(cuda_device)
    new_blk = copy.deepcopy(ref_blk)
    for blk in (ref_blk, new_blk):
        for m in blk.modules():
            if isinstance(m, layers.SN):
                m.compute_dtype = torch.bfloat16
        blk.train()
    g = torch.Generator(device=cuda_device).manual_seed(6)
    x0 = torch.randn(3, 32, 32, 32, device=cuda_device, generator=g).bfloat16()
    outs = []
    old = ops.POOLED_DOWN
    try:
        for flag, blk in ((False, ref_blk), (True, new_blk)):
            ops.POOLED_DOWN = flag
            x = x0.clone().requires_grad_(True)
            out = blk(x)
            gy = torch.randn(out.shape, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(9))
            out.float().backward(gy)
            outs.append((out.float(), x.grad.float(), {k: p.grad.clone() for k, p in blk.named_parameters()}))
    finally:
        ops.POOLED_DOWN = old
    assert getattr(new_blk.conv2, "pooled_down", False) and not getattr(ref_blk.conv2, "pooled_down", False)
    (o0, dx0, g0), (o1, dx1, g1) = outs
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))
    assert rel(o1, o0) <= 2e-2 and rel(dx1, dx0) <= 3e-2, (rel(o1, o0), rel(dx1, dx0))
    for k in g0:
        assert rel(g1[k], g0[k]) <= 5e-2, (k, rel(g1[k], g0[k]))
