"""Kernel-level parity on the B200: every C-ABI kernel against a plain PyTorch fp32 evaluation of
the same op on the same (bf16-rounded) inputs.  Tolerances: one bf16 rounding of the output
(2^-8 relative) for bf16 results, 1e-4 relative for fp32 results (fp32 accumulation order)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def isx():
    import iris_b200
    from iris_b200 import _lib

    _lib.load()
    _lib.call("isx_device_check", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib


def nhwc_bf16(B, H, W, C, seed, scale=1.0, relu=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(B, H, W, C, device="cuda", generator=g) * scale
    if relu:
        t = t.clamp_min(0)
    return t.to(torch.bfloat16).contiguous()


def pack(isx, w):
    Cout, Cin = w.shape[:2]
    wf = torch.empty(9, Cout, Cin, device="cuda", dtype=torch.bfloat16)
    wd = torch.empty(9, Cin, Cout, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_pack_conv3x3_weights", w.contiguous(), Cout, Cin, wf, wd, isx.stream_ptr())
    return wf, wd


def assert_close_bf16(got, ref, what):
    got = got.float()
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-2 * ref.abs().mean() + 1e-6
    bad = (err > tol).sum().item()
    assert bad == 0, "%s: %d / %d elements off (max err %.4g, ref absmax %.4g)" % (
        what, bad, ref.numel(), err.max().item(), ref.abs().max().item())


CONV_SHAPES = [
    # B, H, W, Cin, Cout
    (2, 20, 24, 64, 64),
    (1, 50, 80, 128, 256),
    (3, 13, 9, 256, 128),    # ragged tiles in every dimension
    (1, 25, 40, 512, 512),
    (5, 6, 10, 64, 128),     # several images per tile
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("cfg", [0, 6410, 6420, 12810, 12820, 25610, 25620, 12813])
def test_conv3x3_fwd(isx, shape, cfg):
    B, H, W, Cin, Cout = shape
    bn = cfg // 100
    if bn and Cout % bn:
        pytest.skip("BN does not divide Cout")
    x = nhwc_bf16(B, H, W, Cin, 1, relu=True)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (2.0 / (9 * Cin)) ** 0.5
    bias = torch.randn(Cout, device="cuda") * 0.1
    wf, _ = pack(isx, w)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, out, B, H, W, Cin, Cout, 1, cfg, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1))
    assert_close_bf16(out, ref.permute(0, 2, 3, 1), "conv fwd %s cfg %d" % (shape, cfg))


@pytest.mark.parametrize("shape", CONV_SHAPES + [(2, 40, 48, 64, 64), (1, 100, 160, 128, 256)])
@pytest.mark.parametrize("cfg", [0, 6410, 12820, 25610])
def test_conv3x3_fwd_fused_pool(isx, shape, cfg):
    """Conv + ReLU with MaxPool2d(2,2) emitted from the epilogue: both outputs against torch (pool is exact on the bf16 tile)."""
    B, H, W, Cin, Cout = shape
    bn = cfg // 100
    if bn and Cout % bn:
        pytest.skip("BN does not divide Cout")
    x = nhwc_bf16(B, H, W, Cin, 1, relu=True)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (2.0 / (9 * Cin)) ** 0.5
    bias = torch.randn(Cout, device="cuda") * 0.1
    wf, _ = pack(isx, w)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    pool = torch.full((B, H // 2, W // 2, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_conv3x3_bias_relu_pool_fwd", x, wf, bias, out, pool, B, H, W, Cin, Cout, cfg, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1))
    assert_close_bf16(out, ref.permute(0, 2, 3, 1), "conv fwd (+pool) %s cfg %d" % (shape, cfg))
    ref_pool = F.max_pool2d(out.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(pool.float(), ref_pool)
    # the same launch with the routing bytes of the pool's backward, with and without the full-resolution store
    ref_idx = torch.empty(B, H // 2, W // 2, Cout, device="cuda", dtype=torch.uint8)
    isx.call("isx_maxpool2x2_fwd_idx", out, None, ref_idx, B, H, W, Cout, isx.stream_ptr())
    for skip in (0, 1):
        out2 = torch.full((B, H, W, Cout), 7.0, device="cuda", dtype=torch.bfloat16)
        pool2 = torch.full((B, H // 2, W // 2, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        idx2 = torch.full((B, H // 2, W // 2, Cout), 255, device="cuda", dtype=torch.uint8)
        isx.call("isx_conv3x3_bias_relu_pool_idx_fwd", x, wf, bias, out2, pool2, idx2, skip, B, H, W, Cin, Cout, cfg,
                 isx.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(pool2, pool), "pooled map differs (skip_out=%d)" % skip
        assert torch.equal(idx2, ref_idx), "routing bytes differ (skip_out=%d): %d" % (skip, int((idx2 != ref_idx).sum()))
        untouched = bool((out2 == 7.0).all())
        assert torch.equal(out2, out) or (skip == 1 and untouched), "full-resolution output (skip_out=%d)" % skip


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("mode", ["plain", "mask", "mask_add", "mask_affine"])
def test_conv3x3_dgrad(isx, shape, mode):
    B, H, W, Cin, Cout = shape
    dy = nhwc_bf16(B, H, W, Cout, 2)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (2.0 / (9 * Cout)) ** 0.5
    _, wd = pack(isx, w)
    act = nhwc_bf16(B, H, W, Cin, 3, relu=True)
    add = nhwc_bf16(B, H, W, Cin, 4, scale=0.5)
    aa = torch.randn(B, Cin, device="cuda") * 0.3
    ab = torch.randn(B, Cin, device="cuda") * 0.3
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_conv3x3_dgrad", dy, wd, dx, B, H, W, Cin, Cout,
             act if mode != "plain" else None, add if mode == "mask_add" else None,
             aa if mode == "mask_affine" else None, ab if mode == "mask_affine" else None, 0, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    if mode == "mask_add":
        ref = ref + add.float()
    if mode == "mask_affine":
        ref = ref + aa[:, None, None, :] + ab[:, None, None, :] * act.float()
    if mode != "plain":
        ref = torch.where(act.float() > 0, ref, torch.zeros_like(ref))
    assert_close_bf16(dx, ref, "dgrad %s %s" % (shape, mode))


@pytest.mark.parametrize("shape", [(2, 20, 24, 64, 64), (3, 13, 9, 128, 128), (1, 50, 80, 256, 256), (2, 25, 40, 512, 512),
                                   (2, 16, 24, 64, 128)])
def test_conv3x3_dgrad_fused_gram(isx, shape):
    """dx = relu'(act) * (dgrad(dy) + act . D[b]): the Gram tap gradient as extra K blocks of the dgrad GEMM."""
    B, H, W, Cin, Cout = shape
    dy = nhwc_bf16(B, H, W, Cout, 2)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (2.0 / (9 * Cout)) ** 0.5
    _, wd = pack(isx, w)
    act = nhwc_bf16(B, H, W, Cin, 3, relu=True)
    D = torch.randn(B, Cin, Cin, device="cuda") * 0.05
    D = (D + D.transpose(1, 2)).to(torch.bfloat16).contiguous()
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_conv3x3_dgrad_gram", dy, wd, dx, B, H, W, Cin, Cout, act, D, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    ref = ref + torch.bmm(act.float().reshape(B, H * W, Cin), D.float()).reshape(B, H, W, Cin)
    ref = torch.where(act.float() > 0, ref, torch.zeros_like(ref))
    assert_close_bf16(dx, ref, "dgrad + fused gram %s" % (shape,))


# ---- the specialised Cin = 64 kernel (conv_c64.cu): persistent CTAs, resident weights, halo patch --------------------
C64_SHAPES = [
    (2, 20, 24, 64, 64),      # fewer tiles than SMs
    (3, 13, 9, 64, 64),       # ragged in x and y
    (4, 160, 200, 64, 64),    # ~7 tiles per CTA: the halo ring and both TMEM accumulator sets wrap
    (2, 41, 37, 64, 64),      # odd sizes: fused pool drops the last row/column
]


@pytest.fixture
def c64_forced(isx):
    """Route every applicable call through conv_c64 (option 2 also takes the N = 16 tail and inputs smaller than one
    tile per SM, which the default heuristic leaves to the generic kernel)."""
    lib = isx.load()
    assert lib.isx_set_option(b"c64", 2) == 0
    assert lib.isx_set_option(b"sweep64", 0) == 0
    yield
    assert lib.isx_set_option(b"c64", 1) == 0
    assert lib.isx_set_option(b"sweep64", 1) == 0


@pytest.mark.parametrize("shape", C64_SHAPES)
def test_conv_c64_fwd_and_pool(isx, c64_forced, shape):
    test_conv3x3_fwd(isx, shape, 0)
    test_conv3x3_fwd_fused_pool(isx, shape, 0)


@pytest.mark.parametrize("shape", C64_SHAPES)
@pytest.mark.parametrize("mode", ["plain", "mask", "mask_add", "mask_affine", "gram"])
def test_conv_c64_dgrad(isx, c64_forced, shape, mode):
    if mode == "gram":
        test_conv3x3_dgrad_fused_gram(isx, shape)
    else:
        test_conv3x3_dgrad(isx, shape, mode)


@pytest.mark.parametrize("xc", [3, 1])
@pytest.mark.parametrize("use_mask", [False, True])
def test_conv_c64_tail(isx, c64_forced, xc, use_mask):
    test_conv1_1_fwd_dgrad(isx, xc, use_mask)


# ---- the tap-stacked sweep kernel (conv_sweep.cu): 128-pixel strips, N = 192 MMAs into a ring of eight TMEM accumulators ----
SWEEP_SHAPES = [
    (2, 20, 24, 64, 64),      # one partly filled strip along x, one short segment
    (3, 13, 9, 64, 64),       # ragged, odd sizes
    (2, 41, 37, 64, 64),      # odd sizes: the fused pool drops the last row / column, strips along y
    (1, 256, 40, 64, 64),     # strips along y (two full strips), sweep along x
    (2, 128, 230, 64, 64),    # strips along y, two sweep segments (116 + 114): the accumulator ring wraps many times
    (4, 160, 200, 64, 64),    # strips along x (200 = 128 + 72), two segments of 80 rows, several jobs per CTA
    (1, 300, 128, 64, 64),    # strips along x exactly 128 wide, three segments
]


@pytest.fixture
def sweep_forced(isx):
    lib = isx.load()
    assert lib.isx_set_option(b"sweep64", 2) == 0
    yield
    assert lib.isx_set_option(b"sweep64", 1) == 0


@pytest.mark.parametrize("shape", SWEEP_SHAPES)
def test_conv_sweep_fwd_and_pool(isx, sweep_forced, shape):
    test_conv3x3_fwd(isx, shape, 0)
    test_conv3x3_fwd_fused_pool(isx, shape, 0)


@pytest.mark.parametrize("shape", SWEEP_SHAPES)
@pytest.mark.parametrize("mode", ["plain", "mask", "mask_add", "mask_affine", "gram"])
def test_conv_sweep_dgrad(isx, sweep_forced, shape, mode):
    if mode == "gram":
        test_conv3x3_dgrad_fused_gram(isx, shape)
    else:
        test_conv3x3_dgrad(isx, shape, mode)


def test_conv_sweep_matches_c64_bitwise(isx):
    """Same K order per output element (nine taps x four 16-channel steps, fp32 accumulation in TMEM) is not guaranteed between
    the two kernels, but both must agree to one bf16 rounding; and the sweep kernel must be deterministic run to run."""
    B, H, W = 2, 256, 120
    x = nhwc_bf16(B, H, W, 64, 5, relu=True)
    w = torch.randn(64, 64, 3, 3, device="cuda") * (2.0 / (9 * 64)) ** 0.5
    bias = torch.randn(64, device="cuda") * 0.1
    wf, _ = pack(isx, w)
    lib = isx.load()
    outs = []
    for opt in (0, 2, 2):
        assert lib.isx_set_option(b"sweep64", opt) == 0
        o = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
        isx.call("isx_conv3x3_bias_relu_fwd", x, wf, bias, o, B, H, W, 64, 64, 1, 0, isx.stream_ptr())
        outs.append(o)
    assert lib.isx_set_option(b"sweep64", 1) == 0
    torch.cuda.synchronize()
    assert torch.equal(outs[1], outs[2])
    assert_close_bf16(outs[1], outs[0].float(), "sweep vs c64")


# ---- the halo-patch pair kernel (conv_halo.cu): persistent CTAs, 16x16 / 8x32 pixel pairs, streamed weight slabs ------
HALO_SHAPES = [
    (2, 20, 24, 64, 64),       # BN = 64, fewer items than SMs
    (3, 13, 9, 128, 128),      # ragged in x and y; second tile of the pair entirely outside the image
    (1, 50, 80, 128, 256),     # two Cout tiles
    (2, 37, 41, 256, 128),     # odd sizes (fused pool drops the last row/column), four input blocks
    (2, 100, 200, 64, 128),    # 16x16 pairs, several items per CTA: rings and both accumulator sets wrap
    (3, 160, 100, 128, 64),    # 8x32 pairs (100 = 12.5 x 8), BN = 64
    (1, 25, 40, 512, 512),
]


@pytest.fixture
def halo2_forced(isx):
    lib = isx.load()
    assert lib.isx_set_option(b"halo2", 2) == 0
    assert lib.isx_set_option(b"c64", 0) == 0
    assert lib.isx_set_option(b"sweep64", 0) == 0
    yield
    assert lib.isx_set_option(b"halo2", 1) == 0
    assert lib.isx_set_option(b"c64", 1) == 0
    assert lib.isx_set_option(b"sweep64", 1) == 0


@pytest.mark.parametrize("shape", HALO_SHAPES)
def test_conv_halo_fwd_and_pool(isx, halo2_forced, shape):
    test_conv3x3_fwd(isx, shape, 0)
    test_conv3x3_fwd_fused_pool(isx, shape, 0)


@pytest.mark.parametrize("shape", HALO_SHAPES)
@pytest.mark.parametrize("mode", ["plain", "mask", "mask_add", "mask_affine", "gram"])
def test_conv_halo_dgrad(isx, halo2_forced, shape, mode):
    if mode == "gram":
        test_conv3x3_dgrad_fused_gram(isx, shape)
    else:
        test_conv3x3_dgrad(isx, shape, mode)


@pytest.mark.parametrize("xc", [3, 1])
@pytest.mark.parametrize("use_mask", [False, True])
def test_conv1_1_generic_tail(isx, xc, use_mask):
    """The 36-MMA image-gradient tail of the generic kernel (the default is the taps-in-N kernel, conv1_1_tail.cu)."""
    lib = isx.load()
    assert lib.isx_set_option(b"tail_n", 0) == 0
    try:
        test_conv1_1_fwd_dgrad(isx, xc, use_mask)
    finally:
        assert lib.isx_set_option(b"tail_n", 1) == 0


@pytest.mark.parametrize("B,H,W", [(1, 14, 14), (2, 15, 29), (3, 100, 57), (2, 200, 320)])
@pytest.mark.parametrize("xc", [3, 1])
def test_conv1_1_tail_shapes(isx, B, H, W, xc):
    """Taps-in-N tail on exact, ragged and multi-patch-per-CTA shapes against autograd of the fp32 convolution."""
    w = torch.randn(64, 3, 3, 3, device="cuda") * (2.0 / (9 * 64)) ** 0.5
    mask = (torch.rand(B, 1, H, W, device="cuda") > 0.3).float()
    dy = nhwc_bf16(B, H, W, 64, 7)
    wd0 = torch.empty(9, 16, 64, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_pack_conv1_1_dgrad", w, wd0, isx.stream_ptr())
    dx = torch.full((B, xc, H, W), float("nan"), device="cuda")
    isx.call("isx_conv1_1_dgrad_tc", dy, wd0, mask, B, dx, xc, B, H, W, isx.stream_ptr())
    torch.cuda.synchronize()
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    x = torch.rand(B, xc, H, W, device="cuda", requires_grad=True)
    F.conv2d((x - mean) / std * mask, w.to(torch.bfloat16).float(), None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx, x.grad, rtol=1e-4, atol=1e-4 * x.grad.abs().max().item())


def test_persistent_conv_kernels_are_deterministic(isx):
    """conv_halo / conv_c64 / conv1_1_tail with ~10 work items per CTA (every ring, TMEM set and staging slot wraps several
    times): three runs must agree bit for bit -- a missed barrier shows up as run-to-run noise."""
    B, H, W = 6, 200, 160
    outs = []
    x128 = nhwc_bf16(B, H, W, 128, 21, relu=True)
    x64 = nhwc_bf16(B, H, W, 64, 22, relu=True)
    dy = nhwc_bf16(B, H, W, 128, 23)
    w = torch.randn(128, 128, 3, 3, device="cuda") * 0.03
    wf, wd = pack(isx, w)
    w64 = torch.randn(64, 64, 3, 3, device="cuda") * 0.04
    wf64, wd64 = pack(isx, w64)
    bias = torch.randn(128, device="cuda") * 0.1
    D = (torch.randn(B, 128, 128, device="cuda") * 0.05).to(torch.bfloat16)
    D64 = (torch.randn(B, 64, 64, device="cuda") * 0.05).to(torch.bfloat16)
    w0 = torch.randn(64, 3, 3, 3, device="cuda") * 0.1
    wd0 = torch.empty(9, 16, 64, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_pack_conv1_1_dgrad", w0, wd0, isx.stream_ptr())
    for rep in range(3):
        o1 = torch.empty(B, H, W, 128, device="cuda", dtype=torch.bfloat16)
        p1 = torch.empty(B, H // 2, W // 2, 128, device="cuda", dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        o3 = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
        o4 = torch.empty_like(o3)
        o5 = torch.empty(B, 3, H, W, device="cuda")
        isx.call("isx_conv3x3_bias_relu_pool_fwd", x128, wf, bias, o1, p1, B, H, W, 128, 128, 0, isx.stream_ptr())
        isx.call("isx_conv3x3_dgrad_gram", dy, wd, o2, B, H, W, 128, 128, x128, D, isx.stream_ptr())
        isx.call("isx_conv3x3_bias_relu_fwd", x64, wf64, bias[:64].contiguous(), o3, B, H, W, 64, 64, 1, 0, isx.stream_ptr())
        isx.call("isx_conv3x3_dgrad_gram", x64, wd64, o4, B, H, W, 64, 64, x64, D64, isx.stream_ptr())
        isx.call("isx_conv1_1_dgrad_tc", x64, wd0, None, 0, o5, 3, B, H, W, isx.stream_ptr())
        torch.cuda.synchronize()
        outs.append((o1, p1, o2, o3, o4, o5))
    for later in outs[1:]:
        for a, b in zip(outs[0], later):
            assert torch.equal(a, b)


def test_conv_c64_matches_generic_kernel(isx):
    """Same inputs through the generic tcgen05 kernel and through conv_c64: the MMAs run in the same order
    (tap-major, then the Gram block), so the bf16 results must be identical."""
    lib = isx.load()
    B, H, W = 3, 48, 56
    x = nhwc_bf16(B, H, W, 64, 11, relu=True)
    w = torch.randn(64, 64, 3, 3, device="cuda") * (2.0 / (9 * 64)) ** 0.5
    bias = torch.randn(64, device="cuda") * 0.1
    wf, wd = pack(isx, w)
    act = nhwc_bf16(B, H, W, 64, 3, relu=True)
    D = torch.randn(B, 64, 64, device="cuda") * 0.05
    D = (D + D.transpose(1, 2)).to(torch.bfloat16).contiguous()
    res = {}
    try:
        for mode in (0, 2):
            assert lib.isx_set_option(b"c64", mode) == 0
            out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
            pool = torch.empty(B, H // 2, W // 2, 64, device="cuda", dtype=torch.bfloat16)
            dx = torch.empty_like(out)
            isx.call("isx_conv3x3_bias_relu_pool_fwd", x, wf, bias, out, pool, B, H, W, 64, 64, 0, isx.stream_ptr())
            isx.call("isx_conv3x3_dgrad_gram", x, wd, dx, B, H, W, 64, 64, act, D, isx.stream_ptr())
            torch.cuda.synchronize()
            res[mode] = (out, pool, dx)
    finally:
        lib.isx_set_option(b"c64", 1)
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("C", [64, 128, 256, 512])
@pytest.mark.parametrize("B,H,W", [(1, 50, 80), (3, 17, 23), (2, 100, 160)])
def test_gram_fwd_and_loss(isx, C, B, H, W):
    f = nhwc_bf16(B, H, W, C, 5, relu=True)
    HW = H * W
    inv_n = 1.0 / (C * HW)
    ws = torch.empty(isx.call_i64("isx_gram_workspace_bytes", B, HW, C), device="cuda", dtype=torch.uint8)
    G = torch.full((B, C, C), float("nan"), device="cuda")
    tgt = torch.randn(B, C, C, device="cuda") * 0.01
    loss = torch.zeros(B, device="cuda", dtype=torch.float64)
    D = torch.empty(B, C, C, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_gram_fwd", f, B, HW, C, isx.f32(inv_n), ws, G, tgt, B, isx.f64(0.25), loss, isx.f32(3.0), D,
             isx.stream_ptr())
    torch.cuda.synchronize()
    ff = f.float().reshape(B, HW, C)
    ref = torch.bmm(ff.transpose(1, 2), ff) * inv_n
    assert torch.allclose(G, ref, rtol=2e-4, atol=1e-6 * ref.abs().max().item()), (G - ref).abs().max().item()
    ref_loss = 0.25 * ((ref - tgt).double() ** 2).sum(dim=(1, 2))
    assert torch.allclose(loss, ref_loss, rtol=1e-4)
    assert_close_bf16(D, 3.0 * (ref - tgt), "gram D")
    # single shared target (style batch 1)
    loss.zero_()
    isx.call("isx_gram_fwd", f, B, HW, C, isx.f32(inv_n), ws, None, tgt[:1].contiguous(), 1, isx.f64(0.25), loss,
             isx.f32(0.0), None, isx.stream_ptr())
    ref_loss1 = 0.25 * ((ref - tgt[:1]).double() ** 2).sum(dim=(1, 2))
    assert torch.allclose(loss, ref_loss1, rtol=1e-4)


@pytest.mark.parametrize("C", [64, 128, 256, 512])
@pytest.mark.parametrize("B,H,W,mb", [(2, 50, 80, 2), (3, 17, 23, 1), (1, 100, 160, 1)])
def test_gram_masked_fwd(isx, C, B, H, W, mb):
    """Row G': the mask weights are applied to the operand tiles INSIDE the Gram kernel (all-zero K blocks skipped) and
    F*m^2 is written for the backward -- bit-identical to the separate pre-pass isx_mask_features + isx_gram_fwd, and
    equal to utils.GramMatrix(F * m) computed by torch."""
    f = nhwc_bf16(B, H, W, C, 9, relu=True)
    HW = H * W
    g = torch.Generator().manual_seed(3)
    m = torch.zeros(mb, H, W)
    m[:, H // 3: H // 3 + max(2, H // 4), W // 4: W // 4 + max(3, W // 3)] = 1.0            # a blob: most K blocks are empty
    m[:, H // 3 + 1, W // 4: W // 4 + 3] = torch.tensor([0.25, 0.5, 0.75])                  # pooled-mask style fractions
    if mb > 1:
        m[1] = (torch.rand(H, W, generator=g) > 0.5).float()                               # dense random mask
    m = m.reshape(mb, HW).cuda().contiguous()
    inv_n = 1.0 / (C * HW)
    ws = torch.empty(isx.call_i64("isx_gram_workspace_bytes", B, HW, C), device="cuda", dtype=torch.uint8)
    fl = torch.empty(max(16, isx.call_i64("isx_gram_mask_flags_bytes", mb, HW, C)), device="cuda", dtype=torch.uint8)
    fm2 = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    G = torch.full((B, C, C), float("nan"), device="cuda")
    tgt = torch.randn(B, C, C, device="cuda") * 0.01
    loss = torch.zeros(B, device="cuda", dtype=torch.float64)
    D = torch.empty(B, C, C, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_gram_masked_fwd", f, B, HW, C, m, mb, fl, fm2, isx.f32(inv_n), ws, G, tgt, B, isx.f64(0.25), loss,
             isx.f32(3.0), D, isx.stream_ptr())
    # the pre-pass formulation
    fm_ref = torch.empty_like(f)
    fm2_ref = torch.empty_like(f)
    isx.call("isx_mask_features", f, m, mb, fm_ref, fm2_ref, B, isx.i64(HW), C, isx.stream_ptr())
    G2 = torch.empty_like(G)
    loss2 = torch.zeros_like(loss)
    isx.call("isx_gram_fwd", fm_ref, B, HW, C, isx.f32(inv_n), ws, G2, tgt, B, isx.f64(0.25), loss2, isx.f32(3.0), None,
             isx.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(G, G2) and torch.equal(fm2, fm2_ref)
    assert torch.allclose(loss, loss2, rtol=1e-12)
    ff = (f.float().reshape(B, HW, C) * m.reshape(mb, HW, 1)).to(torch.bfloat16).float()
    ref = torch.bmm(ff.transpose(1, 2), ff) * inv_n
    assert torch.allclose(G, ref, rtol=2e-4, atol=1e-6 * ref.abs().max().item())
    assert torch.equal(G, G.transpose(1, 2))
    assert_close_bf16(D, 3.0 * (ref - tgt), "masked gram D")


@pytest.mark.parametrize("C", [64, 128, 256, 512])
def test_gram_bwd(isx, C):
    B, H, W = 2, 26, 40
    f = nhwc_bf16(B, H, W, C, 6, relu=True)
    Dm = torch.randn(B, C, C, device="cuda") * 0.1
    Dm = (Dm + Dm.transpose(1, 2)).to(torch.bfloat16).contiguous()
    dF = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_gram_bwd", f, Dm, dF, B, H, W, C, None, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = torch.bmm(f.float().reshape(B, H * W, C), Dm.float()).reshape(B, H, W, C)
    assert_close_bf16(dF, ref, "gram bwd C=%d" % C)
    isx.call("isx_gram_bwd", f, Dm, dF, B, H, W, C, f, isx.stream_ptr())
    torch.cuda.synchronize()
    assert_close_bf16(dF, torch.where(f.float() > 0, ref, torch.zeros_like(ref)), "gram bwd masked C=%d" % C)


@pytest.mark.parametrize("xc", [3, 1])
@pytest.mark.parametrize("use_mask", [False, True])
def test_conv1_1_fwd_dgrad(isx, xc, use_mask):
    B, H, W = 2, 21, 30
    x = torch.rand(B, xc, H, W, device="cuda")
    w = torch.randn(64, 3, 3, 3, device="cuda") * (2.0 / (9 * 64)) ** 0.5
    bias = torch.randn(64, device="cuda") * 0.1
    mask = (torch.rand(B, 1, H, W, device="cuda") > 0.3).float() if use_mask else None
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_conv1_1_fwd", x, xc, mask, B, w, bias, out, B, H, W, isx.stream_ptr())
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    xr = x.clone().requires_grad_(True)
    xn = (xr - mean) / std
    if use_mask:
        xn = xn * mask
    y = F.relu(F.conv2d(xn, w, bias, padding=1))
    assert_close_bf16(out, y.detach().permute(0, 2, 3, 1), "conv1_1 fwd")
    # tensor-core head: bf16 weights, input split into bf16 hi + lo
    w0p = torch.empty(64, 64, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_pack_conv1_1_fwd", w, w0p, isx.stream_ptr())
    out2 = torch.full_like(out, float("nan"))
    isx.call("isx_conv1_1_fwd_tc", x, xc, mask, B, w0p, bias, out2, B, H, W, isx.stream_ptr())
    torch.cuda.synchronize()
    y2 = F.relu(F.conv2d(xn.detach(), w.to(torch.bfloat16).float(), bias, padding=1))
    assert_close_bf16(out2, y2.permute(0, 2, 3, 1), "conv1_1 fwd (tensor cores)")
    dy = nhwc_bf16(B, H, W, 64, 7)
    y.backward(dy.float().permute(0, 3, 1, 2) * (y > 0))  # dy is "already ReLU-masked" in the kernel contract
    dyk = (dy.float() * (y.detach().permute(0, 2, 3, 1) > 0)).to(torch.bfloat16).contiguous()
    dx = torch.empty(B, xc, H, W, device="cuda")
    isx.call("isx_conv1_1_dgrad", dyk, w, mask, B, dx, xc, B, H, W, isx.stream_ptr())
    torch.cuda.synchronize()
    ref = torch.autograd.grad(F.conv2d(xn, w, bias, padding=1), xr, dyk.float().permute(0, 3, 1, 2))[0] if False else None
    # recompute the reference gradient from the bf16-rounded masked dy
    xr2 = x.clone().requires_grad_(True)
    xn2 = (xr2 - mean) / std
    if use_mask:
        xn2 = xn2 * mask
    F.conv2d(xn2, w, bias, padding=1).backward(dyk.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx, xr2.grad, rtol=1e-4, atol=1e-4 * xr2.grad.abs().max().item())
    # the same tail on the tensor cores (bf16-rounded weights, fp32 accumulation)
    wd0 = torch.empty(9, 16, 64, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_pack_conv1_1_dgrad", w, wd0, isx.stream_ptr())
    dx2 = torch.full_like(dx, float("nan"))
    isx.call("isx_conv1_1_dgrad_tc", dyk, wd0, mask, B, dx2, xc, B, H, W, isx.stream_ptr())
    torch.cuda.synchronize()
    xr3 = x.clone().requires_grad_(True)
    xn3 = (xr3 - mean) / std
    if use_mask:
        xn3 = xn3 * mask
    F.conv2d(xn3, w.to(torch.bfloat16).float(), bias, padding=1).backward(dyk.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx2, xr3.grad, rtol=1e-4, atol=1e-4 * xr3.grad.abs().max().item())


@pytest.mark.parametrize("B,H,W,C", [(2, 20, 24, 64), (1, 25, 41, 128), (3, 8, 6, 512)])
def test_maxpool_fwd_bwd(isx, B, H, W, C):
    a = nhwc_bf16(B, H, W, C, 8, relu=True)
    out = torch.empty(B, H // 2, W // 2, C, device="cuda", dtype=torch.bfloat16)
    isx.call("isx_maxpool2x2_fwd", a, out, B, H, W, C, isx.stream_ptr())
    ar = a.float().permute(0, 3, 1, 2).requires_grad_(True)
    pr = F.max_pool2d(ar, 2, 2)
    assert torch.equal(out.float(), pr.detach().permute(0, 2, 3, 1))
    dy = nhwc_bf16(B, H // 2, W // 2, C, 9)
    dx = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_maxpool2x2_bwd", dy, a, dx, B, H, W, C, isx.stream_ptr())
    torch.cuda.synchronize()
    pr.backward(dy.float().permute(0, 3, 1, 2))
    ref = (ar.grad * (ar.detach() > 0)).permute(0, 2, 3, 1)
    assert torch.equal(dx.float(), ref)
    # the same pair through routing bytes (what the fused conv epilogues emit): bit-identical results
    out2 = torch.empty_like(out)
    idx = torch.full((B, H // 2, W // 2, C), 255, device="cuda", dtype=torch.uint8)
    isx.call("isx_maxpool2x2_fwd_idx", a, out2, idx, B, H, W, C, isx.stream_ptr())
    dx2 = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    isx.call("isx_maxpool2x2_bwd_idx", dy, idx, dx2, B, H, W, C, isx.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out2, out) and int(idx.max()) <= 4
    assert torch.equal(dx2, dx)
    # ties and zero windows: first maximum wins, an all-zero window routes nothing
    t = torch.zeros(1, 4, 4, 8, device="cuda", dtype=torch.bfloat16)
    t[0, 0, 1, :] = 2.0
    t[0, 1, 0, :] = 2.0          # window (0,0): tie between positions 1 and 2 -> 1
    t[0, 2, 2, 0] = 1.0          # window (1,1), channel 0: position 0; other channels all zero -> 4
    ti = torch.empty(1, 2, 2, 8, device="cuda", dtype=torch.uint8)
    isx.call("isx_maxpool2x2_fwd_idx", t, None, ti, 1, 4, 4, 8, isx.stream_ptr())
    torch.cuda.synchronize()
    assert ti[0, 0, 0].tolist() == [1] * 8 and ti[0, 1, 1].tolist() == [0] + [4] * 7 and ti[0, 0, 1].tolist() == [4] * 8


def test_content_mse(isx):
    B, n = 3, 512 * 50 * 80
    p = nhwc_bf16(B, 50, 80, 512, 10, relu=True)
    t = nhwc_bf16(B, 50, 80, 512, 11, relu=True)
    g = torch.empty_like(p)
    loss = torch.zeros(B, device="cuda", dtype=torch.float64)
    isx.call("isx_content_mse_fwd_bwd", p, t, B, g, B, isx.i64(n), isx.f64(0.5 / n), isx.f32(1.0 / n), loss,
             isx.stream_ptr())
    torch.cuda.synchronize()
    d = p.float() - t.float()
    ref_loss = 0.5 * (d.double() ** 2).reshape(B, -1).mean(dim=1)
    assert torch.allclose(loss, ref_loss, rtol=1e-5)
    assert_close_bf16(g, d / n * (p.float() > 0), "content grad")


@pytest.mark.parametrize("C", [64, 128, 256, 512])
def test_bn_stats(isx, C):
    B, H, W = 3, 37, 52
    f = nhwc_bf16(B, H, W, C, 12, relu=True)
    HW = H * W
    sums = torch.empty(B, C, 2, device="cuda", dtype=torch.float64)
    mean = torch.empty(B, C, device="cuda")
    std = torch.empty(B, C, device="cuda")
    tm = torch.rand(B, C, device="cuda")
    ts = torch.rand(B, C, device="cuda")
    loss = torch.zeros(B, device="cuda", dtype=torch.float64)
    aa = torch.empty(B, C, device="cuda")
    ab = torch.empty(B, C, device="cuda")
    w_l, beta = 1.0, 1e4
    isx.call("isx_bn_stats_fwd", f, B, isx.i64(HW), C, sums, mean, std, tm, ts, B, isx.f64(w_l / C),
             isx.f64(beta * w_l / C), loss, aa, ab, isx.stream_ptr())
    torch.cuda.synchronize()
    fr = f.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    m_ref, s_ref = fr.mean(dim=(-2, -1)), fr.std(dim=(-2, -1))
    assert torch.allclose(mean, m_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(std, s_ref.detach(), rtol=1e-4, atol=1e-6)
    l_ref = (((m_ref - tm) ** 2 + (s_ref - ts) ** 2).sum(dim=1) * w_l / C)
    assert torch.allclose(loss, l_ref.detach().double(), rtol=1e-4)
    (l_ref.sum() * beta).backward()
    g_ref = fr.grad.permute(0, 2, 3, 1)
    g_got = aa[:, None, None, :] + ab[:, None, None, :] * f.float()
    assert torch.allclose(g_got, g_ref, rtol=2e-3, atol=2e-3 * g_ref.abs().max().item())


def test_handles_isolate_options_and_counters(isx):
    """isx_create / isx_make_current / isx_destroy: options, launch counter and profiler belong to a handle, not to the
    process; a thread without a bound handle uses its private default context."""
    lib = isx.load()
    lib.isx_launch_count.restype = ctypes.c_ulonglong

    def opt(name):
        v = ctypes.c_int(-1)
        assert lib.isx_get_option(name, ctypes.byref(v)) == 0
        return v.value

    assert opt(b"c64") == 1 and lib.isx_sm_count() == torch.cuda.get_device_properties(0).multi_processor_count
    base = lib.isx_launch_count()
    h1, h2 = isx.Handle(0), isx.Handle(0)
    x = torch.rand(1000, device="cuda")
    with h1:
        assert lib.isx_set_option(b"c64", 0) == 0 and opt(b"c64") == 0
        assert lib.isx_launch_count() == 0
        isx.call("isx_clamp01", x, isx.i64(x.numel()), isx.stream_ptr())
        assert lib.isx_launch_count() == 1
    with h2:
        assert opt(b"c64") == 1 and lib.isx_launch_count() == 0
    assert opt(b"c64") == 1 and lib.isx_launch_count() == base        # the default context saw none of it
    assert lib.isx_set_option(b"no_such_option", 1) != 0
    bad = ctypes.c_void_p()
    assert lib.isx_create(99, ctypes.byref(bad)) != 0                   # no such device
    h1.destroy(); h2.destroy()
