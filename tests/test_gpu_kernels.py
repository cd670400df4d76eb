"""Parity of the individual CUDA kernels (through the C ABI) against fp32 restatements of the reference ops.
Integer work (window maps, im2col, taps) is bit-exact; floating point within the stated tolerances."""
import pytest
import torch
import torch.nn.functional as F

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from anyref_b200 import ops as _ops

    return _ops


def rel_fro(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm()).item()


# ------------------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("dt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 192), (4900, 3840, 1280), (4096, 1280, 5120),
                                   (4096, 256, 2304), (513, 480, 160), (1000, 160, 640)])
def test_gemm_epilogues(ops, dt, tol, M, N, K):
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV)
    ref = a.float() @ w.float().t()
    # fp32 output: only accumulation-order differences
    assert rel_fro(ops.gemm(a, w, out_dtype=torch.float32), ref) < 1e-5
    # bias + exact-erf GELU, rounded to the operand format (common.py:21-26)
    got = ops.gemm(a, w, bias=bias, act="gelu", out_dtype=dt)
    want = F.gelu(ref + bias)
    assert (got.float() - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    # in-place fp32 residual (image_encoder.py:190-191)
    res = torch.randn(M, N, device=DEV)
    x = res.clone()
    ops.gemm(a, w, bias=bias, residual=x, out=x)
    assert rel_fro(x, ref + bias + res) < 1e-5
    # broadcast residual rows (pos_embed add, image_encoder.py:112-113)
    pos = torch.randn(128, N, device=DEV)
    got = ops.gemm(a, w, bias=bias, residual=pos, res_mod=128, out_dtype=torch.float32)
    assert rel_fro(got, ref + bias + pos[torch.arange(M, device=DEV) % 128]) < 1e-5


# ------------------------------------------------------------------------------------------ LayerNorm folding
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_cast_stats(ops, dt):
    torch.manual_seed(3)
    x = torch.randn(4096 + 24, 1280, device=DEV) * 2 + 0.3
    xb, stats = ops.cast_stats(x, dt)
    assert torch.equal(xb, x.to(dt))                                           # rounding is bit-exact
    xs = x.double().view(x.shape[0], 10, 128)                                  # per 128-column slice: (mean, M2)
    m2 = ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)
    assert (stats[..., 0].double() - xs.mean(-1)).abs().max().item() < 1e-5
    assert ((stats[..., 1].double() - m2).abs() / m2).max().item() < 1e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 1280, 1280), (8192, 1280, 5120), (2048 + 32, 256, 256), (65536, 1280, 1280)])
def test_gemm_residual_ln_producer(ops, dt, M, N, K):
    """x += a.W^T + b in place; xb == round(x) bit-exactly; slice statistics (mean, M2) of the NEW x (image_encoder.py:190-192)."""
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV)
    x0 = torch.randn(M, N, device=DEV) * 2 + 0.1
    x = x0.clone()
    xb, stats = ops.gemm_residual_ln(a, w, x, bias)
    ref = a.float() @ w.float().t() + bias + x0
    assert rel_fro(x, ref) < 1e-5
    assert torch.equal(xb, x.to(dt))
    xs = x.double().view(M, N // 128, 128)
    m2 = ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)
    assert (stats[..., 0].double() - xs.mean(-1)).abs().max().item() < 2e-5
    assert ((stats[..., 1].double() - m2).abs() / m2).max().item() < 2e-5
    # same numbers as the plain in-place residual epilogue (TMA reduce-add) up to the last fp32 bit
    y = x0.clone()
    ops.gemm(a, w, bias=bias, residual=y, out=y)
    assert rel_fro(x, y) < 1e-6


@pytest.mark.parametrize("dt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("M,N,K,act", [(4096, 3840, 1280, "none"), (4096, 5120, 1280, "gelu"), (1024 + 32, 768, 256, "none")])
def test_gemm_ln_consumer_matches_layernorm_then_linear(ops, dt, tol, M, N, K, act):
    """act(LN(x).W^T + b) (image_encoder.py:178-179 -> :238, :191 -> common.py:26) from the folded operands."""
    import types

    from anyref_b200.segment_anything._pack import _fold_layernorm

    torch.manual_seed(M + N)
    x = torch.randn(M, K, device=DEV) * 2 + 0.2 * torch.randn(M, 1, device=DEV)
    gamma = torch.rand(K, device=DEV) * 0.2 + 0.9
    beta = torch.rand(K, device=DEV) * 0.2 - 0.1
    W = (torch.rand(N, K, device=DEV) * 2 - 1) / K ** 0.5
    b = (torch.rand(N, device=DEV) * 2 - 1) / K ** 0.5
    norm = types.SimpleNamespace(weight=gamma, bias=beta)
    lin = types.SimpleNamespace(weight=W, bias=b)
    wg, colsum, bias_fold = _fold_layernorm(norm, lin, dt)
    xb, stats = ops.cast_stats(x, dt)
    got = ops.gemm_ln(xb, wg, bias_fold, colsum, stats, 1e-6, act=act)
    want = F.layer_norm(x.double(), (K,), gamma.double(), beta.double(), 1e-6) @ W.double().t() + b.double()
    if act == "gelu":
        want = F.gelu(want)
    assert rel_fro(got, want.float()) < tol
    # and it is as accurate as the unfused pair LayerNorm kernel -> GEMM
    a16 = ops.layernorm(x, gamma, beta, 1e-6, dt)
    plain = ops.gemm(a16, W.to(dt), bias=b, act=act, out_dtype=dt)
    assert rel_fro(got, want.float()) < 1.5 * rel_fro(plain, want.float()) + 1e-4


def test_fp16_stores_saturate_instead_of_overflowing(ops):
    """fp16 operand mode: a result beyond +-65504 is stored as the largest finite half, not inf (the reference's own fp16
    path would carry the inf into NaNs); in-range values are untouched (every other fp16 test is bit-sensitive to that)."""
    M, K, N = 256, 64, 256
    a = torch.full((M, K), 64.0, device=DEV, dtype=torch.float16)
    w = torch.full((N, K), 32.0, device=DEV, dtype=torch.float16)
    w[N // 2:] = -32.0
    out = ops.gemm(a, w, out_dtype=torch.float16)            # +-131072 in fp32
    assert bool(torch.isfinite(out.float()).all())
    assert torch.equal(out[:, : N // 2].float(), torch.full((M, N // 2), 65504.0, device=DEV))
    assert torch.equal(out[:, N // 2:].float(), torch.full((M, N // 2), -65504.0, device=DEV))
    x = torch.full((128, 1280), 1e6, device=DEV)
    xb, _ = ops.cast_stats(x, torch.float16)
    assert torch.equal(xb.float(), torch.full_like(x, 65504.0))


def _stats_to_mean_var(stats, C):
    """Chan combination of the per-slice (mean, M2) statistics [M, C/128, 2] -> (mean [M], biased variance [M])."""
    mean_i, m2_i = stats[..., 0].double(), stats[..., 1].double()
    mean = mean_i.mean(dim=1)
    m2 = m2_i.sum(dim=1) + 128.0 * ((mean_i - mean[:, None]) ** 2).sum(dim=1)
    return mean, m2 / C


@pytest.mark.parametrize("offset", [0.0, 50.0, 1000.0])
def test_ln_fold_statistics_survive_a_large_row_mean(ops, offset):
    """Rows whose |mean| is up to 1000x their standard deviation: the folded LayerNorm's statistics -- per 128-column
    slice (mean, M2) from cast_stats and from the residual-GEMM producer epilogue, combined with Chan's formula in the
    consumer -- must still give the row variance to 1e-4 (E[x^2] - mean^2 in fp32 is off by several per cent at 1000),
    and the consumer GEMM must match the same formula evaluated in fp64 on the SAME rounded operand."""
    torch.manual_seed(17)
    M, K, N = 2048, 1280, 1280
    dt = torch.float16
    x = torch.randn(M, K, device=DEV) + offset * (1.0 + torch.rand(M, 1, device=DEV))
    mean_ref, var_ref = x.double().mean(1), x.double().var(1, unbiased=False)
    xb, stats = ops.cast_stats(x.clone(), dt)
    mean, var = _stats_to_mean_var(stats, K)
    assert ((mean - mean_ref).abs() / (mean_ref.abs() + 1.0)).max().item() < 1e-6
    assert ((var - var_ref).abs() / var_ref).max().item() < 1e-4
    # producer epilogue: x <- x + a.w^T + b, statistics of the NEW rows
    a = (torch.randn(M, 256, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, 256, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV) * 0.1
    x2 = x.clone()
    xb2, stats2 = ops.gemm_residual_ln(a, w, x2, bias)
    mean2, var2 = _stats_to_mean_var(stats2, N)
    assert ((mean2 - x2.double().mean(1)).abs() / (x2.double().mean(1).abs() + 1.0)).max().item() < 1e-6
    assert ((var2 - x2.double().var(1, unbiased=False)).abs() / x2.double().var(1, unbiased=False)).max().item() < 1e-4
    # consumer: rstd * (xb . W'^T - mean * colsum) + bias'  against fp64 with the exact statistics of x
    import types
    from anyref_b200.segment_anything._pack import _fold_layernorm

    gamma = torch.rand(K, device=DEV) * 0.2 + 0.9
    beta = torch.rand(K, device=DEV) * 0.2 - 0.1
    W = (torch.rand(256, K, device=DEV) * 2 - 1) / K ** 0.5
    b = (torch.rand(256, device=DEV) * 2 - 1) / K ** 0.5
    wg, colsum, bias_fold = _fold_layernorm(types.SimpleNamespace(weight=gamma, bias=beta),
                                            types.SimpleNamespace(weight=W, bias=b), dt)
    got = ops.gemm_ln(xb, wg, bias_fold, colsum, stats, 1e-6).float()
    rstd = 1.0 / torch.sqrt(var_ref + 1e-6)
    want = rstd[:, None] * (xb.double() @ wg.double().t() - mean_ref[:, None] * colsum.double()[None]) + bias_fold.double()
    # the products xb.W' and mean.colsum are ~offset times larger than their difference: allow their fp32 rounding
    tol = 2e-3 * (1.0 + offset / 10.0)
    assert ((got.double() - want).norm() / want.norm()).item() < tol


def _guarded(shape, dtype, pad=4096):
    """Output view in the middle of a sentinel-filled allocation (compute-sanitizer is closed on the GPU pool)."""
    n = 1
    for d in shape:
        n *= d
    raw = torch.full((n + 2 * pad,), 77, device=DEV, dtype=torch.float32).to(dtype)
    return raw, raw[pad:pad + n].view(shape), pad


def _guards_intact(raw, pad):
    return bool((raw[:pad].float() == 77).all()) and bool((raw[-pad:].float() == 77).all())


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (513, 480, 160), (4096, 1280, 5120), (4900, 3840, 1280), (16384 + 256, 1280, 320)])
def test_gemm_half_tiles_of_the_last_round_are_bit_identical(ops, dt, M, N, K):
    """An underfull last round of 256 x 256 tiles is cut into 256 x 128 halves (gemm2.cu, Sched): every output column is
    still accumulated in the same order, so each epilogue must return the same bits with the split forced on and off."""
    from anyref_b200 import _lib
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    got = {}
    try:
        for mode in (0, 1):
            _lib.gemm_set_tile_split(mode)
            x = res.clone()
            ops.gemm(a, w, bias=bias, residual=x, out=x)
            got[mode] = (ops.gemm(a, w, out_dtype=torch.float32), ops.gemm(a, w, bias=bias, act="gelu", out_dtype=dt), x)
    finally:
        _lib.gemm_set_tile_split(-1)
    for u, v in zip(got[0], got[1]):
        assert torch.equal(u, v)
    assert rel_fro(got[1][0], a.float() @ w.float().t()) < 1e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4096, 1280, 1280), (4096, 1280, 5120), (2048 + 32, 256, 256), (16384 + 256, 1280, 320)])
def test_ln_fold_with_half_tiles(ops, dt, M, N, K):
    """LayerNorm folding through half tiles: the residual producer's x / xb / slice statistics (two 64-column halves, held
    by the two warps of a lane quadrant instead of one warp) are bit-identical to the whole-tile schedule and agree with
    fp64, and the consumer returns the same bits."""
    from anyref_b200 import _lib
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV)
    x0 = torch.randn(M, N, device=DEV) * 2 + 3.0
    colsum = torch.randn(N, device=DEV)
    st_in = torch.rand(M, K // 128, 2, device=DEV) if K % 128 == 0 else None     # (mean, M2) per 128-column slice
    got = {}
    try:
        for mode in (0, 1):
            _lib.gemm_set_tile_split(mode)
            x = x0.clone()
            xb, stats = ops.gemm_residual_ln(a, w, x, bias)
            y = ops.gemm_ln(a, w, bias, colsum, st_in, 1e-6, act="gelu") if st_in is not None else None
            got[mode] = (x, xb, stats, y)
    finally:
        _lib.gemm_set_tile_split(-1)
    assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1])
    if got[0][3] is not None:
        assert torch.equal(got[0][3], got[1][3])
    x, _, stats, _ = got[1]
    xs = x.double().view(M, N // 128, 128)
    m2 = ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)
    assert (stats[..., 0].double() - xs.mean(-1)).abs().max().item() < 2e-5
    assert ((stats[..., 1].double() - m2).abs() / m2).max().item() < 2e-5
    # a slice is always the Chan combination of its two 64-column halves, whoever computed them: same bits either way
    assert torch.equal(got[0][2], stats)


@pytest.mark.parametrize("M,N,K", [(2048 + 32, 256, 256), (4096, 1280, 1280), (288, 768, 2304)])
def test_ln_fold_kernels_stay_inside_their_outputs(ops, M, N, K):
    """Guard bands around every output of the folded-LayerNorm kernels (M not a multiple of the 256-row tile, K long
    enough for the vector-load x path): nothing outside the [M, *] views is written."""
    dt = torch.bfloat16
    torch.manual_seed(7)
    a = (torch.randn(M, K, device=DEV) * 0.5).to(dt)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(dt)
    bias = torch.randn(N, device=DEV)
    raw_x, x, pad = _guarded((M, N), torch.float32)
    x.copy_(torch.randn(M, N, device=DEV))
    x0 = x.clone()
    raw_xb, xb, _ = _guarded((M, N), dt)
    raw_st, st, _ = _guarded((M, N // 128, 2), torch.float32)
    ops.gemm_residual_ln(a, w, x, bias, xb=xb, stats=st)
    torch.cuda.synchronize()
    assert _guards_intact(raw_x, pad) and _guards_intact(raw_xb, pad) and _guards_intact(raw_st, pad)
    assert rel_fro(x, a.float() @ w.float().t() + bias + x0) < 1e-5 and torch.equal(xb, x.to(dt))
    # cast_stats and the consumer GEMM on the same buffers
    raw_xb2, xb2, _ = _guarded((M, N), dt)
    raw_st2, st2, _ = _guarded((M, N // 128, 2), torch.float32)
    lib_xb, lib_st = ops.cast_stats(x, dt)
    from anyref_b200 import _lib
    rc = _lib.load().sam_cast_stats(x.data_ptr(), N, xb2.data_ptr(), N, _lib.fmt_of(dt), st2.data_ptr(), M, N,
                                    _lib.stream_ptr(x.device))
    assert rc == 0
    torch.cuda.synchronize()
    assert _guards_intact(raw_xb2, pad) and _guards_intact(raw_st2, pad)
    assert torch.equal(xb2, lib_xb) and torch.equal(st2, lib_st)
    w2 = (torch.randn(512, N, device=DEV) * 0.05).to(dt)
    raw_o, out, _ = _guarded((M, 512), dt)
    ops.gemm_ln(xb2, w2, torch.randn(512, device=DEV), torch.randn(512, device=DEV), st2, 1e-6, act="gelu", out=out)
    torch.cuda.synchronize()
    assert _guards_intact(raw_o, pad) and bool(torch.isfinite(out.float()).all())


def test_gemm_rejects_bad_arguments(ops):
    a = torch.zeros(64, 60, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(64, 60, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.gemm(a, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a.cpu(), w.cpu())


@pytest.mark.parametrize("a_mode,b_mode,N,K", [(0, 0, 208, 128), (1, 1, 208, 16), (0, 2, 128, 64), (0, 3, 64, 128),
                                               (0, 4, 80, 128), (0, 6, 96, 128)])
def test_umma_descriptor_conventions(ops, a_mode, b_mode, N, K):
    """Pins the shared-memory descriptor / swizzle conventions the attention kernels rely on."""
    torch.manual_seed(0)
    a = torch.randn(128, K, device=DEV).to(torch.bfloat16)
    if b_mode <= 2:
        b = torch.randn(N, K, device=DEV).to(torch.bfloat16)
        ref = a.float() @ b.float().t()
    else:
        b = torch.randn(K, N, device=DEV).to(torch.bfloat16)
        ref = a.float() @ b.float()
    assert (ops.umma_probe(a, b, N, K, a_mode, b_mode) - ref).abs().max().item() < 1e-3


# ------------------------------------------------------------------------------------------------------------ glue
def test_layernorm_rows(ops):
    torch.manual_seed(1)
    x = torch.randn(4096 + 3, 1280, device=DEV) * 2 + 0.3
    g = torch.rand(1280, device=DEV) + 0.5
    b = torch.randn(1280, device=DEV) * 0.1
    ref = F.layer_norm(x, (1280,), g, b, 1e-6)
    assert (ops.layernorm(x, g, b, 1e-6, torch.float32) - ref).abs().max().item() < 5e-6
    assert (ops.layernorm(x, g, b, 1e-6, torch.bfloat16).float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    r = torch.randn_like(x)
    assert (ops.layernorm(x, g, b, 1e-5, torch.float32, residual=r) - F.layer_norm(x + r, (1280,), g, b, 1e-5)).abs().max() < 5e-6
    assert torch.equal(ops.layernorm(x, None, None, 0.0, torch.bfloat16, normalize=False), x.to(torch.bfloat16))


def test_patch_im2col_is_exact(ops):
    img = torch.randn(2, 3, 1024, 1024, device=DEV)
    for dt_in in (torch.float32, torch.bfloat16):
        got = ops.patch_im2col(img.to(dt_in), 16, torch.bfloat16)
        want = F.unfold(img.to(dt_in).float(), 16, stride=16).transpose(1, 2).reshape(-1, 768).to(torch.bfloat16)
        assert torch.equal(got, want)


def test_im2col3x3_is_exact(ops):
    y = torch.randn(2, 64, 64, 256, device=DEV).to(torch.float16)
    got = ops.im2col3x3(y, 2, 64)
    yp = F.pad(y, (0, 0, 1, 1, 1, 1))
    want = torch.stack([yp[:, ky:ky + 64, kx:kx + 64] for ky in range(3) for kx in range(3)], dim=3).reshape(-1, 9 * 256)
    assert torch.equal(got, want)


def test_layernorm2d_to_nchw(ops):
    z = torch.randn(2 * 4096, 256, device=DEV) * 3
    g = torch.rand(256, device=DEV) + 0.5
    b = torch.randn(256, device=DEV) * 0.1
    got = ops.ln_nhwc_to_nchw(z, g, b, 1e-6, 2, 64, torch.float32)
    x_nchw = z.view(2, 64, 64, 256).permute(0, 3, 1, 2)
    want = O.layer_norm_2d(x_nchw, g, b, 1e-6)
    assert (got - want).abs().max().item() < 5e-6


# ------------------------------------------------------------------------------------------------------- attention
def _window_reference(qkv, bias, rel_h, rel_w, B, heads):
    hd = qkv.shape[1] // 3 // heads
    """image_encoder.py:263-288 + :235-257 + :354-392 on projected qkv; padded tokens carry the qkv bias because the
    reference pads the LayerNorm output with zeros BEFORE the qkv Linear (image_encoder.py:179-183)."""
    E = qkv.shape[1] // 3
    xp = bias.float().view(1, 1, 1, -1).expand(B, 70, 70, 3 * E).clone()
    xp[:, :64, :64] = qkv.float().view(B, 64, 64, 3 * E)
    win = xp.view(B, 5, 14, 5, 14, 3 * E).permute(0, 1, 3, 2, 4, 5).reshape(-1, 14, 14, 3 * E)
    b = win.shape[0]
    q, k, v = win.reshape(b, 196, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, b * heads, 196, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    rh, rw = O._rel_table(14, 14, rel_h.float()), O._rel_table(14, 14, rel_w.float())
    rq = q.reshape(b * heads, 14, 14, hd)
    a = torch.einsum("bhwc,hkc->bhwk", rq, rh)
    c = torch.einsum("bhwc,wkc->bhwk", rq, rw)
    attn = (attn.view(b * heads, 14, 14, 14, 14) + a[..., None] + c[:, :, :, None, :]).view(b * heads, 196, 196)
    o = (attn.softmax(-1) @ v).view(b, heads, 14, 14, hd).permute(0, 2, 3, 1, 4).reshape(b, 14, 14, E)
    return O._unpartition(o, 14, (70, 70), (64, 64)).reshape(B * 4096, E)


def _global_reference(qkv, rel_h, rel_w, heads):
    hd = qkv.shape[1] // 3 // heads
    E = qkv.shape[1] // 3
    x = qkv.float().view(1, 4096, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, heads, 4096, hd)
    q, k, v = x.unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    rh, rw = O._rel_table(64, 64, rel_h.float()), O._rel_table(64, 64, rel_w.float())
    rq = q.reshape(heads, 64, 64, hd)
    a = torch.einsum("bhwc,hkc->bhwk", rq, rh)
    c = torch.einsum("bhwc,wkc->bhwk", rq, rw)
    attn = (attn.view(heads, 64, 64, 64, 64) + a[..., None] + c[:, :, :, None, :]).view(heads, 4096, 4096)
    return (attn.softmax(-1) @ v).permute(1, 0, 2).reshape(4096, E)


@pytest.mark.parametrize("dt,tol", [(torch.float16, 1e-3), (torch.bfloat16, 5e-3)])
@pytest.mark.parametrize("B,heads", [(1, 2), (2, 16)])
def test_windowed_attention(ops, dt, tol, B, heads):
    torch.manual_seed(3)
    E = heads * 80
    qkv = torch.randn(B * 4096, 3 * E, device=DEV).to(dt)
    bias = (torch.randn(3 * E, device=DEV) * 0.5).to(dt)
    rel_h = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
    rel_w = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
    got = ops.attn_window(qkv, bias, ops.window_rel_table(rel_h, rel_w, dt), B, heads)
    assert rel_fro(got, _window_reference(qkv, bias, rel_h, rel_w, B, heads)) < tol


@pytest.mark.parametrize("dt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 8e-3)])
def test_windowed_attention_peaked_logits(ops, dt, tol):
    """Logits with a std of ~13 log2 units: keys beyond the first 32 beat the single-pass reference maximum by more than
    the 2^10 headroom for many rows, so the two-pass redo must run; the result is still the exact softmax."""
    torch.manual_seed(6)
    B, heads = 1, 2
    E = heads * 80
    qkv = torch.randn(B * 4096, 3 * E, device=DEV)
    qkv[:, :2 * E] *= 3.0
    qkv = qkv.to(dt)
    bias = (torch.randn(3 * E, device=DEV) * 0.5).to(dt)
    rel_h = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
    rel_w = (torch.randn(27, 80, device=DEV) * 0.2).to(dt)
    got = ops.attn_window(qkv, bias, ops.window_rel_table(rel_h, rel_w, dt), B, heads)
    assert bool(torch.isfinite(got.float()).all())
    assert rel_fro(got, _window_reference(qkv, bias, rel_h, rel_w, B, heads)) < tol


def test_attention_is_bit_reproducible_under_timing_noise(ops):
    """Regression test for a cross-proxy race: the rel-pos gather scratch of the windowed kernel (generic-proxy stores
    into the dead Q tile) was not fenced against the next TMA load into the same bytes, and ~4 % of runs had a few wrong
    rows in one (window, head) item.  300 launches with L2 / timing perturbation must all give the bits of the first."""
    torch.manual_seed(3)
    dt, B, heads = torch.bfloat16, 2, 16
    E = heads * 80
    qkv = torch.randn(B * 4096, 3 * E, device=DEV).to(dt)
    bias = (torch.randn(3 * E, device=DEV) * 0.5).to(dt)
    tab = ops.window_rel_table((torch.randn(27, 80, device=DEV) * 0.2).to(dt), (torch.randn(27, 80, device=DEV) * 0.2).to(dt), dt)
    gh = ops.global_rel_table((torch.randn(127, 80, device=DEV) * 0.2).to(dt), dt)
    gw = ops.global_rel_table((torch.randn(127, 80, device=DEV) * 0.2).to(dt), dt)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    first_w = ops.attn_window(qkv, bias, tab, B, heads).clone()
    first_g = ops.attn_global(qkv, gh, gw, B, heads).clone()
    for it in range(300):
        if it % 3 == 0:
            junk.random_(0, 255)
        assert torch.equal(ops.attn_window(qkv, bias, tab, B, heads), first_w), f"windowed attention differs in run {it}"
        if it % 4 == 0:
            assert torch.equal(ops.attn_global(qkv, gh, gw, B, heads), first_g), f"global attention differs in run {it}"


def test_windowed_attention_token_map_is_exact(ops):
    """q = k = 0 and zero rel-pos make the softmax uniform, so every output is the mean of v over its window:
    checks the partition / padding / un-partition index maps (image_encoder.py:263-318) independently of the maths."""
    heads, E, B = 2, 160, 2
    qkv = torch.zeros(B * 4096, 3 * E, device=DEV)
    ids = torch.arange(B * 4096, device=DEV, dtype=torch.float32)
    qkv[:, 2 * E:] = (ids % 251)[:, None] / 8.0          # exactly representable in fp16
    qkv = qkv.to(torch.float16)
    bias = torch.zeros(3 * E, device=DEV, dtype=torch.float16)
    bias[2 * E:] = 7.0
    tab = torch.zeros(64, 80, device=DEV, dtype=torch.float16)
    got = ops.attn_window(qkv, bias, tab, B, heads)
    want = _window_reference(qkv, bias, tab[:27], tab[32:59], B, heads)
    assert (got.float() - want).abs().max().item() < 2e-2    # fp16 rounding of p = 1/196 only


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("heads", [2, 16])
def test_windowed_attention_window_map_is_bit_exact(ops, dt, heads):
    """INT-exact partition / pad / un-partition map of the CUDA path (image_encoder.py:263-318), independent of any
    arithmetic: every token carries a random +-8 sign code as its key, and the query of token t is the code of ONE other
    real token pi(t) of the same window.  The matching logit is 64 * 80 * scale = 572, every other one is <= 572 - 150, so
    the fp32 softmax is EXACTLY one-hot (exp underflows to 0) and the kernel must return v[pi(t)] bit for bit -- any
    permutation inside the gather of q / k / v or the scatter of the output changes bits.  Padded window slots carry the
    qkv bias (own sign code): they must never be selected by these queries, and their own outputs must not be stored."""
    g = torch.Generator(device="cpu").manual_seed(11 + heads)
    B, hd = 2, 80
    E = heads * hd
    code = (torch.randint(0, 2, (B, 64, 64, heads, hd), generator=g) * 2 - 1).float() * 8.0     # keys
    bias_code = (torch.randint(0, 2, (3, heads, hd), generator=g) * 2 - 1).float() * 8.0
    v = torch.randn(B, 64, 64, heads, hd, generator=g).to(dt).float()
    q = torch.empty_like(code)
    want = torch.empty_like(v)
    worst = 0.0
    for b in range(B):
        for wy in range(5):
            for wx in range(5):
                ys = torch.arange(14 * wy, min(14 * wy + 14, 64))
                xs = torch.arange(14 * wx, min(14 * wx + 14, 64))
                yy, xx = torch.meshgrid(ys, xs, indexing="ij")
                yy, xx = yy.reshape(-1), xx.reshape(-1)
                perm = torch.randperm(yy.numel(), generator=g)
                q[b, yy, xx] = code[b, yy[perm], xx[perm]]
                want[b, yy, xx] = v[b, yy[perm], xx[perm]]
                # the codes of a window (plus the padding code) must be far from each other, else the test is void
                kk = code[b, yy, xx]                                        # [t, heads, hd]
                kk = torch.cat([kk, bias_code[1][None]], dim=0)
                corr = torch.einsum("thd,uhd->htu", kk, kk) / 64.0
                corr = corr - torch.eye(kk.shape[0])[None] * 1e9
                worst = max(worst, corr.max().item())
    assert worst <= 80 - 24, worst            # margin >= 24 * 64 * 80^-0.5 = 171 > 104 (fp32 exp underflow incl. denormals)
    qkv = torch.cat([q.reshape(B * 4096, E), code.reshape(B * 4096, E), v.reshape(B * 4096, E)], dim=1).to(dt).to(DEV)
    bias = bias_code.reshape(3 * E).to(dt).to(DEV)
    rel_h = (torch.randn(27, hd, generator=g) * 0.01).to(dt).to(DEV)
    rel_w = (torch.randn(27, hd, generator=g) * 0.01).to(dt).to(DEV)
    got = ops.attn_window(qkv, bias, ops.window_rel_table(rel_h, rel_w, dt), B, heads)
    assert torch.equal(got.float().cpu(), want.reshape(B * 4096, E))


@pytest.mark.parametrize("dt,tol", [(torch.float16, 1.5e-3), (torch.bfloat16, 5e-3)])
def test_global_attention(ops, dt, tol):
    torch.manual_seed(4)
    heads, E, B = 2, 160, 2
    qkv = torch.randn(B * 4096, 3 * E, device=DEV).to(dt)
    gh = (torch.randn(127, 80, device=DEV) * 0.2).to(dt)
    gw = (torch.randn(127, 80, device=DEV) * 0.2).to(dt)
    got = ops.attn_global(qkv, ops.global_rel_table(gh, dt), ops.global_rel_table(gw, dt), B, heads)
    want = torch.cat([_global_reference(qkv[i * 4096:(i + 1) * 4096], gh, gw, heads) for i in range(B)])
    assert rel_fro(got, want) < tol


@pytest.mark.parametrize("dt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 8e-3)])
def test_global_attention_peaked_logits(ops, dt, tol):
    """Logits with a std of ~13 log2 units: later key rows exceed the reference maximum taken from the first key row by
    far more than the 2^10 headroom, so the lazy-reference path (rescale O in TMEM, redo the block) must run -- the
    result is still the exact softmax of image_encoder.py:250-256."""
    torch.manual_seed(5)
    heads, E, B = 2, 160, 1
    qkv = torch.randn(B * 4096, 3 * E, device=DEV)
    qkv[:, :2 * E] *= 3.0
    qkv = qkv.to(dt)
    gh = (torch.randn(127, 80, device=DEV) * 0.2).to(dt)
    gw = (torch.randn(127, 80, device=DEV) * 0.2).to(dt)
    got = ops.attn_global(qkv, ops.global_rel_table(gh, dt), ops.global_rel_table(gw, dt), B, heads)
    want = _global_reference(qkv, gh, gw, heads)
    assert bool(torch.isfinite(got.float()).all())
    assert rel_fro(got, want) < tol


@pytest.mark.parametrize("dt,tol", [(torch.float16, 1.5e-3), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("peaked", [False, True])
def test_attention_head_dim_64(ops, dt, tol, peaked):
    """ViT-L / ViT-B (build_sam.py:28-45) have head_dim 64: the same kernels without the 16-wide operand tail."""
    torch.manual_seed(9)
    heads, hd, B = 3, 64, 1
    E = heads * hd
    qkv = torch.randn(B * 4096, 3 * E, device=DEV)
    if peaked:
        qkv[:, :2 * E] *= 3.0
    qkv = qkv.to(dt)
    bias = (torch.randn(3 * E, device=DEV) * 0.5).to(dt)
    rel_h = (torch.randn(27, hd, device=DEV) * 0.2).to(dt)
    rel_w = (torch.randn(27, hd, device=DEV) * 0.2).to(dt)
    got = ops.attn_window(qkv, bias, ops.window_rel_table(rel_h, rel_w, dt), B, heads)
    assert rel_fro(got, _window_reference(qkv, bias, rel_h, rel_w, B, heads)) < (tol * 1.5 if peaked else tol)
    gh = (torch.randn(127, hd, device=DEV) * 0.2).to(dt)
    gw = (torch.randn(127, hd, device=DEV) * 0.2).to(dt)
    got = ops.attn_global(qkv, ops.global_rel_table(gh, dt), ops.global_rel_table(gw, dt), B, heads)
    assert bool(torch.isfinite(got.float()).all())
    assert rel_fro(got, _global_reference(qkv, gh, gw, heads)) < (tol * 1.5 if peaked else tol)


def test_rel_pos_index_tables_match_get_rel_pos(ops):
    """INT-exact: row j of the reversed table is rel_pos[126 - j]; get_rel_pos index (image_encoder.py:345-351)."""
    t = torch.arange(127 * 80, device=DEV, dtype=torch.float32).view(127, 80)
    rev = ops.global_rel_table(t, torch.float32)
    idx = O.rel_pos_index(64)
    q, k = 10, 50
    assert torch.equal(rev[63 - q + k], t[idx[q, k]])
    assert int(idx.min()) == 0 and int(idx.max()) == 126
    assert int(O.rel_pos_index(14).max()) == 26
