"""GPU parity tests of the individual kernels, called through the C ABI (vitgan_b200.ops -> libvitgan_b200.so)
and compared with the CPU oracle / plain torch fp32 formulas on the same seeded inputs.
Tolerances (BASELINE.json): 1e-4 relative for fp32, 2e-2 for bf16 (max|a-b| / max|b| per tensor)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import harness, v1 as o1

pytestmark = pytest.mark.gpu

FP32_TOL, BF16_TOL = 1e-4, 2e-2


@pytest.fixture(scope="module")
def vb():
    import vitgan_b200
    return vitgan_b200


def rel(a, b):
    return harness.rel_err(a, b)


def gen(seed=0):
    return torch.Generator("cpu").manual_seed(seed)


def act_ref(act, v, aux, prm):
    if act == 1: return F.gelu(v)
    if act == 2: return torch.tanh(v)
    if act == 3: return torch.sin(prm * v)
    if act == 4: return torch.sigmoid(v)
    if act == 5: return v * (0.5 * (1 + torch.erf(aux / math.sqrt(2))) + aux * torch.exp(-0.5 * aux * aux) / math.sqrt(2 * math.pi))
    if act == 6: return v * (1 - aux * aux)
    if act == 7: return v * prm * torch.cos(prm * aux)
    if act == 8: return v * aux * (1 - aux)
    return v


GEMM_SHAPES = [(128, 128, 64), (256, 384, 128), (1000, 200, 72), (33280 // 8, 128, 256), (65, 10, 128), (7, 8, 16)]


@pytest.mark.parametrize("path", ["simt_f32", "simt_bf16", "tc"])
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_layouts(vb, path, ta, tb, M, N, K):
    L = vb.lib
    g = gen(M + N + K)
    dt = torch.float32 if path == "simt_f32" else torch.bfloat16
    a = torch.randn((K, M) if ta else (M, K), generator=g).to(dt)
    b = torch.randn((N, K) if tb else (K, N), generator=g).to(dt)
    if path == "tc" and (a.shape[1] % 8 or b.shape[1] % 8):
        pytest.skip("tcgen05 path needs 16-byte row pitch (checked separately in test_gemm_tc_rejects)")
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float())
    out = vb.ops.gemm(a.cuda(), b.cuda(), trans_a=bool(ta), trans_b=bool(tb), out_dtype=torch.float32,
                      path=L.GEMM_TCGEN05 if path == "tc" else L.GEMM_SIMT)
    assert rel(out, ref) < 1e-5 * max(1, K / 64)


@pytest.mark.parametrize("path", ["simt", "tc"])
def test_gemm_split_k_accumulate(vb, path):
    L = vb.lib
    g = gen(3)
    dy = torch.randn(33280 // 4, 384, generator=g).bfloat16()
    x = torch.randn(33280 // 4, 128, generator=g).bfloat16()
    ref = dy.float().t() @ x.float()
    out = vb.ops.gemm(dy.cuda(), x.cuda(), trans_a=True, trans_b=False, accumulate=True,
                      path=L.GEMM_TCGEN05 if path == "tc" else L.GEMM_SIMT)
    assert out.dtype == torch.float32 and rel(out, ref) < 2e-5


@pytest.mark.parametrize("rows,N,K", [(33280 // 4, 384, 128), (1000, 256, 128), (70, 128, 256), (4096, 136, 64)])
def test_gemm_wgrad_fused_bias_gradient(vb, rows, N, K):
    """a_rowsum: the weight-gradient GEMM dW += dY^T X also accumulates db += colsum(dY) from an N=16 MMA against ones
    (tcgen05 accumulate mode); both outputs ADD into their buffers; the CUDA-core path rejects it before launching."""
    L = vb.lib
    g = gen(rows + N)
    dy = torch.randn(rows, N, generator=g).bfloat16()
    x = torch.randn(rows, K, generator=g).bfloat16()
    dw0, db0 = torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    dw, db = dw0.cuda(), db0.cuda()
    vb.ops.gemm(dy.cuda(), x.cuda(), trans_a=True, trans_b=False, accumulate=True, out=dw, rowsum_out=db, path=L.GEMM_TCGEN05)
    assert rel(dw, dw0 + dy.float().t() @ x.float()) < 2e-5
    assert rel(db, db0 + dy.float().sum(0)) < 2e-5
    with pytest.raises(L.VitganError, match="a_rowsum"):
        vb.ops.gemm(dy.cuda(), x.cuda(), trans_a=True, trans_b=False, accumulate=True, out=dw, rowsum_out=db, path=L.GEMM_SIMT)


@pytest.mark.parametrize("path", ["simt_f32", "simt_bf16", "tc"])
@pytest.mark.parametrize("act", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_gemm_epilogue(vb, path, act):
    L = vb.lib
    g = gen(act)
    dt = torch.float32 if path == "simt_f32" else torch.bfloat16
    M, N, K = 300, 136, 64
    a = (torch.randn(M, K, generator=g) * 0.5).to(dt)
    w = (torch.randn(N, K, generator=g) * 0.2).to(dt)
    bias = torch.randn(N, generator=g)
    aux = (torch.rand(M, N, generator=g) * 0.9).to(dt)
    res = torch.randn(M, N, generator=g).to(dt)
    pre_ref = a.float() @ w.float().t() + bias
    ref = act_ref(act, pre_ref, aux.float(), 30.0 if act in (3, 7) else 0.0) + res.float()
    out, pre = vb.ops.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), act=act, act_param=30.0 if act in (3, 7) else 0.0,
                           aux=aux.cuda() if act >= 5 else None, residual=res.cuda(), want_pre=True,
                           path=L.GEMM_TCGEN05 if path == "tc" else L.GEMM_SIMT)
    tol = FP32_TOL if dt == torch.float32 else BF16_TOL
    # SIN / MUL_DSIN (omega_0 = 30): the sine is taken on the fp32 accumulator, and aux is the same bf16 tensor on both sides,
    # so the bf16 path meets the plain 2e-2 as well
    assert rel(pre, pre_ref) < tol and rel(out, ref) < tol


def test_gemm_row_remaps(vb):
    """CLS slot (c_row_group) + positional-embedding broadcast (res_row_mod/off) of the patch-embed epilogue."""
    g = gen(5)
    B, n, K, E = 3, 16, 48, 64
    a = torch.randn(B * n, K, generator=g)
    w = torch.randn(E, K, generator=g)
    pos = torch.randn(n + 1, E, generator=g)
    ref = torch.zeros(B, n + 1, E)
    ref[:, 1:] = (a @ w.t()).view(B, n, E) + pos[1:]
    out = torch.zeros(B * (n + 1), E, device="cuda")
    vb.ops.gemm(a.cuda(), w.cuda(), residual=pos.cuda(), res_row_mod=n, res_row_off=1, c_row_group=n, out=out)
    assert rel(out.view(B, n + 1, E), ref) < FP32_TOL
    for path in (vb.lib.GEMM_SIMT, vb.lib.GEMM_TCGEN05):
        out = torch.zeros(B * (n + 1), E, device="cuda", dtype=torch.bfloat16)
        vb.ops.gemm(a.cuda().bfloat16(), w.cuda().bfloat16(), residual=pos.cuda().bfloat16(), res_row_mod=n, res_row_off=1,
                    c_row_group=n, out=out, path=path)
        assert rel(out.view(B, n + 1, E), ref) < BF16_TOL


@pytest.mark.parametrize("B,n,K,E,off", [(5, 64, 48, 128, 0), (3, 256, 192, 768, 0), (4, 64, 432, 432 + 16, 1), (2, 32, 64, 64, 0)])
def test_gemm_row_remaps_staged_3d(vb, B, n, K, E, off):
    """Row groups that are multiples of 32 (the real patch-embedding shapes: 64 / 256 patches per image) take the staged
    epilogue with 3-D TMA stores (slot 0 of every group untouched) and the broadcast positional table as a TMA-loaded
    residual; must equal the direct-epilogue result and the fp32 formula."""
    g = gen(B + n + K)
    a = torch.randn(B * n, K, generator=g).bfloat16()
    w = (torch.randn(E, K, generator=g) * 0.2).bfloat16()
    bias = torch.randn(E, generator=g)
    pos = torch.randn(n + off, E, generator=g).bfloat16()
    ref = torch.full((B, n + 1, E), 7.0)
    ref[:, 1:] = (a.float() @ w.float().t() + bias).view(B, n, E) + pos.float()[off:]
    out = torch.full((B * (n + 1), E), 7.0, device="cuda", dtype=torch.bfloat16)
    vb.ops.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), residual=pos.cuda(), res_row_mod=n, res_row_off=off, c_row_group=n, out=out,
                path=vb.lib.GEMM_TCGEN05)
    assert rel(out.view(B, n + 1, E), ref) < BF16_TOL
    assert torch.equal(out.view(B, n + 1, E)[:, 0].float().cpu(), torch.full((B, E), 7.0))       # CLS slots untouched
    out2 = torch.full((B * (n + 1), E), 7.0, device="cuda", dtype=torch.bfloat16)
    vb.ops.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), residual=pos.cuda(), res_row_mod=n, res_row_off=off, c_row_group=n, out=out2,
                path=vb.lib.GEMM_SIMT)
    assert rel(out, out2) < BF16_TOL


def test_gemm_tc_rejects_and_errors(vb):
    a = torch.randn(16, 10, device="cuda").bfloat16()       # K=10 -> 20-byte pitch: not TMA-able
    w = torch.randn(32, 10, device="cuda").bfloat16()
    with pytest.raises(vb.lib.VitganError, match="unsupported"):
        vb.ops.gemm(a, w, path=vb.lib.GEMM_TCGEN05)
    vb.ops.gemm(a, w)   # AUTO falls to the library's own CUDA-core kernel, never to torch/CPU
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        vb.ops.gemm(a.cpu(), w.cpu())


@pytest.mark.parametrize("dt,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("rows,E", [(33, 32), (130, 128), (65 * 4, 432), (50, 768), (5000, 768), (4097, 768), (4099, 384), (2500, 432), (3001, 512)])
def test_layernorm(vb, dt, tol, rows, E):
    g = gen(rows + E)
    x = (torch.randn(rows, E, generator=g) * 2 + 0.5).to(dt)
    gam, bet = torch.randn(E, generator=g), torch.randn(E, generator=g)
    dy = torch.randn(rows, E, generator=g).to(dt)
    dres = torch.randn(rows, E, generator=g).to(dt)
    xr = x.float().requires_grad_(True); gr = gam.clone().requires_grad_(True); br = bet.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (E,), gr, br, 1e-5)
    yr.backward(dy.float())
    y, mean, rstd = vb.ops.layernorm_fwd(x.cuda(), gam.cuda(), bet.cuda())
    assert rel(y, yr) < tol
    assert rel(mean, x.float().mean(1)) < 1e-5
    cr, cx = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    dx, dg, db = vb.ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dres=dres.cuda(), dres_colsum=cr, dx_colsum=cx)
    assert rel(dx, xr.grad + dres.float()) < tol
    assert rel(dg, gr.grad) < max(tol, 1e-4) and rel(db, br.grad) < max(tol, 1e-4)
    # fused column sums (bias gradients of the neighbouring Linear layers): fp32 sums of the un-rounded values
    assert rel(cr, dres.float().sum(0)) < 1e-4 and rel(cx, (xr.grad + dres.float()).sum(0)) < max(tol, 1e-3)
    # without the two bias-gradient column sums (a different kernel instantiation at wide rows), and dx only (no parameter gradients)
    dx3, dg3, db3 = vb.ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dres=dres.cuda())
    assert rel(dx3, xr.grad + dres.float()) < tol and rel(dg3, gr.grad) < max(tol, 1e-4) and rel(db3, br.grad) < max(tol, 1e-4)
    dx4, _, _ = vb.ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dx_only=True)
    assert rel(dx4, xr.grad) < tol
    if E <= 128:      # deferred column reductions: per-CTA partials + vg_fold_partials give the same five results
        dx2, part = vb.ops.layernorm_bwd_partials(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dres=dres.cuda())
        outs = [torch.zeros(E, device="cuda") for _ in range(4)]
        vb.ops.fold_partials(part, E, *outs)
        assert torch.equal(dx2, dx)
        for got, want in zip(outs, (dg, db, cr, cx)):
            assert rel(got, want) < 1e-5


@pytest.mark.parametrize("dt,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("bcast", [False, True])
def test_sln(vb, dt, tol, bcast):
    g = gen(9)
    B, S, Fd = 3, 10, 384
    h = torch.randn((S, Fd) if bcast else (B, S, Fd), generator=g).to(dt)
    w = torch.randn(B, S, Fd, generator=g).to(dt)
    p = {"layer_norm.weight": torch.randn(Fd, generator=g), "layer_norm.bias": torch.randn(Fd, generator=g),
         "gamma": torch.randn(1, 1, 1, generator=g), "beta": torch.randn(1, 1, 1, generator=g)}
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    hr, wr = h.float().requires_grad_(True), w.float().requires_grad_(True)
    yr = o1.sln(p, "", hr, wr)
    dy = torch.randn(B, S, Fd, generator=g).to(dt)
    yr.backward(dy.float())
    c = lambda t: t.detach().cuda()
    y, mean, rstd = vb.ops.sln_fwd(c(h).reshape(-1, Fd), c(w).reshape(-1, Fd), c(p["layer_norm.weight"]), c(p["layer_norm.bias"]),
                                   c(p["gamma"]).reshape(1), c(p["beta"]).reshape(1))
    assert rel(y, yr.reshape(-1, Fd)) < tol
    dh, dw, dgs, dbs, dlg, dlb = vb.ops.sln_bwd(c(dy).reshape(-1, Fd), c(h).reshape(-1, Fd), c(w).reshape(-1, Fd), mean, rstd,
                                                c(p["layer_norm.weight"]), c(p["layer_norm.bias"]), c(p["gamma"]).reshape(1),
                                                c(p["beta"]).reshape(1))
    assert rel(dh, hr.grad.reshape(-1, Fd)) < tol and rel(dw, wr.grad.reshape(-1, Fd)) < tol
    t2 = max(tol, 2e-4)
    assert rel(dgs, p["gamma"].grad.reshape(1)) < t2 and rel(dbs, p["beta"].grad.reshape(1)) < t2
    assert rel(dlg, p["layer_norm.weight"].grad) < t2 and rel(dlb, p["layer_norm.bias"].grad) < t2


def attn_ref(q, k, v, scale, mode):
    # q,k,v: (B,H,S,d) fp32
    if mode == 1:
        s = torch.cdist(q, k, p=2)       # S > 25 -> matmul path, same as the reference
    else:
        s = q @ k.transpose(-1, -2)
    return torch.softmax(s * scale, -1) @ v


@pytest.mark.parametrize("dt,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,H,S,d", [(3, 4, 65, 32), (2, 4, 65, 108), (2, 4, 64, 96), (1, 2, 257, 192), (2, 2, 30, 12)])
def test_attention(vb, dt, tol, mode, B, H, S, d):
    g = gen(S + d + mode)
    hd = H * d
    qkv = (torch.randn(B * S, 3 * hd, generator=g) * (1.0 if mode else 0.7)).to(dt)
    d_o = torch.randn(B * S, hd, generator=g).to(dt)
    scale = 1.0 / math.sqrt(d if mode == 0 else hd)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    oref = attn_ref(q, k, v, scale, mode).permute(0, 2, 1, 3).reshape(B * S, hd)
    oref.backward(d_o.float())
    qc = qkv.cuda()
    o, lse = vb.ops.attention_fwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], B, H, S, d, scale, mode)
    assert rel(o, oref) < tol
    # backward consumes the GPU forward's own (rounded) o / lse, like the real pipeline
    dqkv = vb.ops.attention_bwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], o, d_o.cuda(), lse, B, H, S, d, scale, mode)
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * hd:(i + 1) * hd], ref_in.grad[:, i * hd:(i + 1) * hd]) < tol, name


def test_attention_l2_zero_distance_is_finite(vb):
    """q == k rows give dist = 0 on the diagonal: gradients must stay finite (guarded 1/dist), SURVEY Q6."""
    B, H, S, d = 1, 1, 40, 16
    x = torch.randn(B * S, d, generator=gen(1)).cuda()
    qkv = torch.cat([x, x, x], 1).contiguous()
    o, lse = vb.ops.attention_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, H, S, d, 0.25, 1)
    dqkv = vb.ops.attention_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], o, torch.ones_like(o), lse, B, H, S, d, 0.25, 1)
    assert torch.isfinite(o).all() and torch.isfinite(dqkv).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_patch_gathers(vb, dt):
    g = gen(2)
    B, C, I, P = 3, 3, 32, 4
    img = torch.randn(B, C, I, I, generator=g)
    ref = F.unfold(img, kernel_size=P, stride=P).transpose(1, 2).reshape(B * (I // P) ** 2, C * P * P)
    out = vb.ops.im2col(img.cuda(), P, dt)
    assert torch.equal(out.float().cpu(), ref.to(dt).float())            # pure data movement: bit exact
    back = vb.ops.col2im(out, B, C, I, P)
    assert torch.equal(back.cpu(), img.to(dt).float())
    # v1 scrambled tokens against the oracle restatement of patch_encoder.py:54-73, 32 and 64 px
    for I in (32, 64):
        cfg = o1.V1Config(image_size=I)
        n = int(round(math.sqrt(cfg.number_of_tokens)))
        img = torch.randn(2, 3, I, I, generator=g)
        tok = vb.ops.v1_tokens_fwd(img.cuda(), cfg.window, cfg.stride, n, dt)
        ref = o1.get_tokens(img, cfg).reshape(-1, cfg.token_size)
        assert torch.equal(tok.float().cpu(), ref.to(dt).float())
        # adjoint check: <gather(x), y> == <x, scatter(y)>
        y = torch.randn(ref.shape, generator=g)
        imgr = img.clone().requires_grad_(True)
        (o1.get_tokens(imgr, cfg).reshape(-1, cfg.token_size) * y).sum().backward()
        dimg = vb.ops.v1_tokens_bwd(y.to(dt).cuda(), 2, 3, I, cfg.window, cfg.stride, n)
        assert rel(dimg, imgr.grad) < (1e-5 if dt == torch.float32 else BF16_TOL)


def test_embed_split_colsum_fill(vb):
    g = gen(4)
    B, S, E = 5, 17, 64
    dx = torch.randn(B, S, E, generator=g)
    for has_cls in (False, True):
        dtok, dcls, dpos = vb.ops.embed_bwd_split(dx.cuda(), has_cls)
        assert torch.equal(dtok.cpu().view(B, S - 1, E), dx[:, 1:])
        assert rel(dcls, dx[:, 0].sum(0)) < 1e-6
        assert rel(dpos, dx.sum(0) if has_cls else dx[:, 1:].sum(0)) < 1e-6
    for Bb, Ss, Ee, dt_ in ((512, 65, 128, torch.bfloat16), (37, 9, 128, torch.bfloat16), (70, 5, 768, torch.float32), (3, 4, 36, torch.float32)):
        dxl = torch.randn(Bb, Ss, Ee, generator=g).to(dt_)        # vectorised kernel (16 B column groups) and its tails
        dtok, dcls, dpos = vb.ops.embed_bwd_split(dxl.cuda(), False)
        assert torch.equal(dtok.cpu().view(Bb, Ss - 1, Ee), dxl[:, 1:])
        assert rel(dcls, dxl[:, 0].float().sum(0)) < 1e-5 and rel(dpos, dxl[:, 1:].float().sum(0)) < 1e-5
    x = torch.randn(1000, 200, generator=g)
    assert rel(vb.ops.colsum(x.cuda()), x.sum(0)) < 1e-5
    assert rel(vb.ops.colsum(x.bfloat16().cuda()), x.bfloat16().float().sum(0)) < 1e-5
    t = torch.zeros(B, S, E, device="cuda")
    v, v2 = torch.randn(E, generator=g), torch.randn(E, generator=g)
    vb.ops.fill_rows(t, 0, v.cuda(), v2.cuda())
    assert torch.equal(t[:, 0].cpu(), (v + v2).expand(B, E)) and t[:, 1:].abs().sum() == 0


def test_sigma_max_power_iteration(vb):
    """vs the reference's full SVD (attention.py:54-58) on v1-discriminator-shaped weights."""
    g = gen(6)
    mats = [(torch.rand(108, 432, generator=g) * 2 - 1) / math.sqrt(432) for _ in range(12)]
    dev = [m.cuda() for m in mats]
    ptrs = torch.tensor([m.data_ptr() for m in dev], dtype=torch.int64).cuda()
    u = torch.zeros(12, 108, device="cuda")
    s_cold = vb.ops.sigma_max(ptrs, 12, 108, 432, u, 400)
    ref = torch.stack([torch.linalg.svdvals(m).max() for m in mats])
    assert rel(s_cold, ref) < 1e-4
    s_warm = vb.ops.sigma_max(ptrs, 12, 108, 432, u, 4)        # persistent vector: already converged
    assert rel(s_warm, s_cold) < 1e-6
    # a matrix too large for shared memory takes the global-memory variant
    big = torch.randn(300, 400, generator=g)
    bp = torch.tensor([big.cuda().data_ptr()], dtype=torch.int64).cuda()
    bigc = big.cuda()
    bp = torch.tensor([bigc.data_ptr()], dtype=torch.int64).cuda()
    sb = vb.ops.sigma_max(bp, 1, 300, 400, torch.zeros(1, 300, device="cuda"), 600)
    assert rel(sb, torch.linalg.svdvals(big).max().reshape(1)) < 1e-3


@pytest.mark.parametrize("decoupled,wd,betas", [(True, 1e-3, (0.9, 0.999)), (False, 0.0, (0.5, 0.999))])
def test_fused_adam_matches_torch(vb, decoupled, wd, betas):
    g = gen(8)
    p0 = torch.randn(10007, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([pr], lr=5e-4, betas=betas, weight_decay=wd)
    p = p0.cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(5):
        gr = torch.randn(10007, generator=g)
        pr.grad = gr.clone()
        opt.step()
        vb.ops.adam_step(p, gr.cuda(), m, v, cnt, 5e-4, betas[0], betas[1], 1e-8, wd, decoupled)
    assert int(cnt.item()) == 5 and rel(p, pr) < 1e-6


def test_cast_scale_and_act_backward(vb):
    g = gen(10)
    x = torch.randn(1001, generator=g)
    num, den = torch.tensor([3.0]).cuda(), torch.tensor([4.0]).cuda()
    assert rel(vb.ops.cast(x.cuda(), torch.bfloat16, num=num, den=den), (x * 0.75).bfloat16().float()) < 4e-3
    assert torch.equal(vb.ops.cast(x.cuda(), torch.bfloat16).cpu(), x.bfloat16())
    dy, aux = torch.randn(64, 33, generator=g), torch.rand(64, 33, generator=g)
    for act, prm in ((1, 0.0), (2, 0.0), (3, 30.0), (4, 0.0)):
        out = vb.ops.act_backward(dy.cuda(), aux.cuda(), act, prm)
        assert rel(out, act_ref(act + 4, dy, aux, prm)) < 1e-5


@pytest.mark.parametrize("B,H,S,d", [(3, 4, 65, 32), (300, 4, 65, 32), (2, 8, 65, 32), (4, 2, 64, 64), (3, 4, 128, 64), (5, 4, 17, 32), (2, 4, 1, 32),
                                     (7, 2, 96, 32), (3, 1, 80, 64), (2, 6, 33, 32), (3, 4, 100, 32), (600, 2, 48, 32)])
def test_attention_tensor_core_path(vb, B, H, S, d):
    """bf16 dot-product attention with S <= 128 and d in {32, 64} runs on tcgen05 (attention_tc.cu): forward and the
    five-GEMM backward against the fp32 formula, and against this library's CUDA-core flash kernel (VG_ATTN_PATH=simt
    is read once per process, so the cross-check uses the fp32 path which always takes the CUDA-core kernel)."""
    g = gen(B + S + d)
    hd = H * d
    qkv = (torch.randn(B * S, 3 * hd, generator=g) * 0.8).bfloat16()
    d_o = torch.randn(B * S, hd, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(d)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(-1, -2)) * scale
    oref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, hd)
    lse_ref = torch.logsumexp(s, -1).reshape(-1)
    oref.backward(d_o.float())
    qc = qkv.cuda()
    o, lse = vb.ops.attention_fwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], B, H, S, d, scale, 0)
    assert rel(o, oref) < BF16_TOL
    assert rel(lse, lse_ref) < 1e-3
    dqkv = vb.ops.attention_bwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], o, d_o.cuda(), lse, B, H, S, d, scale, 0)
    assert torch.isfinite(dqkv).all()
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * hd:(i + 1) * hd], ref_in.grad[:, i * hd:(i + 1) * hd]) < BF16_TOL, name
    # same inputs through the CUDA-core kernel in fp32
    qf = qkv.float().cuda()
    o32, lse32 = vb.ops.attention_fwd(qf[:, :hd], qf[:, hd:2 * hd], qf[:, 2 * hd:], B, H, S, d, scale, 0)
    assert rel(o, o32) < BF16_TOL and rel(lse, lse32) < 1e-3


def test_softmax_ce_fused_head(vb):
    """vg_softmax_ce == nn.CrossEntropyLoss per group (values and gradient of the summed losses), fp32 1e-5."""
    g = gen(77)
    for rows, C, rpg in ((1024, 10, 512), (512, 10, 512), (96, 7, 32)):
        z = (torch.randn(rows, C, generator=g) * 3).requires_grad_(True)
        t = torch.randint(0, C, (rows,), generator=g)
        ref = torch.stack([F.cross_entropy(z[i:i + rpg], t[i:i + rpg]) for i in range(0, rows, rpg)])
        ref.sum().backward()
        zc = z.detach().cuda().requires_grad_(True)
        got = vb.functional.softmax_ce(zc, t.cuda(), rpg)
        got.sum().backward()
        assert rel(got, ref.detach()) < 1e-5 and rel(zc.grad, z.grad) < 1e-5


def test_bce_fused_head(vb):
    """vg_bce == nn.BCELoss per group (src/v1/gan.py:16-20): values and gradient of the summed losses at fp32 1e-5, including
    saturated probabilities (torch's -100 log clamp and 1e-12 denominator clamp)."""
    g = gen(78)
    for rows, rpg in ((256, 128), (128, 128), (96, 32)):
        p = torch.sigmoid(torch.randn(rows, 1, generator=g) * 4)
        p[0, 0], p[1, 0] = 0.0, 1.0
        p.requires_grad_(True)
        t = (torch.rand(rows, 1, generator=g) > 0.5).float()
        t[0, 0], t[1, 0] = 0.0, 1.0                  # saturated and correct: loss 0, gradient 0
        ref = torch.stack([F.binary_cross_entropy(p[i:i + rpg], t[i:i + rpg]) for i in range(0, rows, rpg)])
        ref.sum().backward()
        pc = p.detach().cuda().requires_grad_(True)
        got = vb.functional.bce(pc, t.cuda(), rpg)
        got.sum().backward()
        assert rel(got, ref.detach()) < 1e-5 and rel(pc.grad, p.grad) < 1e-5
    p = torch.tensor([[1.0], [0.0], [0.3]])          # saturated and wrong: -100 clamp, 1e12 gradient
    t = torch.tensor([[0.0], [1.0], [1.0]])
    pr = p.clone().requires_grad_(True)
    ref = F.binary_cross_entropy(pr, t)
    ref.backward()
    pc = p.cuda().requires_grad_(True)
    got = vb.functional.bce(pc, t.cuda())
    got.sum().backward()
    assert rel(got, ref.detach().reshape(1)) < 1e-6 and rel(pc.grad, pr.grad) < 1e-6


def test_denorm_u8_is_byte_exact(vb):
    """vg_denorm_u8 == utils.convert_to_uint8 of the reference (src/v2/utils.py:194-196), bit for bit."""
    from oracle import v2 as o2
    g = gen(5)
    x = torch.cat([torch.randn(3 * 32 * 32 * 7 + 3, generator=g) * 1.2, torch.tensor([-1.0, 1.0, 0.0, -2.0, 2.0, 0.999999, -0.999999])])
    assert torch.equal(vb.ops.denorm_u8(x.cuda()).cpu(), o2.convert_to_uint8(x))
    xb = x.bfloat16()
    assert torch.equal(vb.ops.denorm_u8(xb.cuda()).cpu(), o2.convert_to_uint8(xb.float()))
    # more elements than one pass of the capped launch grid covers (148 SMs x 16 CTAs x 256 threads x 4 = 2.4 M): the c5 sampling
    # batch (4096 x 3 x 32 x 32 = 12.6 M) must be written completely, odd tail included
    big = torch.randn(4096 * 3 * 32 * 32 + 3, generator=g)
    got = vb.ops.denorm_u8(big.cuda()).cpu()
    assert torch.equal(got, o2.convert_to_uint8(big))
    assert torch.equal(vb.ops.denorm_u8(big.bfloat16().cuda()).cpu(), o2.convert_to_uint8(big.bfloat16().float()))


@pytest.mark.parametrize("M,K,res", [(33280 // 8, 128, True), (300, 256, True), (65, 128, False)])
def test_gemm_fused_layernorm_epilogue(vb, M, K, res):
    """vg_gemm ln_*: C = A W^T + b (+ residual) and LayerNorm(C) with its statistics from ONE launch (N == 128); the
    normalised output must equal a separate vg_layernorm_fwd on the stored bf16 C to bf16 rounding, and the fp32 formula."""
    L = vb.lib
    g = gen(M + K)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(128, K, generator=g) * 0.1).bfloat16()
    b = torch.randn(128, generator=g)
    r = torch.randn(M, 128, generator=g).bfloat16() if res else None
    gam, bet = torch.randn(128, generator=g), torch.randn(128, generator=g)
    c, xn, mean, rstd = vb.ops.gemm(a.cuda(), w.cuda(), bias=b.cuda(), residual=None if r is None else r.cuda(),
                                    ln=(gam.cuda(), bet.cuda(), 1e-5), path=L.GEMM_TCGEN05)
    c_ref = a.float() @ w.float().t() + b + (0 if r is None else r.float())
    assert rel(c, c_ref) < BF16_TOL
    xn2, mean2, rstd2 = vb.ops.layernorm_fwd(c, gam.cuda(), bet.cuda())
    assert rel(mean, mean2) < 1e-5 and rel(rstd, rstd2) < 1e-5
    assert rel(xn, xn2) < 1e-2          # both are bf16 roundings of the same fp32 value up to summation order
    assert rel(xn, F.layer_norm(c.float().cpu(), (128,), gam, bet, 1e-5)) < BF16_TOL
    with pytest.raises(L.VitganError, match="LayerNorm"):
        vb.ops.gemm(a.cuda(), torch.randn(256, K).bfloat16().cuda(), ln=(torch.ones(256).cuda(), torch.zeros(256).cuda(), 1e-5))


def test_gemm_gelu_epilogue_large_arguments(vb):
    """The bf16 GELU / GELU' epilogues evaluate erf through a fitted tanh polynomial whose argument is clamped to [-8, 8]:
    pre-activations far outside that range (outliers) must still give GELU(x) = x / 0 and GELU'(x) = 1 / 0."""
    L = vb.lib
    bias = torch.cat([torch.linspace(-60, 60, 120), torch.tensor([-1e4, 1e4, -11.7, 11.7, -9.0, 9.0, -3.0, 3.0])])
    a = torch.zeros(256, 64).bfloat16()
    w = torch.zeros(128, 64).bfloat16()
    out, pre = vb.ops.gemm(a.cuda(), w.cuda(), bias=bias.cuda(), act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
    ref = F.gelu(bias.bfloat16().float()).expand(256, 128)
    assert torch.isfinite(out).all() and rel(out, ref) < BF16_TOL
    assert (out.float().cpu()[0] - ref[0]).abs().max() <= 8e-3 * ref.abs().max()
    dy = torch.ones(256, 128).bfloat16()
    eye = torch.eye(128).bfloat16()                              # dX = dY I  x  GELU'(aux)
    dgrad = vb.ops.gemm(dy.cuda(), eye.cuda(), trans_b=False, act=L.ACT_MUL_DGELU, aux=pre, path=L.GEMM_TCGEN05)
    xr = bias.bfloat16().float().requires_grad_(True)
    F.gelu(xr).sum().backward()
    assert torch.isfinite(dgrad).all() and (dgrad.float().cpu()[0] - xr.grad).abs().max() < 1e-2


# ---------------------------------------------------------------------------------------------------------------
# Production shapes of BASELINE configs[1] (C2): M = 512 x 65 = 33 280 rows (66 560 for the merged discriminator pass), i.e. 5-11
# tiles per CTA through the persistent loop of the 128-wide tcgen05 GEMM: accumulator-stage phase wrap, the alternating epilogue
# staging tiles (`wait_group.read 1`), and the fused epilogues at more than one tile per CTA.  Checker: the fp32 formula on the
# same bf16-rounded operands, evaluated by torch on the device.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [33280, 66560])
def test_gemm_production_rows_all_epilogues(vb, M):
    L = vb.lib
    g = gen(M)
    E, m = 128, 2
    dev = "cuda"
    x = (torch.randn(M, E, generator=g) * 0.7).bfloat16().to(dev)
    res = torch.randn(M, E, generator=g).bfloat16().to(dev)
    wqkv = (torch.randn(3 * E, E, generator=g) * 0.1).bfloat16().to(dev)
    w1 = (torch.randn(m * E, E, generator=g) * 0.1).bfloat16().to(dev)
    w2 = (torch.randn(E, m * E, generator=g) * 0.1).bfloat16().to(dev)
    bq, b1, b2 = (torch.randn(n, generator=g).to(dev) for n in (3 * E, m * E, E))
    gam, bet = (1 + 0.1 * torch.randn(E, generator=g)).to(dev), (0.1 * torch.randn(E, generator=g)).to(dev)
    xf = x.float()
    # forward, bias
    qkv = vb.ops.gemm(x, wqkv, bias=bq, path=L.GEMM_TCGEN05)
    assert rel(qkv, xf @ wqkv.float().t() + bq) < BF16_TOL
    # forward, GELU + pre-activation
    h, u = vb.ops.gemm(x, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
    u_ref = xf @ w1.float().t() + b1
    assert rel(u, u_ref) < BF16_TOL and rel(h, F.gelu(u_ref)) < BF16_TOL
    # forward, residual + fused LayerNorm epilogue (N = 128: one tile is one complete row)
    y, xn, mean, rstd = vb.ops.gemm(h, w2, bias=b2, residual=res, ln=(gam, bet, 1e-5), path=L.GEMM_TCGEN05)
    y_ref = h.float() @ w2.float().t() + b2 + res.float()
    assert rel(y, y_ref) < BF16_TOL
    yb = y.float()                                                     # the kernel normalises the bf16-rounded row
    assert rel(mean, yb.mean(1)) < 1e-3 and rel(rstd, (yb.var(1, unbiased=False) + 1e-5).rsqrt()) < 1e-3
    assert rel(xn, F.layer_norm(yb, (E,), gam, bet, 1e-5)) < BF16_TOL
    # dgrad with GELU' epilogue:  dU = (dH W2) * gelu'(u)
    dy = torch.randn(M, E, generator=g).bfloat16().to(dev)
    du = vb.ops.gemm(dy, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=u, path=L.GEMM_TCGEN05)
    du_ref = act_ref(5, dy.float() @ w2.float(), u.float(), 0.0)
    assert rel(du, du_ref) < BF16_TOL
    # wgrad + fused bias gradient (a_rowsum), accumulating
    dw0, db0 = torch.randn(3 * E, E, generator=g).to(dev), torch.randn(3 * E, generator=g).to(dev)
    dw, db = dw0.clone(), db0.clone()
    vb.ops.gemm(qkv, x, trans_a=True, trans_b=False, accumulate=True, out=dw, rowsum_out=db, path=L.GEMM_TCGEN05)
    assert rel(dw, dw0 + qkv.float().t() @ xf) < 1e-4
    assert rel(db, db0 + qkv.float().sum(0)) < 1e-4


@pytest.mark.parametrize("M", [257 * 256, 130 * 128 + 37])
def test_gemm_cta_pair_wide_tiles(vb, M):
    """Compute-bound shapes (E = 768, the scaled config of BASELINE.json configs[3]) run on the CTA-pair kernel (256 x 256 tile per
    2-CTA cluster, tcgen05.mma.cta_group::2): every Linear product of the encoder block at production rows (65 792 = 257 pair
    tiles) and at a row count that leaves the last pair half empty plus a ragged 37-row tail -- forward + bias, forward + GELU +
    pre-activation, forward + bias + residual (side tile prefetched one group ahead), dgrad (B MN-major), dgrad * GELU'(u)
    (aux chain), split-K wgrad (A and B MN-major, TMA reduce-add) -- against the fp32 formula on the same bf16 operands."""
    L = vb.lib
    g = gen(M % 1000)
    E, m, dev = 768, 2, "cuda"
    x = (torch.randn(M, E, generator=g) * 0.7).bfloat16().to(dev)
    wqkv = (torch.randn(3 * E, E, generator=g) * 0.05).bfloat16().to(dev)
    w1 = (torch.randn(m * E, E, generator=g) * 0.05).bfloat16().to(dev)
    w2 = (torch.randn(E, m * E, generator=g) * 0.05).bfloat16().to(dev)
    bq, b1, b2 = (torch.randn(n, generator=g).to(dev) for n in (3 * E, m * E, E))
    xf = x.float()
    qkv = vb.ops.gemm(x, wqkv, bias=bq, path=L.GEMM_TCGEN05)
    assert rel(qkv, xf @ wqkv.float().t() + bq) < BF16_TOL
    h, u = vb.ops.gemm(x, w1, bias=b1, act=L.ACT_GELU, want_pre=True, path=L.GEMM_TCGEN05)
    u_ref = xf @ w1.float().t() + b1
    assert rel(u, u_ref) < BF16_TOL and rel(h, F.gelu(u_ref)) < BF16_TOL
    y = vb.ops.gemm(h, w2, bias=b2, residual=x, path=L.GEMM_TCGEN05)
    assert rel(y, h.float() @ w2.float().t() + b2 + xf) < BF16_TOL
    dqkv = (torch.randn(M, 3 * E, generator=g) * 0.5).bfloat16().to(dev)
    dx = vb.ops.gemm(dqkv, wqkv, trans_b=False, path=L.GEMM_TCGEN05)
    assert rel(dx, dqkv.float() @ wqkv.float()) < BF16_TOL
    dy = torch.randn(M, E, generator=g).bfloat16().to(dev)
    du = vb.ops.gemm(dy, w2, trans_b=False, act=L.ACT_MUL_DGELU, aux=u, path=L.GEMM_TCGEN05)
    assert rel(du, act_ref(5, dy.float() @ w2.float(), u.float(), 0.0)) < BF16_TOL
    dw0 = torch.randn(3 * E, E, generator=g).to(dev)
    dw = dw0.clone()
    vb.ops.gemm(dqkv, x, trans_a=True, trans_b=False, accumulate=True, out=dw, path=L.GEMM_TCGEN05)
    assert rel(dw, dw0 + dqkv.float().t() @ xf) < 1e-4


@pytest.mark.parametrize("M,N,K", [(9605, 512, 520), (19000, 256, 1000), (2304, 768, 8200)])
def test_gemm_cta_pair_ragged_shapes(vb, M, N, K):
    """CTA-pair kernel at shapes that exercise its edges: a ragged last pair tile (M % 256 != 0, incl. a pair whose second CTA has
    no rows), K % 64 != 0 (the last k-block is zero-filled by TMA), a single n-tile (N = 256), and -- third case, the wgrad
    geometry -- split-K with a ragged last split reduce-added into a pre-filled fp32 output."""
    L = vb.lib
    g = gen(M + N + K)
    dev = "cuda"
    if K < 4096:
        x = (torch.randn(M, K, generator=g) * 0.5).bfloat16().to(dev)
        w = (torch.randn(N, K, generator=g) * 0.05).bfloat16().to(dev)
        b = torch.randn(N, generator=g).to(dev)
        y = vb.ops.gemm(x, w, bias=b, path=L.GEMM_TCGEN05)
        assert rel(y, x.float() @ w.float().t() + b) < BF16_TOL
        dy = (torch.randn(M, N, generator=g) * 0.5).bfloat16().to(dev)
        dx = vb.ops.gemm(dy, w, trans_b=False, path=L.GEMM_TCGEN05) if K % 256 == 0 else None     # dgrad output width K must suit the wide tile
        if dx is not None:
            assert rel(dx, dy.float() @ w.float()) < BF16_TOL
    else:
        dy = (torch.randn(K, M, generator=g) * 0.5).bfloat16().to(dev)       # [rows, out_features]
        x = (torch.randn(K, N, generator=g) * 0.5).bfloat16().to(dev)        # [rows, in_features]
        dw0 = torch.randn(M, N, generator=g).to(dev)
        dw = dw0.clone()
        vb.ops.gemm(dy, x, trans_a=True, trans_b=False, accumulate=True, out=dw, path=L.GEMM_TCGEN05)
        assert rel(dw, dw0 + dy.float().t() @ x.float()) < 1e-4


MT_SHAPES = [(2, 4, 257, 192, 0), (40, 4, 257, 192, 0), (1, 1, 128, 192, 0), (1, 2, 129, 192, 0), (3, 4, 64, 96, 0), (2, 4, 65, 112, 0),
             (2, 4, 65, 112, 1), (300, 4, 65, 112, 1), (2, 4, 64, 96, 1), (1, 2, 200, 112, 1), (2, 2, 272, 96, 0), (2, 2, 17, 96, 0),
             (3, 2, 1, 112, 1)]


@pytest.mark.parametrize("B,H,S,d,mode", MT_SHAPES)
def test_attention_multi_tile_tensor_core_path(vb, B, H, S, d, mode):
    """bf16 attention with head widths 96 / 112 / 192 runs on the multi-tile tcgen05 kernels (attention_mt.cu) -- the scaled v2
    config (S = 257, d = 192; src/v2/modules.py:142-159) and both v1 head shapes incl. the L2-distance scores of the
    discriminator (src/v1/attention.py:43-52,66-70; torch.cdist matmul-path semantics): forward (o, lse) and the backward
    (dq, dk, dv) against the fp32 formula at 2e-2, and against this library's CUDA-core flash kernel in fp32."""
    assert vb.lib.lib.vg_attention_path(1, mode, B, H, S, d) == 2
    g = gen(B + S + d + mode)
    hd = H * d
    qkv = (torch.randn(B * S, 3 * hd, generator=g) * (1.0 if mode else 0.7)).bfloat16()
    d_o = torch.randn(B * S, hd, generator=g).bfloat16()
    scale = 1.0 / math.sqrt(d if mode == 0 else hd)
    ref_in = qkv.float().cuda().requires_grad_(True)                    # fp32 formula, evaluated by torch on the device
    q, k, v = [ref_in[:, i * hd:(i + 1) * hd].reshape(B, S, H, d).permute(0, 2, 1, 3) for i in range(3)]
    if mode == 1:
        s = ((q * q).sum(-1, keepdim=True) + (k * k).sum(-1, keepdim=True).transpose(-1, -2) - 2 * q @ k.transpose(-1, -2)).clamp_min(0).sqrt() * scale
    else:
        s = (q @ k.transpose(-1, -2)) * scale
    oref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, hd)
    lse_ref = torch.logsumexp(s, -1).reshape(-1)
    oref.backward(d_o.float().cuda())
    qc = qkv.cuda()
    o, lse = vb.ops.attention_fwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], B, H, S, d, scale, mode)
    assert rel(o, oref) < BF16_TOL and rel(lse, lse_ref) < 1e-3
    dqkv = vb.ops.attention_bwd(qc[:, :hd], qc[:, hd:2 * hd], qc[:, 2 * hd:], o, d_o.cuda(), lse, B, H, S, d, scale, mode)
    assert torch.isfinite(dqkv).all()
    for i, name in enumerate("qkv"):
        assert rel(dqkv[:, i * hd:(i + 1) * hd], ref_in.grad[:, i * hd:(i + 1) * hd]) < BF16_TOL, name
    if B * H * S * S < 5e7:                                             # same inputs through the CUDA-core kernel in fp32
        qf = qkv.float().cuda()
        o32, lse32 = vb.ops.attention_fwd(qf[:, :hd], qf[:, hd:2 * hd], qf[:, 2 * hd:], B, H, S, d, scale, mode)
        assert rel(o, o32) < BF16_TOL and rel(lse, lse32) < 1e-3


def test_attention_multi_tile_l2_zero_distance_is_finite(vb):
    """q == k rows give dist = 0 on the diagonal: the guarded 1/dist of the tcgen05 L2 backward keeps the gradients finite (Q6)."""
    B, H, S, d = 2, 2, 65, 112
    x = torch.randn(B * S, H * d, generator=gen(1)).bfloat16().cuda()
    qkv = torch.cat([x, x, x], 1).contiguous()
    hd = H * d
    assert vb.lib.lib.vg_attention_path(1, 1, B, H, S, d) == 2
    o, lse = vb.ops.attention_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, H, S, d, 0.1, 1)
    dqkv = vb.ops.attention_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], o, torch.ones_like(o), lse, B, H, S, d, 0.1, 1)
    assert torch.isfinite(o).all() and torch.isfinite(dqkv).all()


def test_pack_pad_grouped_heads(vb):
    """vg_pack_pad: 3H per-head [108, 432] weights (pointer array) -> one bf16 [3H x 112, 432] operand, rescaled per head by
    sigma_init / sigma_now and zero padded, in one launch; with n = 1 it pads the columns of the out-proj weight."""
    g = gen(9)
    ws = [torch.randn(108, 432, generator=g).cuda() for _ in range(12)]
    num, den = (torch.rand(12, generator=g) + 0.5).cuda(), (torch.rand(12, generator=g) + 0.5).cuda()
    ptrs = torch.tensor([w.data_ptr() for w in ws], dtype=torch.int64).cuda()
    out = vb.ops.pack_pad(ptrs, 12, 108, 432, 112, 432, torch.bfloat16, num, den).view(12, 112, 432)
    for i, w in enumerate(ws):
        assert torch.equal(out[i, :108], (w * num[i] / den[i]).bfloat16()) and not out[i, 108:].any()
    wo = torch.randn(432, 432, generator=g).cuda()
    ptr = torch.tensor([wo.data_ptr()], dtype=torch.int64).cuda()
    pad = vb.ops.pack_pad(ptr, 1, 432 * 4, 108, 432 * 4, 112, torch.float32).view(432, 4, 112)
    assert torch.equal(pad[:, :, :108], wo.view(432, 4, 108)) and not pad[:, :, 108:].any()
