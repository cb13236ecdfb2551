"""Per-kernel numerics checks: each CUDA kernel behind the C ABI against a plain PyTorch fp32 evaluation of the same
op on the same (bf16-rounded) inputs.  Every check returns a dict of error statistics and raises AssertionError with
a diagnostic message when out of tolerance.  Used by tests/test_kernels_gpu.py and tools/gpu_probe.py."""
import math

import torch

from simple_tad_b200 import _lib as L

DEV = "cuda"


def _stats(got, ref, name, atol, rtol):
    got = got.float()
    ref = ref.float()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    finite = bool(torch.isfinite(got).all())
    diff = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = diff > tol
    nbad = int(bad.sum())
    rel_l2 = float((got - ref).norm() / (ref.norm() + 1e-30))
    out = {"name": name, "max_abs": float(diff.max()), "rel_l2": rel_l2, "n_bad": nbad, "n": got.numel(),
           "finite": finite}
    if nbad or not finite:
        idx = bad.nonzero()[:8].tolist()
        flat = bad.reshape(-1, bad.shape[-1]) if bad.dim() > 1 else bad.reshape(1, -1)
        rows_bad = flat.any(1).nonzero().flatten()[:16].tolist()
        cols_bad = flat.any(0).nonzero().flatten()[:16].tolist()
        raise AssertionError(
            f"{name}: {nbad}/{got.numel()} outside tol (atol={atol}, rtol={rtol}); finite={finite}; "
            f"max_abs={out['max_abs']:.4g} rel_l2={rel_l2:.4g}; first bad idx={idx}; bad rows~{rows_bad}; "
            f"bad cols~{cols_bad}; got={[float(got[tuple(i)]) for i in idx[:4]]} ref={[float(ref[tuple(i)]) for i in idx[:4]]}")
    return out


def _bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _f32(*shape, scale=1.0, seed=0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(DEV)


# ----------------------------------------------------------------------------------------------------- row kernels
def check_cast(n=3 * 16 * 224 * 224 + 5):
    x = _f32(n, seed=1)
    y = L.cast_f32_bf16(x)
    torch.cuda.synchronize()
    assert torch.equal(y, x.to(torch.bfloat16)), "cast: not bit-identical to torch's round-to-nearest-even"
    return {"name": "cast", "n": n}


def check_row_stats(M=1000, D=768, eps=1e-6):
    x = (_f32(M, D, seed=2, scale=1.5, shift=0.3)).to(torch.bfloat16)
    st = L.row_stats(x, eps)
    torch.cuda.synchronize()
    xf = x.float()
    mean = xf.mean(1)
    rstd = (xf.var(1, unbiased=False) + eps).rsqrt()
    a = _stats(st[:, 0], mean, f"row_stats.mean[{M}x{D}]", 1e-5, 1e-5)
    b = _stats(st[:, 1], rstd, f"row_stats.rstd[{M}x{D}]", 1e-5, 1e-4)
    return {"name": "row_stats", "mean": a, "rstd": b}


def check_layernorm(M=777, D=768, eps=1e-6):
    x = (_f32(M, D, seed=3, scale=2.0, shift=-0.2)).to(torch.bfloat16)
    g = _f32(D, seed=4, scale=0.2, shift=1.0)
    b = _f32(D, seed=5, scale=0.2)
    y = L.layernorm(x, g, b, eps)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), g, b, eps)
    return _stats(y, ref, f"layernorm[{M}x{D}]", 1e-4, 1e-4)


def check_pool_head(B=5, N=1568, D=768, Cn=2, eps=1e-6):
    x = (_f32(B, N, D, seed=6, scale=1.0, shift=0.1)).to(torch.bfloat16)
    g = _f32(D, seed=7, scale=0.2, shift=1.0)
    b = _f32(D, seed=8, scale=0.2)
    w = _f32(Cn, D, seed=9, scale=0.05)
    bh = _f32(Cn, seed=10, scale=0.1)
    logits, probs, feats = L.pool_norm_head(x, g, b, w, bh, eps, want_probs=True, want_features=True)
    torch.cuda.synchronize()
    pooled = x.float().mean(1)
    normed = torch.nn.functional.layer_norm(pooled, (D,), g, b, eps)
    ref = normed @ w.t() + bh
    a = _stats(logits, ref, f"pool_head.logits[B{B},N{N},D{D}]", 2e-4, 2e-4)
    p = _stats(probs, ref.softmax(-1), "pool_head.probs", 1e-4, 1e-4)
    f = _stats(feats, normed, "pool_head.features", 1e-4, 1e-4)
    return {"name": "pool_head", "logits": a, "probs": p, "features": f}


def check_rows_norm_head(B=3, S=1568, D=768, Cn=2, rows="all", eps=1e-6, seed=60):
    """final_reduction 'cls' / 'none' (mf:323-334): LayerNorm of token 0 / every token, head, softmax."""
    x = (_f32(B, S, D, seed=seed, scale=1.2, shift=0.2)).to(torch.bfloat16)
    g = _f32(D, seed=seed + 1, scale=0.2, shift=1.0)
    b = _f32(D, seed=seed + 2, scale=0.2)
    w = _f32(Cn, D, seed=seed + 3, scale=0.05)
    bh = _f32(Cn, seed=seed + 4, scale=0.1)
    logits, probs, feats = L.rows_norm_head(x, g, b, w, bh, eps, rows=rows, want_probs=True, want_features=True)
    torch.cuda.synchronize()
    xf = x.float()[:, 0] if rows == "cls" else x.float()
    normed = torch.nn.functional.layer_norm(xf, (D,), g, b, eps)
    ref = normed @ w.t() + bh
    tag = f"rows_norm_head[{rows},B{B},S{S},D{D},C{Cn}]"
    a = _stats(logits, ref, tag + ".logits", 2e-4, 2e-4)
    p = _stats(probs, ref.softmax(-1), tag + ".probs", 1e-4, 1e-4)
    f = _stats(feats, normed, tag + ".features", 1e-4, 1e-4)
    return {"name": "rows_norm_head", "logits": a, "probs": p, "features": f}


def check_prepend_cls(B=3, N=1568, D=384, eps=1e-6, seed=70):
    """x = cat(cls_token, emb) per clip (MVD mf:431-435) + LayerNorm statistics of every row."""
    emb = _bf16(B, N, D, seed=seed, scale=1.1)
    cls = _f32(D, seed=seed + 1, scale=0.5)
    x, st = L.prepend_cls(emb, cls, eps)
    torch.cuda.synchronize()
    ref = torch.cat([cls.to(torch.bfloat16)[None, None].expand(B, 1, D), emb], dim=1)
    assert torch.equal(x, ref), f"prepend_cls[{B}x{N}x{D}]: rows are not bit-identical to torch.cat"
    xf = ref.float().reshape(B * (N + 1), D)
    a = _stats(st[:, 0], xf.mean(1), f"prepend_cls.mean[{B}x{N}x{D}]", 1e-5, 1e-5)
    b = _stats(st[:, 1], (xf.var(1, unbiased=False) + eps).rsqrt(), f"prepend_cls.rstd[{B}x{N}x{D}]", 1e-5, 1e-4)
    return {"name": "prepend_cls", "mean": a, "rstd": b}


# ----------------------------------------------------------------------------------------------------------- GEMMs
def check_gemm(M=1568, N=768, K=768, mode="plain", seed=0):
    """mode: plain | bias | resid | ln | ln_gelu"""
    a = _bf16(M, K, seed=seed + 11, scale=1.0)
    w = _bf16(N, K, seed=seed + 12, scale=0.05)
    bias = _f32(N, seed=seed + 13, scale=0.5)
    af, wf = a.float(), w.float()
    if mode in ("plain", "bias", "resid"):
        res = _bf16(M, N, seed=seed + 14) if mode == "resid" else None
        out = L.gemm_bias_residual(a, w, None if mode == "plain" else bias, res)
        torch.cuda.synchronize()
        ref = af @ wf.t()
        if mode != "plain":
            ref = ref + bias
        if mode == "resid":
            ref = ref + res.float()
    else:
        eps = 1e-6
        # x has a per-row offset so the folded mean term matters
        a = (a.float() + _f32(M, 1, seed=seed + 15, scale=0.5)).to(torch.bfloat16)
        af = a.float()
        stats = L.row_stats(a, eps)
        colsum = wf.sum(1).contiguous()
        out = L.ln_gemm(a, stats, w, bias, colsum, gelu=(mode == "ln_gelu"))
        torch.cuda.synchronize()
        mean = af.mean(1, keepdim=True)
        rstd = (af.var(1, unbiased=False, keepdim=True) + eps).rsqrt()
        ref = ((af - mean) * rstd) @ wf.t() + bias
        if mode == "ln_gelu":
            ref = torch.nn.functional.gelu(ref)
    # bf16 output rounding: 2^-8 relative; fp32 accumulation error is far smaller
    return _stats(out, ref, f"gemm.{mode}[{M}x{N}x{K}]", 2e-2, 1e-2)


def check_gemm_stats(M=3136, D=768, N2=3072, K1=768, seed=0):
    """residual GEMM whose epilogue emits the LayerNorm statistics of the rows it stores -> LN-folded GEMM, against
    layer_norm(x) @ w2^T in fp32 on the bf16 x the first GEMM stored; the statistics also against the two-pass
    statistics kernel."""
    eps = 1e-6
    a = _bf16(M, K1, seed=seed + 41)
    w1 = _bf16(D, K1, seed=seed + 42, scale=0.05)
    b1 = _f32(D, seed=seed + 43, scale=0.5)
    res = (_f32(M, D, seed=seed + 44) + _f32(M, 1, seed=seed + 45, scale=2.0)).to(torch.bfloat16)  # per-row offset
    x, st = L.gemm_bias_residual_stats(a, w1, b1, res, eps)
    w2 = _bf16(N2, D, seed=seed + 46, scale=0.05)
    b2 = _f32(N2, seed=seed + 47, scale=0.5)
    colsum = w2.float().sum(1).contiguous()
    y = L.ln_gemm(x, st, w2, b2, colsum, gelu=True)
    st2 = L.row_stats(x, eps)
    torch.cuda.synchronize()
    xr = a.float() @ w1.float().t() + b1 + res.float()
    r1 = _stats(x, xr, f"gemm_stats.x[{M}x{D}x{K1}]", 3e-2, 1e-2)
    xf = x.float()
    s1 = _stats(st[:, 0], xf.mean(1), "gemm_stats.mean", 1e-4, 1e-4)
    s2 = _stats(st[:, 1], (xf.var(1, unbiased=False) + eps).rsqrt(), "gemm_stats.rstd", 1e-4, 2e-4)
    s3 = _stats(st, st2, "gemm_stats vs row_stats kernel", 1e-4, 2e-4)
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xf, (D,), None, None, eps) @ w2.float().t() + b2)
    r2 = _stats(y, ref, f"gemm_stats.ln_gelu[{M}x{N2}x{D}]", 2e-2, 1e-2)
    return {"name": "gemm_stats", "x": r1, "mean": s1, "rstd": s2, "vs_kernel": s3, "y": r2}


# ------------------------------------------------------------------------------------------------------- attention
def check_attention_outliers(B=2, H=2, S=1568, seed=0, gains=(8.0, 16.0)):
    """Adversarial scores for the lazy reference max: the reference of a row is the exact max of the unit's FIRST key tile
    (the ragged tail of the sequence, 32 keys at S = 1568).  Every 5th query row gets keys planted in later tiles that
    score 90+ and 180+ octaves above everything before them (k_j = gain * q_row), in increasing order, so that the row
    sum guard trips once or twice in the same unit (slow path: exact max of the tile, O and row sum rescaled, tile
    redone); every 7th row gets its dominant key INSIDE the first tile.  The fp32 reference puts ~all weight on the last
    planted key."""
    qkv = _bf16(B, S, 3, H, 64, seed=seed + 31, scale=1.0)
    g = torch.Generator().manual_seed(seed + 32)
    n_tiles = (S + 95) // 96
    for b in range(B):
        for h in range(H):
            for r in range(0, S, 5):
                # two planted keys at increasing positions in DIFFERENT key tiles (tile order of the kernel: ragged tail
                # first, then keys 0, 96, 192, ...), never in the ragged tail itself
                full = (S // 96) * 96
                if full < 192:
                    continue
                j1 = int(torch.randint(0, full // 2, (1,), generator=g))
                j2 = int(torch.randint(full // 2, full, (1,), generator=g))
                qkv[b, j1, 1, h] = (gains[0] * qkv[b, r, 0, h].float()).to(torch.bfloat16)
                qkv[b, j2, 1, h] = (gains[1] * qkv[b, r, 0, h].float()).to(torch.bfloat16)
            if S % 96:
                for r in range(3, S, 7):
                    j0 = (S // 96) * 96 + int(torch.randint(0, S % 96, (1,), generator=g))
                    qkv[b, j0, 1, h] = (4.0 * qkv[b, r, 0, h].float()).to(torch.bfloat16)
    out = L.attention(qkv)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))  # [B,H,S,64]
    att = (q * 64 ** -0.5) @ k.transpose(-1, -2)
    ref = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, S, H * 64)
    return _stats(out, ref, f"attention_outliers[B{B},H{H},S{S}]", 2e-2, 2e-2)


def check_attention(B=2, H=3, S=1568, peaky=1.0, seed=0):
    qkv = _bf16(B, S, 3, H, 64, seed=seed + 21, scale=1.0)
    if peaky != 1.0:
        qkv[:, :, 0] *= peaky  # sharper softmax -> exercises the running-max rescale
    out = L.attention(qkv)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))  # [B,H,S,64]
    att = (q * 64 ** -0.5) @ k.transpose(-1, -2)
    ref = (att.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, S, H * 64)
    return _stats(out, ref, f"attention[B{B},H{H},S{S},peaky{peaky}]", 2e-2, 2e-2)


# ----------------------------------------------------------------------------------------------------- patch embed
def _sinusoid(n, d):
    pos = torch.arange(n, dtype=torch.float64)[:, None]
    j = torch.arange(d, dtype=torch.float64)[None, :]
    ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / d)
    tab = torch.where((torch.arange(d) % 2 == 0)[None, :], ang.sin(), ang.cos())
    return tab.float()


def check_patch_embed(B=2, D=384, mode="clips", masked=False, seed=0, T=16, tubelet=2, img=224, frame_step=1):
    Cc, Hh, Ww = 3, img, img
    K = Cc * tubelet * 16 * 16
    N = (T // tubelet) * (img // 16) ** 2
    w5 = _bf16(D, Cc, tubelet, 16, 16, seed=seed + 31, scale=0.03)
    bias = _f32(D, seed=seed + 32, scale=0.2)
    pos_bias = (_sinusoid(N, D).to(DEV) + bias).contiguous()
    dims = L.make_dims(img_h=img, img_w=img, tubelet=tubelet, frames=T, dim=D, depth=1, heads=D // 64, hidden=4 * D)
    if mode == "clips":
        x = _bf16(B, Cc, T, Hh, Ww, seed=seed + 33)
        clips = x
        kw = dict(mode=L.STAD_IN_CLIPS)
    else:
        span = (T - 1) * frame_step + 1  # frame t of clip b = frames[start + b*stride + t*frame_step]
        F, start, stride = span + 3 * (B - 1) + 2, 1, 3
        frames = _bf16(F, Cc, Hh, Ww, seed=seed + 34)
        x = frames
        clips = torch.stack([frames[start + b * stride: start + b * stride + span: frame_step] for b in range(B)])
        clips = clips.permute(0, 2, 1, 3, 4).contiguous()  # [B,T,C,H,W] -> [B,C,T,H,W]
        kw = dict(mode=L.STAD_IN_FRAMES, n_frames=F, start=start, stride=stride, frame_step=frame_step)
    ref = torch.nn.functional.conv3d(clips.float(), w5.float(), None, stride=(tubelet, 16, 16))  # [B,D,8,14,14]
    ref = ref.flatten(2).transpose(1, 2) + pos_bias  # [B,N,D]
    tok_idx = None
    n_tok = N
    if masked:
        g = torch.Generator().manual_seed(seed + 35)
        keep = torch.stack([torch.randperm(196, generator=g)[:20].sort().values for _ in range(B)])  # [B,20]
        tok = (torch.arange(8)[None, :, None] * 196 + keep[:, None, :]).reshape(B, -1)                # [B,160]
        tok_idx = tok.to(torch.int32).to(DEV).contiguous()
        n_tok = tok.shape[1]
        ref = torch.gather(ref, 1, tok.to(DEV)[:, :, None].expand(-1, -1, D))
    out = L.patch_embed(x, w5.reshape(D, K).contiguous(), pos_bias, dims, B, n_tok, tok_idx=tok_idx, **kw)
    torch.cuda.synchronize()
    return _stats(out.reshape(B, n_tok, D), ref,
                  f"patch_embed[{mode},masked={masked},B{B},D{D},T{T},tubelet{tubelet},img{img}]", 2e-2, 1e-2)


# -------------------------------------------------------------------------------- MAE decoder glue / frame preparation
def check_decoder_assemble(B=3, N=1568, n_vis=160, D=384, eps=1e-6, seed=40):
    """x_full = cat(vis, mask_token + pos[mask_idx]) (mp:283-288) + LayerNorm statistics of every row."""
    vis = _bf16(B, n_vis, D, seed=seed, scale=1.3)
    pos = _f32(N, D, seed=seed + 1)
    mask_token = _f32(D, seed=seed + 2, scale=0.5)
    g = torch.Generator().manual_seed(seed + 3)
    mask_idx = torch.stack([torch.randperm(N, generator=g)[: N - n_vis].sort().values for _ in range(B)])
    x, st = L.decoder_assemble(vis, pos, mask_token, mask_idx.to(torch.int32).to(DEV), N, eps)
    torch.cuda.synchronize()
    ref = torch.cat([vis, (mask_token[None, None] + pos[mask_idx.to(DEV)]).to(torch.bfloat16)], dim=1)
    assert torch.equal(x, ref), f"decoder_assemble[{B}x{N}x{D}]: rows are not bit-identical to the torch composition"
    xf = ref.float().reshape(B * N, D)
    a = _stats(st[:, 0], xf.mean(1), f"decoder_assemble.mean[{B}x{N}x{D}]", 1e-5, 1e-5)
    b = _stats(st[:, 1], (xf.var(1, unbiased=False) + eps).rsqrt(), f"decoder_assemble.rstd[{B}x{N}x{D}]", 1e-5, 1e-4)
    return {"name": "decoder_assemble", "mean": a, "rstd": b}


def check_tail_rows(B=3, N=200, n_keep=171, Cn=1536, seed=44):
    x = _bf16(B, N, Cn, seed=seed)
    y = L.tail_rows_f32(x, n_keep)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and torch.equal(y, x[:, N - n_keep:].float()), "tail_rows_f32: not an exact copy"
    return {"name": "tail_rows", "n": y.numel()}


def check_normalize_u8(F_=5, H=224, W=224, bgr=True, seed=46):
    """prepare_image (ri:15-34): BGR uint8 HWC -> RGB CHW, /255, ImageNet mean/std — evaluated in fp32 by torch."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (F_, H, W, 3), generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    out = L.normalize_frames_u8(u8.to(DEV), mean, std, bgr=bgr)
    torch.cuda.synchronize()
    rgb = u8.flip(-1) if bgr else u8
    img = rgb.permute(0, 3, 1, 2).float().div_(255.0)
    ref = (img - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    assert out.shape == (F_, 3, H, W) and out.dtype == torch.bfloat16
    # the kernel evaluates v * (1/(255 std)) - mean/std in one FMA; the two fp32 evaluations differ by <= 1 ulp of
    # fp32, which can flip a bf16 rounding in rare ties: allow one bf16 ulp
    return _stats(out.cpu(), ref, f"normalize_u8[{F_}x{H}x{W},bgr={bgr}]", 1e-3, 4e-3)


def check_resize_cubic():
    """stad_resize_cubic_u8 against the oracle's integer restatement of cv2 INTER_CUBIC (bit-exact) and against the
    cv2 outputs of the fixture (<= 1 LSB on < 1e-4 of the pixels: OpenCV's float SIMD pass)."""
    import numpy as np
    from oracle import resize_oracle as ro
    from simple_tad_b200 import frames
    from tests import parity
    g = parity.golden("resize_cubic")
    out = []
    for i in range(3):
        h, w, dh, dw = (int(v) for v in g[f"shape_{i}"])
        imgs = np.stack([ro.synthetic_frame(h, w, seed=i), ro.synthetic_frame(h, w, seed=i + 10)])
        got = frames.resize_cubic_u8(torch.from_numpy(imgs).to(DEV), (dh, dw)).cpu().numpy()
        for k in range(2):
            ref = ro.resize_cubic_u8(imgs[k], dh, dw)
            assert np.array_equal(got[k], ref), f"resize_cubic[{h}x{w}->{dh}x{dw}]: {(got[k] != ref).sum()} pixels differ from the oracle"
        d = np.abs(got[0].astype(int) - g[f"cv2_{i}"].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-4
        out.append({"name": f"resize_cubic[{h}x{w}->{dh}x{dw}]", "vs_cv2_mismatch": float((d > 0).mean())})
    # upscale and identity sizes, odd sizes
    for (h, w, dh, dw) in ((50, 70, 224, 224), (224, 224, 224, 224), (33, 17, 5, 9)):
        img = ro.synthetic_frame(h, w, seed=7)
        got = frames.resize_cubic_u8(torch.from_numpy(img[None]).to(DEV), (dh, dw)).cpu().numpy()[0]
        assert np.array_equal(got, ro.resize_cubic_u8(img, dh, dw)), (h, w, dh, dw)
    return out


def check_gemm_ln_pos(M=320, N=384, K=768, n_rows=1568, eps=1e-6, seed=48):
    """encoder_to_decoder with the encoder norm folded in and the gathered decoder position rows added (mp:107,281,287):
    only reachable through stad_mae_forward, so checked there; here the LN-fold + plain GEMM pieces on the same shape."""
    return check_gemm(M, N, K, "ln", seed=seed)


def check_gemm_pair():
    """CTA-pair (cta_group::2, 256 x 256) tiles: every epilogue that has a pair instantiation, on shapes the dispatcher
    sends there (K >= 2048, even number of M-tiles, >= 74 pair tiles), incl. a ragged last M-tile pair and several
    tiles per pair (persistent loop, accumulator double-buffering across the two CTAs)."""
    out = []
    for M, N, K, mode in ((9472, 512, 2048, "plain"), (9472, 512, 2048, "bias"), (9472, 768, 3072, "resid"),
                          (9472, 512, 2048, "ln"), (9472, 1024, 2048, "ln_gelu"), (25088, 768, 3072, "resid"),
                          (9472 - 70, 512, 2048, "ln"), (18944, 1024, 2048, "ln_gelu"), (9472, 512, 4096, "resid")):
        out.append(check_gemm(M, N, K, mode, seed=60 + len(out)))
    out.append(check_gemm_stats(9472, 768, 512, 3072, seed=7))      # first GEMM: resid + stats on the pair tile
    out.append(check_gemm_stats(25088 - 128 - 5, 1024, 256, 4096, seed=8))
    # odd number of M-tiles: the last pair's second M-tile is virtual (zero-filled loads, clipped stores); 16000 rows =
    # the DAPT encoder batch (100 clips x 160 visible tokens), 100416 = 64 clips x 1569 tokens (MVD class token)
    for M, N, K, mode in ((9600, 512, 2048, "plain"), (9600 - 37, 512, 768, "ln"), (16000, 3072, 768, "ln_gelu"),
                          (16000, 768, 3072, "resid"), (9600 - 128 - 1, 768, 768, "bias")):
        out.append(check_gemm(M, N, K, mode, seed=80 + len(out)))
    out.append(check_gemm_stats(16000, 768, 2304, 768, seed=9))
    out.append(check_gemm_stats(9600 - 91, 1024, 512, 4096, seed=10))
    # N = 384 / 1152 (ViT-S, the MAE decoder) at pair-tile row counts: these run single-CTA 192-column tiles (256 x 192
    # pair tiles measured slower, gemm.cu kPairTiles192): three 32-column chunks per epilogue warpgroup
    for M, N, K, mode in ((9472, 1152, 384, "plain"), (9472, 1152, 384, "ln"), (25088, 384, 384, "bias"),
                          (9472, 384, 1536, "resid"), (9600 - 37, 1152, 384, "ln_gelu"), (50176, 1152, 384, "ln"),
                          (50176, 384, 1536, "resid")):
        out.append(check_gemm(M, N, K, mode, seed=100 + len(out)))
    out.append(check_gemm_stats(9472, 384, 1152, 1536, seed=11))
    out.append(check_gemm_stats(31360 - 5, 384, 1152, 384, seed=12))
    return out


# ------------------------------------------------------------------------------- full-size, size-independent properties
def check_gemm_integer_exact(M=64 * 1568, N=2304, K=768, mode="bias", seed=0):
    """The bench step's GEMM shapes with small-integer operands: every product and every partial sum is an integer
    below 2^24 and every result an integer of magnitude <= 256, so fp32 accumulation in ANY order and the bf16 store are
    exact — the kernel must reproduce the fp32 matmul bit for bit on all M x N outputs (a dropped k-block, a tile
    written twice or a mis-addressed row in any of the ~3600 tiles shows as an integer error)."""
    g = torch.Generator(device=DEV).manual_seed(seed + 500)

    def sparse_pm1(rows, cols):  # entries in {-1, 0, 1}, three quarters of them zero
        t = torch.randint(-1, 2, (rows, cols), generator=g, device=DEV, dtype=torch.int8)
        keep = torch.rand(rows, cols, generator=g, device=DEV) < 0.25
        return (t * keep).to(torch.bfloat16)

    a, w = sparse_pm1(M, K), sparse_pm1(N, K)
    bias = torch.randint(-3, 4, (N,), generator=g, device=DEV).float()
    res = torch.randint(-8, 9, (M, N), generator=g, device=DEV).to(torch.bfloat16) if mode == "resid" else None
    out = L.gemm_bias_residual(a, w, bias, res)
    torch.cuda.synchronize()
    worst = 0.0
    for r0 in range(0, M, 16384):  # fp32 reference by row blocks (the full product would be another M x N fp32 buffer)
        ref = a[r0:r0 + 16384].float() @ w.float().t() + bias
        if res is not None:
            ref = ref + res[r0:r0 + 16384].float()
        assert float(ref.abs().max()) <= 256.0
        d = (out[r0:r0 + 16384].float() - ref).abs().max()
        worst = max(worst, float(d))
    assert worst == 0.0, f"gemm.{mode}[{M}x{N}x{K}] integer operands: max |diff| = {worst} (must be exact)"
    return {"name": f"gemm_integer_exact.{mode}[{M}x{N}x{K}]", "max_abs": worst}


def check_attention_full_size_properties(B=64, H=12, S=1568, seed=0):
    """The bench step's attention launch (5376 units on 148 CTAs), through properties that need no S x S reference:
    (1) rows of softmax sum to one: with V constant along the keys the output equals that constant, whatever Q and K;
    (2) a joint permutation of the keys of K and V leaves the output unchanged up to rounding (different tiling of the
        same sum);
    (3) sampled (clip, head) pairs against the fp32 softmax(Q K^T) V."""
    qkv = _bf16(B, S, 3, H, 64, seed=seed + 600, scale=1.0)
    out = L.attention(qkv)
    # (1)
    const = _bf16(B, 1, H, 64, seed=seed + 601)
    qkv_c = qkv.clone()
    qkv_c[:, :, 2] = const
    out_c = L.attention(qkv_c)
    torch.cuda.synchronize()
    want = const.float().reshape(B, 1, H * 64).expand(B, S, H * 64)
    s1 = _stats(out_c, want, f"attention.const_v[B{B},H{H},S{S}]", 1e-3, 1e-2)
    # (2)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(seed + 602)).to(DEV)
    qkv_p = qkv.clone()
    qkv_p[:, :, 1:] = qkv[:, perm, 1:]
    out_p = L.attention(qkv_p)
    torch.cuda.synchronize()
    s2 = _stats(out_p, out, f"attention.key_permutation[B{B},H{H},S{S}]", 1e-2, 2e-2)
    # (3)
    g = torch.Generator().manual_seed(seed + 603)
    for _ in range(6):
        b, h = int(torch.randint(0, B, (1,), generator=g)), int(torch.randint(0, H, (1,), generator=g))
        q, k, v = (qkv[b, :, i, h].float() for i in range(3))
        ref = ((q * 64 ** -0.5) @ k.t()).softmax(-1) @ v
        _stats(out[b, :, h * 64:(h + 1) * 64], ref, f"attention.sample[b{b},h{h}]", 2e-2, 2e-2)
    return [s1, s2]



CHECKS = {
    "gemm_integer_exact": lambda: [check_gemm_integer_exact(64 * 1568, 2304, 768, "bias"),
                                   check_gemm_integer_exact(64 * 1568, 768, 3072, "resid", seed=1),
                                   check_gemm_integer_exact(128 * 1568, 384, 384, "resid", seed=2)],
    "attention_full_size": check_attention_full_size_properties,
    "rows_norm_head": lambda: [check_rows_norm_head(3, 1568, 768, 2, "all"), check_rows_norm_head(5, 1569, 384, 2, "cls"),
                               check_rows_norm_head(2, 100, 1024, 40, "all", seed=61),
                               check_rows_norm_head(4, 7, 384, 400, "cls", seed=62)],
    "prepend_cls": lambda: [check_prepend_cls(), check_prepend_cls(2, 196, 1024, seed=71)],
    # sequence lengths of the sibling models: 1569 = 1568 + class token (33-row / 33-key tails), 3136, 4608
    "attention_siblings": lambda: [check_attention(2, 6, 1569, seed=7), check_attention(1, 6, 3136, seed=8),
                                   check_attention(1, 3, 4608, peaky=3.0, seed=9)],
    "gemm_pair": check_gemm_pair,
    "resize_cubic": check_resize_cubic,
    "decoder_assemble": lambda: [check_decoder_assemble(), check_decoder_assemble(2, 1568, 392, 192, seed=50),
                                 check_decoder_assemble(1, 1568, 160, 512, seed=51)],
    "tail_rows": lambda: [check_tail_rows(), check_tail_rows(2, 1568, 1408, 1536)],
    "normalize_u8": lambda: [check_normalize_u8(), check_normalize_u8(3, 224, 224, bgr=False), check_normalize_u8(2, 720, 1280)],
    "gemm_e2d_shapes": lambda: [check_gemm_ln_pos(), check_gemm(320, 192, 384, "ln", seed=49),
                                check_gemm(3136, 1536, 384, "ln", seed=50), check_gemm(3136, 1536, 192, "ln", seed=51),
                                check_gemm(3136, 576, 192, "ln", seed=52), check_gemm(3136, 192, 768, "resid", seed=53)],
    "cast": lambda: check_cast(),
    "row_stats": lambda: [check_row_stats(1000, 768), check_row_stats(333, 384), check_row_stats(129, 1024)],
    "layernorm": lambda: check_layernorm(),
    "pool_head": lambda: [check_pool_head(5, 1568, 768), check_pool_head(3, 160, 384), check_pool_head(2, 1568, 1024)],
    "gemm_plain_small": lambda: check_gemm(128, 64, 64, "plain"),
    "gemm_plain_k": lambda: check_gemm(128, 128, 768, "plain"),
    "gemm_plain": lambda: [check_gemm(1568, 768, 768, "plain"), check_gemm(6272, 2304, 768, "plain"),
                           check_gemm(25088, 768, 3072, "plain")],
    "gemm_bn": lambda: [check_gemm(4 * 1568, 384, 384, "bias"), check_gemm(40 * 1568, 1152, 384, "bias"),
                        check_gemm(40 * 1568, 1024, 1024, "bias"), check_gemm(300, 1536, 384, "bias")],
    "gemm_resid": lambda: [check_gemm(3136, 768, 768, "resid"), check_gemm(1568 * 3, 1024, 4096, "resid")],
    "gemm_ln": lambda: [check_gemm(3136, 2304, 768, "ln"), check_gemm(1568, 1152, 384, "ln")],
    "gemm_ln_gelu": lambda: [check_gemm(3136, 3072, 768, "ln_gelu"), check_gemm(1568, 4096, 1024, "ln_gelu")],
    "gemm_stats": lambda: [check_gemm_stats(3136, 768, 3072, 768), check_gemm_stats(1568, 384, 1152, 1536, seed=3),
                           check_gemm_stats(200, 1024, 1024, 4096, seed=5)],
    # planted keys far above the lazy reference (slow path: guard trip, rescale, redo), also in multi-unit CTAs, a
    # class-token length and a key-split tail; S = 160 / 392: two / five tiles per unit
    "attention_outliers": lambda: [check_attention_outliers(2, 2, 1568), check_attention_outliers(30, 6, 392, seed=2),
                                   check_attention_outliers(1, 3, 1569, seed=3), check_attention_outliers(80, 4, 288, seed=4),
                                   check_attention_outliers(3, 12, 1568, seed=5, gains=(30.0, 200.0))],
    "attention_small": lambda: check_attention(1, 1, 128),
    "attention_tail": lambda: [check_attention(1, 2, 160), check_attention(2, 1, 392)],
    "attention": lambda: [check_attention(2, 3, 1568), check_attention(1, 12, 1568, peaky=6.0)],
    # ragged shapes: a single K/V tile, 1-key / 127-key tails, one- and two-slot units, query tails in every warp
    "attention_ragged": lambda: [check_attention(1, 1, 32), check_attention(2, 2, 100), check_attention(1, 2, 129),
                                 check_attention(1, 1, 255), check_attention(2, 1, 257), check_attention(1, 3, 300),
                                 check_attention(1, 2, 640, peaky=4.0)],
    # one row, one row short of / past a 128-row tile, N with 64 as its only tile width (320), a single k-block
    "gemm_edges": lambda: [check_gemm(M, N, K, mode, seed=40 + i) for i, (M, N, K, mode) in enumerate(
        ((1, 64, 64, "bias"), (127, 320, 128, "resid"), (129, 192, 64, "ln"), (255, 320, 192, "ln_gelu"),
         (1, 768, 768, "resid"), (385, 64, 3072, "plain")))],
    # sequence lengths at the seams of the unit / tile geometry: fewer keys than one 32-key chunk, a 1-key ragged tile
    # (97, 193), exact multiples of the 96-key tile, a 33-row tail (one row too many for the key-split unit), 1 token
    "attention_edges": lambda: [check_attention(2, 2, S, seed=30 + i) for i, S in
                                enumerate((1, 7, 31, 33, 64, 96, 97, 128, 192, 193, 289, 384))],
    # more work units than SMs: every CTA of the persistent kernel loops over several units (phase bookkeeping)
    "attention_persistent": lambda: [check_attention(5, 12, 1568, seed=3), check_attention(40, 12, 160, seed=4),
                                     check_attention(16, 6, 392, peaky=5.0, seed=5)],
    "patch_embed": lambda: [check_patch_embed(2, 384, "clips"), check_patch_embed(3, 768, "clips")],
    "patch_embed_frames": lambda: [check_patch_embed(3, 384, "frames"),
                                   # in-window frame step 3: a 30 fps video scored at 10 fps (sequencing.py:45-58)
                                   check_patch_embed(3, 384, "frames", frame_step=3, seed=7),
                                   check_patch_embed(2, 768, "frames", masked=True, frame_step=2, seed=8)],
    # the UMT sibling's geometries: tubelet 1 (K = 768) on 8 and 16 frames, a 384 px image (24 x 24 grid, 4 h' per tile)
    "patch_embed_siblings": lambda: [check_patch_embed(2, 768, "clips", T=8, tubelet=1, seed=3),
                                     check_patch_embed(1, 384, "clips", T=16, tubelet=1, seed=4),
                                     check_patch_embed(1, 384, "clips", T=8, tubelet=1, img=384, seed=5),
                                     check_patch_embed(2, 384, "frames", T=8, tubelet=1, seed=6)],
    "patch_embed_masked": lambda: [check_patch_embed(2, 768, "clips", masked=True),
                                   check_patch_embed(2, 384, "frames", masked=True)],
}
