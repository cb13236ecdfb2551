"""ctypes binding of libstad.so (include/stad.h).  PyTorch is plumbing here: it owns device memory and streams and
hands raw pointers to the C ABI.  There is no fallback: if the library is missing or the device is not sm_100 every
entry point raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STAD_LIB") or os.path.join(_HERE, "libstad.so")  # STAD_LIB: development A/B builds

STAD_OK, STAD_E_SHAPE, STAD_E_ALIGN, STAD_E_ARCH, STAD_E_CUDA = 0, -1, -2, -3, -4
STAD_EPI_BIAS, STAD_EPI_BIAS_GELU = 0, 1
STAD_IN_CLIPS, STAD_IN_FRAMES = 0, 1
STAD_REDUCE_MEAN, STAD_REDUCE_CLS, STAD_REDUCE_NONE = 0, 1, 2

EXPORTS = (
    "stad_abi_version", "stad_init", "stad_last_error", "stad_cast_f32_bf16", "stad_row_stats", "stad_layernorm",
    "stad_pool_norm_head", "stad_patch_embed", "stad_ln_gemm", "stad_gemm_bias_residual", "stad_attention",
    "stad_workspace_bytes", "stad_vit_forward", "stad_profile_enable", "stad_profile_read", "stad_stat_parts",
    "stad_gemm_bias_residual_stats", "stad_stats_finalize", "stad_decoder_assemble", "stad_tail_rows_f32",
    "stad_mae_workspace_bytes", "stad_mae_forward", "stad_normalize_frames_u8", "stad_eval_hist",
    "stad_resize_cubic_u8", "stad_rows_norm_head", "stad_prepend_cls",
)


class StadInput(C.Structure):
    _fields_ = [("data", C.c_void_p), ("mode", C.c_int32), ("n_frames", C.c_int32), ("start", C.c_int32),
                ("stride", C.c_int32), ("frame_step", C.c_int32), ("tubelet_reuse", C.c_int32),
                ("window_starts", C.c_void_p)]


class StadDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("img_h", "img_w", "patch", "tubelet", "frames", "in_chans", "dim", "depth",
                                          "heads", "hidden", "num_classes")]


class StadBlock(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "b_qkv", "cs_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "cs_fc1",
                                           "w_fc2", "b_fc2")]


class StadModel(C.Structure):
    _fields_ = [("dims", StadDims), ("w_patch", C.c_void_p), ("pos_bias", C.c_void_p), ("blocks", C.POINTER(StadBlock)),
                ("norm_g", C.c_void_p), ("norm_b", C.c_void_p), ("w_head", C.c_void_p), ("b_head", C.c_void_p),
                ("eps", C.c_float), ("attn_scale", C.c_float), ("reduction", C.c_int32), ("cls_token", C.c_void_p)]


class StadMaeModel(C.Structure):
    _fields_ = [("encoder", StadModel), ("dec_dims", StadDims), ("w_e2d", C.c_void_p), ("b_e2d", C.c_void_p),
                ("cs_e2d", C.c_void_p), ("pos_dec", C.c_void_p), ("mask_token", C.c_void_p),
                ("dec_blocks", C.POINTER(StadBlock)), ("w_pix", C.c_void_p), ("b_pix", C.c_void_p),
                ("cs_pix", C.c_void_p)]


class StadOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("logits", "probs", "features", "tokens")]


class StadProfileRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("epi", C.c_int32), ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
                ("ms", C.c_float)]


KIND_NAMES = {0: "cast", 1: "gather", 2: "gemm", 3: "attention", 4: "row_stats", 5: "layernorm", 6: "pool",
              7: "assemble", 8: "tail", 9: "normalize", 10: "eval", 11: "resize"}

_lib = None
_inited_devices = set()


def load():
    """dlopen libstad.so (no CUDA call is made here, so this also works on a CPU-only box)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(simple-tad_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    protos = {
        "stad_abi_version": (C.c_int, []),
        "stad_init": (C.c_int, [i32]),
        "stad_last_error": (C.c_char_p, []),
        "stad_cast_f32_bf16": (C.c_int, [vp, vp, sz, vp]),
        "stad_row_stats": (C.c_int, [vp, vp, i32, i32, f32, vp]),
        "stad_layernorm": (C.c_int, [vp, vp, vp, vp, i32, i32, f32, vp]),
        "stad_pool_norm_head": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
        "stad_patch_embed": (C.c_int, [C.POINTER(StadInput), vp, vp, vp, vp, vp, C.POINTER(StadDims), i32, i32, vp]),
        "stad_ln_gemm": (C.c_int, [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp]),
        "stad_gemm_bias_residual": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
        "stad_attention": (C.c_int, [vp, vp, i32, i32, i32, f32, vp]),
        "stad_stat_parts": (C.c_int, [i32, i32]),
        "stad_gemm_bias_residual_stats": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
        "stad_stats_finalize": (C.c_int, [vp, i32, vp, i32, i32, f32, vp]),
        "stad_workspace_bytes": (sz, [C.POINTER(StadDims), i32, i32]),
        "stad_profile_enable": (C.c_int, [i32]),
        "stad_profile_read": (C.c_int, [C.POINTER(StadProfileRecord), i32]),
        "stad_vit_forward": (C.c_int, [C.POINTER(StadModel), C.POINTER(StadInput), vp, i32, i32,
                                       C.POINTER(StadOutputs), vp, sz, vp]),
        "stad_decoder_assemble": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp]),
        "stad_tail_rows_f32": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
        "stad_mae_workspace_bytes": (sz, [C.POINTER(StadMaeModel), i32, i32]),
        "stad_mae_forward": (C.c_int, [C.POINTER(StadMaeModel), C.POINTER(StadInput), vp, vp, i32, i32, vp, vp, sz, vp]),
        "stad_normalize_frames_u8": (C.c_int, [vp, vp, i32, i32, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), i32,
                                               vp]),
        "stad_eval_hist": (C.c_int, [vp, vp, C.c_longlong, vp, i32, vp, vp, vp]),
        "stad_resize_cubic_u8": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
        "stad_rows_norm_head": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, C.c_longlong, C.c_longlong, i32, i32, f32,
                                          vp]),
        "stad_prepend_cls": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, f32, vp]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().stad_last_error().decode("utf-8", "replace")


def check(rc, what):
    """Map a negative STAD_E_* code to the reference's error convention (Python exceptions, mf:188)."""
    if rc >= 0:
        return rc
    msg = f"{what}: {last_error()} (code {rc})"
    if rc in (STAD_E_SHAPE, STAD_E_ALIGN):
        raise ValueError(msg)
    raise RuntimeError(msg)


def init(device=None):
    """stad_init on the tensor's / current device. Raises unless the device is an sm_100 GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("simple-tad_b200 needs a CUDA device (sm_100a); there is no CPU path")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _inited_devices:
        with torch.cuda.device(idx):
            check(load().stad_init(idx), "stad_init")
        _inited_devices.add(idx)
    # the library launches on the CURRENT device and the wrappers pass torch's current stream of that device: a tensor
    # on another GPU of the same process would be read through a stream of the wrong device, so refuse it here
    if device is not None and idx != torch.cuda.current_device():
        raise RuntimeError(f"simple-tad_b200: tensor lives on cuda:{idx} but the current device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device({idx})")
    return idx


def profile_enable(capacity):
    """Bracket every library launch with CUDA events (bench.py's per-kernel roofline); 0 disables."""
    check(load().stad_profile_enable(int(capacity)), "stad_profile_enable")


def profile_read(max_records=1 << 16):
    """[(kind_name, epi, m, n, k, ms)] for every launch since the last read (synchronises on the events)."""
    buf = (StadProfileRecord * max_records)()
    n = check(load().stad_profile_read(buf, max_records), "stad_profile_read")
    return [(KIND_NAMES.get(r.kind, str(r.kind)), r.epi, r.m, r.n, r.k, r.ms) for r in buf[:n]]


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of {dtype}, got {t.dtype} on {t.device} "
                         f"contiguous={t.is_contiguous()}")
    return t


# ------------------------------------------------------------------------------------------------- thin op wrappers
def cast_f32_bf16(x):
    init(x.device)
    _req(x, torch.float32, "x")
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(load().stad_cast_f32_bf16(ptr(x), ptr(y), x.numel(), stream_ptr()), "stad_cast_f32_bf16")
    return y


def row_stats(x, eps):
    init(x.device)
    _req(x, torch.bfloat16, "x")
    M, D = x.shape
    stats = torch.empty(M, 2, dtype=torch.float32, device=x.device)
    check(load().stad_row_stats(ptr(x), ptr(stats), M, D, eps, stream_ptr()), "stad_row_stats")
    return stats


def layernorm(x, g, b, eps):
    init(x.device)
    _req(x, torch.bfloat16, "x")
    M, D = x.shape
    y = torch.empty(M, D, dtype=torch.float32, device=x.device)
    check(load().stad_layernorm(ptr(x), ptr(_req(g, torch.float32, "g")), ptr(_req(b, torch.float32, "b")), ptr(y), M,
                                D, eps, stream_ptr()), "stad_layernorm")
    return y


def pool_norm_head(x, g, b, w_head, b_head, eps, want_probs=False, want_features=False):
    init(x.device)
    _req(x, torch.bfloat16, "x")
    B, N, D = x.shape
    Cn = w_head.shape[0]
    logits = torch.empty(B, Cn, dtype=torch.float32, device=x.device)
    probs = torch.empty(B, Cn, dtype=torch.float32, device=x.device) if want_probs else None
    feats = torch.empty(B, D, dtype=torch.float32, device=x.device) if want_features else None
    scratch = torch.empty(B * 16 * D, dtype=torch.float32, device=x.device)
    check(load().stad_pool_norm_head(ptr(x), ptr(g), ptr(b), ptr(_req(w_head, torch.float32, "w_head")), ptr(b_head),
                                     ptr(logits), ptr(probs), ptr(feats), ptr(scratch), B, N, D, Cn, eps,
                                     stream_ptr()), "stad_pool_norm_head")
    res = (logits,) + ((probs,) if want_probs else ()) + ((feats,) if want_features else ())
    return res if len(res) > 1 else logits


def rows_norm_head(x, g, b, w_head, b_head, eps, rows="all", want_probs=False, want_features=False):
    """x [B, S, D] bf16 -> LayerNorm of token 0 of every clip (rows="cls", mf:327-328) or of every token (rows="all",
    mf:329-330), then the head (mf:334).  Returns logits (, probs) (, features), leading shape [B] or [B, S]."""
    init(x.device)
    _req(x, torch.bfloat16, "x")
    B, S, D = x.shape
    lead = (B,) if rows == "cls" else (B, S)
    R, stride = (B, S) if rows == "cls" else (B * S, 1)
    Cn = w_head.shape[0]
    logits = torch.empty(lead + (Cn,), dtype=torch.float32, device=x.device)
    probs = torch.empty(lead + (Cn,), dtype=torch.float32, device=x.device) if want_probs else None
    feats = torch.empty(lead + (D,), dtype=torch.float32, device=x.device) if want_features else None
    check(load().stad_rows_norm_head(ptr(x), ptr(g), ptr(b), ptr(_req(w_head, torch.float32, "w_head")), ptr(b_head),
                                     ptr(logits), ptr(probs), ptr(feats), R, stride, 0, D, Cn, float(eps), stream_ptr()),
          "stad_rows_norm_head")
    res = (logits,) + ((probs,) if want_probs else ()) + ((feats,) if want_features else ())
    return res if len(res) > 1 else logits


def prepend_cls(emb, cls_token, eps):
    """emb [B, N, D] bf16, cls_token [D] fp32 -> (x [B, N + 1, D] bf16 = cat(cls, emb), stats [B*(N+1), 2] fp32)
    (other_models/MVD/modeling_finetune.py:431-435)."""
    init(emb.device)
    _req(emb, torch.bfloat16, "emb")
    _req(cls_token, torch.float32, "cls_token")
    B, N, D = emb.shape
    if cls_token.numel() != D:
        raise ValueError(f"prepend_cls: cls_token has {cls_token.numel()} values for D={D}")
    x = torch.empty(B, N + 1, D, dtype=torch.bfloat16, device=emb.device)
    stats = torch.empty(B * (N + 1), 2, dtype=torch.float32, device=emb.device)
    check(load().stad_prepend_cls(ptr(emb), ptr(cls_token), ptr(x), ptr(stats), B, N, D, float(eps), stream_ptr()),
          "stad_prepend_cls")
    return x, stats


def ln_gemm(x, stats, w, bias, colsum, gelu=False, out=None):
    init(x.device)
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"ln_gemm: x is [{M},{K}] but w is {tuple(w.shape)}")
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
    check(load().stad_ln_gemm(ptr(x), ptr(_req(stats, torch.float32, "stats")), ptr(w), ptr(bias), ptr(colsum),
                              STAD_EPI_BIAS_GELU if gelu else STAD_EPI_BIAS, ptr(out), M, N, K, stream_ptr()),
          "stad_ln_gemm")
    return out


def gemm_bias_residual(a, w, bias=None, residual=None, out=None):
    init(a.device)
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"gemm: a is [{M},{K}] but w is {tuple(w.shape)}")
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    check(load().stad_gemm_bias_residual(ptr(a), ptr(w), ptr(bias), ptr(residual), ptr(out), M, N, K, stream_ptr()),
          "stad_gemm_bias_residual")
    return out


def gemm_bias_residual_stats(a, w, bias, residual, eps, out=None):
    """out = a w^T + bias + residual, plus the LayerNorm statistics (mean, rstd) [M, 2] of the rows of out, obtained
    from the partial sums the GEMM epilogue emits (no pass over out): returns (out, stats)."""
    init(a.device)
    _req(a, torch.bfloat16, "a")
    _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"gemm: a is [{M},{K}] but w is {tuple(w.shape)}")
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    P = check(load().stad_stat_parts(M, N), "stad_stat_parts")
    parts = torch.empty(P, M, 2, dtype=torch.float32, device=a.device)
    check(load().stad_gemm_bias_residual_stats(ptr(a), ptr(w), ptr(bias), ptr(residual), ptr(out), ptr(parts), M, N, K,
                                               stream_ptr()), "stad_gemm_bias_residual_stats")
    stats = torch.empty(M, 2, dtype=torch.float32, device=a.device)
    check(load().stad_stats_finalize(ptr(parts), P, ptr(stats), M, N, float(eps), stream_ptr()), "stad_stats_finalize")
    return out, stats


def attention(qkv, scale=None, out=None):
    """qkv [B, S, 3, H, 64] bf16 -> [B, S, H*64] bf16 (the tensors of FlashAttention.forward, fac:26-51)."""
    init(qkv.device)
    _req(qkv, torch.bfloat16, "qkv")
    if qkv.dim() != 5 or qkv.shape[2] != 3 or qkv.shape[4] != 64:
        raise ValueError(f"attention: qkv must be [B, S, 3, H, 64], got {tuple(qkv.shape)}")
    B, S, _, H, Dh = qkv.shape
    if scale is None:
        scale = Dh ** -0.5
    if out is None:
        out = torch.empty(B, S, H * Dh, dtype=torch.bfloat16, device=qkv.device)
    check(load().stad_attention(ptr(qkv), ptr(out), B, H, S, float(scale), stream_ptr()), "stad_attention")
    return out


def make_dims(img_h=224, img_w=224, patch=16, tubelet=2, frames=16, in_chans=3, dim=768, depth=12, heads=12,
              hidden=3072, num_classes=2):
    return StadDims(img_h, img_w, patch, tubelet, frames, in_chans, dim, depth, heads, hidden, num_classes)


def make_input(data, mode=STAD_IN_CLIPS, n_frames=0, start=0, stride=1, frame_step=1, window_starts=None,
               tubelet_reuse=False):
    """window_starts: optional int32 CUDA tensor [B] with the first frame of every clip (ABI v6); the caller keeps it
    alive until the launch that reads it has run."""
    ws = None
    if window_starts is not None:
        _req(window_starts, torch.int32, "window_starts")
        ws = window_starts.data_ptr()
    return StadInput(data.data_ptr(), mode, n_frames, start, stride, frame_step, int(bool(tubelet_reuse)), ws)


def patch_embed(x, w, pos_bias, dims, B, n_tok, tok_idx=None, mode=STAD_IN_CLIPS, n_frames=0, start=0, stride=1,
                frame_step=1):
    """x: bf16 planes ([B,C,T,H,W] clips or [F,C,H,W] frames) -> [B*n_tok, D] bf16."""
    init(x.device)
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    _req(pos_bias, torch.float32, "pos_bias")
    D = w.shape[0]
    out = torch.empty(B * n_tok, D, dtype=torch.bfloat16, device=x.device)
    gather = None
    if tok_idx is not None:
        _req(tok_idx, torch.int32, "tok_idx")
        gather = torch.empty(B * n_tok, w.shape[1], dtype=torch.bfloat16, device=x.device)
    inp = make_input(x, mode, n_frames, start, stride, frame_step)
    check(load().stad_patch_embed(C.byref(inp), ptr(w), ptr(pos_bias), ptr(tok_idx), ptr(out), ptr(gather),
                                  C.byref(dims), B, n_tok, stream_ptr()), "stad_patch_embed")
    return out


def decoder_assemble(vis, pos, mask_token, mask_idx, n_tokens, eps):
    """vis [B, n_vis, D] bf16 (x_vis + pos), pos [N, D] fp32, mask_token [D] fp32, mask_idx int32 [B, N - n_vis]
    -> (x_full [B, N, D] bf16, stats [B*N, 2] fp32)   (mp:283-288)."""
    init(vis.device)
    _req(vis, torch.bfloat16, "vis")
    _req(pos, torch.float32, "pos")
    _req(mask_token, torch.float32, "mask_token")
    _req(mask_idx, torch.int32, "mask_idx")
    B, n_vis, D = vis.shape
    N = int(n_tokens)
    if tuple(mask_idx.shape) != (B, N - n_vis) or tuple(pos.shape) != (N, D) or mask_token.numel() != D:
        raise ValueError(f"decoder_assemble: vis {tuple(vis.shape)}, pos {tuple(pos.shape)}, mask_idx {tuple(mask_idx.shape)}")
    x = torch.empty(B, N, D, dtype=torch.bfloat16, device=vis.device)
    stats = torch.empty(B * N, 2, dtype=torch.float32, device=vis.device)
    check(load().stad_decoder_assemble(ptr(vis), ptr(pos), ptr(mask_token), ptr(mask_idx), ptr(x), ptr(stats), B, N,
                                       n_vis, D, float(eps), stream_ptr()), "stad_decoder_assemble")
    return x, stats


def tail_rows_f32(x, n_keep):
    """x [B, N, C] bf16 -> x[:, -n_keep:] as fp32 [B, n_keep, C]   (mp:174)."""
    init(x.device)
    _req(x, torch.bfloat16, "x")
    B, N, Cc = x.shape
    y = torch.empty(B, n_keep, Cc, dtype=torch.float32, device=x.device)
    check(load().stad_tail_rows_f32(ptr(x), ptr(y), B, N, int(n_keep), Cc, stream_ptr()), "stad_tail_rows_f32")
    return y


def normalize_frames_u8(frames, mean, std, bgr=False, out=None):
    """uint8 frames [F, H, W, 3] (HWC, as cv2 delivers them) -> bf16 planes [F, 3, H, W] = (v/255 - mean) / std, RGB
    order   (prepare_image, ri:15-34)."""
    init(frames.device)
    _req(frames, torch.uint8, "frames")
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError(f"normalize_frames_u8: expected [F, H, W, 3] uint8, got {tuple(frames.shape)}")
    F_, H, W, _ = frames.shape
    if out is None:
        out = torch.empty(F_, 3, H, W, dtype=torch.bfloat16, device=frames.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    check(load().stad_normalize_frames_u8(ptr(frames), ptr(out), F_, H, W, m, sd, int(bool(bgr)), stream_ptr()),
          "stad_normalize_frames_u8")
    return out


def eval_hist(probs, labels, thresholds):
    """probs [n, 2] fp32, labels int32 [n], thresholds fp32 [T] ascending -> (hist int64 [2, T+1], conf int64 [4]):
    the exact threshold histogram / arg-max confusion counts of stad_eval_hist (eff:461-488)."""
    init(probs.device)
    _req(probs, torch.float32, "probs")
    _req(labels, torch.int32, "labels")
    _req(thresholds, torch.float32, "thresholds")
    n, T = probs.shape[0], thresholds.numel()
    if probs.dim() != 2 or probs.shape[1] != 2 or labels.numel() != n:
        raise ValueError(f"eval_hist: probs {tuple(probs.shape)}, labels {tuple(labels.shape)}")
    hist = torch.empty(2, T + 1, dtype=torch.int64, device=probs.device)
    conf = torch.empty(4, dtype=torch.int64, device=probs.device)
    check(load().stad_eval_hist(ptr(probs), ptr(labels), n, ptr(thresholds), T, ptr(hist), ptr(conf), stream_ptr()),
          "stad_eval_hist")
    return hist, conf
