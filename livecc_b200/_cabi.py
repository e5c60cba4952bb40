"""ctypes binding of liblivecc_sm100a.so (include/livecc_b200.h).

There is no fallback: if the library is missing or the device is not sm_100, this raises.
Only argument marshalling lives here; tensors are PyTorch CUDA tensors passed as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblivecc_sm100a.so"
HEADER_PATH = PKG_DIR.parent / "include" / "livecc_b200.h"

ABI_VERSION = 8
PAGE_SIZE = 64

# epilogue codes (LCC_EPI_*)
EPI_NONE, EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_BIAS_RESIDUAL, EPI_SWIGLU = range(7)
# stream scalar slots (LCC_SC_*)
SC_KV_LEN, SC_ROPE_POS, SC_FINISHED, SC_N_GENERATED, SC_SEQ_LEN, SC_LAST_TOKEN, SC_VIDEO_TOKENS, SC_NATIVE_ERROR = range(8)
SC_COUNT = 8
MAX_BATCH = 8  # streams per persistent decode launch (MG_MAXB)


class LiveCCNativeError(RuntimeError):
    pass


class Sampling(C.Structure):
    _fields_ = [("repetition_penalty", C.c_float), ("thr_token", C.c_int32), ("thr_base", C.c_float),
                ("thr_step", C.c_float), ("eos_token_id", C.c_int32), ("max_new_tokens", C.c_int32),
                ("inv_repetition_penalty", C.c_float), ("eos_token_id2", C.c_int32)]


class ModelConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vit_depth", "vit_dim", "vit_heads", "vit_mlp", "patch_dim", "merge", "vit_out",
                                         "hidden", "inter", "layers", "q_heads", "kv_heads", "vocab")] + \
               [("rms_eps", C.c_float), ("rope_theta", C.c_float), ("mrope_t", C.c_int32), ("mrope_h", C.c_int32),
                ("video_token_id", C.c_int64)]


class VitBlockW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("norm1_w", "norm1_b", "norm2_w", "norm2_b", "qkv_w", "qkv_b", "proj_w",
                                          "proj_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class LayerW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "qkv_w", "qkv_b", "o_w", "ln2_w", "gate_up_w", "down_w")]


class ModelWeights(C.Structure):
    _fields_ = [("patch_w", C.c_void_p), ("vit_blocks", C.POINTER(VitBlockW)),
                ("merger_ln_w", C.c_void_p), ("merger_ln_b", C.c_void_p), ("merger_fc1_w", C.c_void_p),
                ("merger_fc1_b", C.c_void_p), ("merger_fc2_w", C.c_void_p), ("merger_fc2_b", C.c_void_p),
                ("embed", C.c_void_p), ("layers", C.POINTER(LayerW)), ("final_norm_w", C.c_void_p),
                ("lm_head", C.c_void_p), ("text_inv_freq", C.c_void_p), ("vit_inv_freq", C.c_void_p)]


class StreamState(C.Structure):
    _fields_ = [("k_pool", C.c_void_p), ("v_pool", C.c_void_p), ("layer_stride", C.c_int64),
                ("page_table", C.c_void_p), ("scalars", C.c_void_p), ("seq", C.c_void_p)]


_lib = None


def header_symbols() -> list[str]:
    """Names of every function declared in include/livecc_b200.h (used by the CPU export test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcc_[a-z0-9_]+)\s*\(", text)))


def load_library(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise LiveCCNativeError(f"{LIB_PATH} is missing (run `python -m livecc_b200.build`)")
        from . import build as _build

        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    lib.lcc_abi_version.restype = C.c_int
    if lib.lcc_abi_version() != ABI_VERSION:
        raise LiveCCNativeError("liblivecc_sm100a.so ABI version mismatch; rebuild it (python -m livecc_b200.build --force)")
    lib.lcc_create.restype = C.c_void_p
    lib.lcc_create.argtypes = [C.c_int]
    lib.lcc_destroy.argtypes = [C.c_void_p]
    lib.lcc_last_error.restype = C.c_char_p
    lib.lcc_last_error.argtypes = [C.c_void_p]
    lib.lcc_num_sms.argtypes = [C.c_void_p]
    lib.lcc_launch_count.restype = C.c_uint64
    lib.lcc_launch_count.argtypes = []
    lib.lcc_model_create.restype = C.c_void_p
    lib.lcc_model_create.argtypes = [C.c_void_p, C.POINTER(ModelConfig), C.POINTER(ModelWeights)]
    lib.lcc_model_destroy.argtypes = [C.c_void_p]
    lib.lcc_workspace_bytes.restype = C.c_size_t
    lib.lcc_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.lcc_ws_offset.restype = C.c_size_t
    lib.lcc_ws_offset.argtypes = [C.c_void_p, C.c_int]
    lib.lcc_resize_aa_taps.argtypes = [C.c_int, C.c_int]
    lib.lcc_resize_aa_table.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lcc_resize_plan_create.restype = C.c_void_p
    lib.lcc_resize_plan_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lcc_resize_plan_destroy.argtypes = [C.c_void_p]
    lib.lcc_resize_plan_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def launch_count() -> int:
    """Kernel launches issued by the library so far (graph-captured launches count once, at capture)."""
    return int(load_library().lcc_launch_count())


def resize_aa_table(in_size: int, out_size: int):
    """Host-only: (xmin int32[out], xsize int32[out], weights float32[taps, out]) of one axis of the antialiased bicubic
    resize (lcc_resize_aa_table; ATen _compute_indices_min_size_weights_aa<float>). Needs no GPU."""
    import numpy as np

    lib = load_library()
    taps = lib.lcc_resize_aa_taps(int(in_size), int(out_size))
    if taps <= 0:
        raise LiveCCNativeError(f"lcc_resize_aa_taps({in_size}, {out_size}) failed")
    xmin = np.zeros(out_size, np.int32)
    xsize = np.zeros(out_size, np.int32)
    w = np.zeros((taps, out_size), np.float32)
    rc = lib.lcc_resize_aa_table(int(in_size), int(out_size), xmin.ctypes.data_as(C.c_void_p),
                                 xsize.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    if rc:
        raise LiveCCNativeError(f"lcc_resize_aa_table({in_size}, {out_size}) failed ({rc})")
    return xmin, xsize, w


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _i(x):
    return C.c_int(int(x))


class Context:
    """One lcc_ctx bound to one CUDA device."""

    def __init__(self, device_index: int):
        self.lib = load_library()
        self.handle = self.lib.lcc_create(int(device_index))
        if not self.handle:
            raise LiveCCNativeError(
                f"lcc_create({device_index}) failed: an sm_100 (B200) device is required; no fallback path exists")
        self.device_index = device_index
        self.num_sms = self.lib.lcc_num_sms(C.c_void_p(self.handle))

    def close(self):
        for plan in getattr(self, "_resize_plans", {}).values():
            self.lib.lcc_resize_plan_destroy(C.c_void_p(plan))
        self._resize_plans = {}
        if getattr(self, "handle", None):
            self.lib.lcc_destroy(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.lcc_last_error(C.c_void_p(self.handle)).decode()
            raise LiveCCNativeError(f"{what} failed ({rc}): {msg}")

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(C.c_void_p(self.handle), *args)
        self.check(rc, name)

    @staticmethod
    def stream_ptr():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- op level ----------------------------------------------------------------------------
    def gemm(self, a, b, out=None, bias=None, residual=None, epilogue=EPI_NONE, block_n=0, splitk_ws=None):
        """out[M,N] = epilogue(a[M,K] @ b[N,K]^T); 2-D bf16 CUDA tensors, last dim contiguous.
        splitk_ws: optional fp32 scratch tensor (experimental split-K, LIVECC_B200_GEMM_SPLITK=1)."""
        import torch

        M, K = a.shape
        N = b.shape[0]
        assert b.shape[1] == K and a.stride(1) == 1 and b.stride(1) == 1
        n_out = N // 2 if epilogue == EPI_SWIGLU else N
        if out is None:
            out = torch.empty((M, n_out), dtype=torch.bfloat16, device=a.device)
        assert out.stride(1) == 1 and out.shape == (M, n_out)
        ldr = residual.stride(0) if residual is not None else 0
        self.call("lcc_gemm_bf16", _ptr(a), _i(a.stride(0)), _ptr(b), _i(b.stride(0)), _ptr(out), _i(out.stride(0)),
                  _i(M), _i(N), _i(K), _ptr(bias), _ptr(residual), _i(ldr), _i(epilogue), _i(block_n), _ptr(splitk_ws),
                  C.c_int64(splitk_ws.numel() * splitk_ws.element_size() if splitk_ws is not None else 0),
                  self.stream_ptr())
        return out

    def cast_f32_bf16(self, x):
        import torch

        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        self.call("lcc_cast_f32_bf16", _ptr(x), _ptr(out), C.c_int64(x.numel()), self.stream_ptr())
        return out

    def resize_plan(self, h, w, H, W, rows_per_cta=0):
        """Device tables for (h, w) -> (H, W), created once per size and owned by this context."""
        plans = self.__dict__.setdefault("_resize_plans", {})
        key = (int(h), int(w), int(H), int(W), int(rows_per_cta))
        if key not in plans:
            plan = self.lib.lcc_resize_plan_create(C.c_void_p(self.handle), *map(int, key))
            if not plan:
                raise LiveCCNativeError("lcc_resize_plan_create failed: " + self.lib.lcc_last_error(C.c_void_p(self.handle)).decode())
            plans[key] = plan
        return plans[key]

    def resize_plan_info(self, plan):
        th, rows, smem = C.c_int(0), C.c_int(0), C.c_int64(0)
        self.lib.lcc_resize_plan_info(C.c_void_p(plan), C.byref(th), C.byref(rows), C.byref(smem))
        return {"rows_per_cta": th.value, "max_rows": rows.value, "smem_bytes": smem.value}

    def resize_bicubic_aa_u8(self, clip, size, out=None, rows_per_cta=0):
        """torchvision F.resize(clip, size, BICUBIC, antialias=True) of a contiguous uint8 CUDA tensor [..., h, w]
        (video_process_patch.py:101-106,150-155), bit-identical, in one kernel."""
        import torch

        if clip.dtype != torch.uint8 or not clip.is_cuda or not clip.is_contiguous() or clip.dim() < 2:
            raise ValueError("resize_bicubic_aa_u8 takes a contiguous uint8 CUDA tensor [..., h, w]")
        h, w = clip.shape[-2:]
        H, W = int(size[0]), int(size[1])
        planes = clip.numel() // (h * w)
        if out is None:
            out = torch.empty(clip.shape[:-2] + (H, W), dtype=torch.uint8, device=clip.device)
        plan = self.resize_plan(h, w, H, W, rows_per_cta)
        self.call("lcc_resize_bicubic_aa_u8", C.c_void_p(plan), _ptr(clip), _i(planes), _ptr(out), self.stream_ptr())
        return out

    def layernorm(self, x, w, b, eps=1e-6):
        import torch

        y = torch.empty_like(x)
        self.call("lcc_layernorm", _ptr(x), _i(x.stride(0)), _ptr(w), _ptr(b), _ptr(y), _i(y.stride(0)),
                  _i(x.shape[0]), _i(x.shape[1]), C.c_float(eps), self.stream_ptr())
        return y

    def rmsnorm(self, x, w, eps=1e-6):
        import torch

        y = torch.empty_like(x)
        self.call("lcc_rmsnorm", _ptr(x), _i(x.stride(0)), _ptr(w), _ptr(y), _i(y.stride(0)), _i(x.shape[0]),
                  _i(x.shape[1]), C.c_float(eps), self.stream_ptr())
        return y

    def vit_rope_table(self, t, h, w, head_dim, inv_freq, merge=2):
        import torch

        n = t * h * w
        cos = torch.empty((n, head_dim // 2), dtype=torch.float32, device=inv_freq.device)
        sin = torch.empty_like(cos)
        self.call("lcc_vit_rope_table", _ptr(cos), _ptr(sin), _i(t), _i(h), _i(w), _i(merge), _i(head_dim),
                  _ptr(inv_freq), self.stream_ptr())
        return cos, sin

    def vit_rope_apply(self, qkv, cos, sin, heads, head_dim):
        self.call("lcc_vit_rope_apply", _ptr(qkv), _i(qkv.stride(0)), _ptr(cos), _ptr(sin), _i(qkv.shape[0]),
                  _i(heads), _i(head_dim), self.stream_ptr())
        return qkv

    def vit_attention(self, qkv, cu_seqlens, max_seg_len, heads, head_dim, impl=0):
        """impl: 0 default, 1 mma.sync kernels, 2 tcgen05 kernel (LCC_VIT_ATTN_*)"""
        import torch

        out = torch.empty((qkv.shape[0], heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
        self.call("lcc_vit_attention", _ptr(qkv), _i(qkv.stride(0)), C.c_int64(qkv.shape[0]), _ptr(out),
                  _i(out.stride(0)), _ptr(cu_seqlens), _i(cu_seqlens.numel() - 1), _i(max_seg_len), _i(heads),
                  _i(head_dim), _i(impl), self.stream_ptr())
        return out

    def embed_gather(self, ids, table, video_embeds, video_token_id):
        import torch

        S, H = ids.numel(), table.shape[1]
        out = torch.empty((S, H), dtype=torch.bfloat16, device=table.device)
        rank = torch.empty((S + 2,), dtype=torch.int32, device=table.device)
        n_rows = video_embeds.shape[0] if video_embeds is not None else 0
        self.call("lcc_embed_gather", _ptr(ids), _ptr(table), _ptr(video_embeds), _i(n_rows), C.c_int64(video_token_id), _ptr(out),
                  _ptr(rank), _i(S), _i(H), C.c_int64(table.shape[0]), self.stream_ptr())
        return out, rank

    def mrope_kv_write(self, qkv, pos3, inv_freq, sec_t, sec_h, Hq, Hkv, k_cache, v_cache, page_table, kv_start):
        self.call("lcc_mrope_kv_write", _ptr(qkv), _i(qkv.stride(0)), _ptr(pos3), _i(qkv.shape[0]), _ptr(inv_freq),
                  _i(sec_t), _i(sec_h), _i(Hq), _i(Hkv), _ptr(k_cache), _ptr(v_cache), _ptr(page_table), _i(kv_start),
                  self.stream_ptr())

    def attn_prefill(self, q, k_cache, v_cache, page_table, Hq, Hkv, past, impl=0, split=False):
        """impl: 0 default, 1 mma.sync kernel, 2 tcgen05 kernel (LCC_ATTN_*); split: provide split-KV scratch"""
        import torch

        S = q.shape[0]
        out = torch.empty((S, Hq * 128), dtype=torch.bfloat16, device=q.device)
        part_rows = 8 * S * Hq if split else 0
        part_o = torch.empty((max(part_rows, 1), 128), dtype=torch.float32, device=q.device)
        part_ml = torch.empty((max(part_rows, 1), 2), dtype=torch.float32, device=q.device)
        self.call("lcc_attn_prefill", _ptr(q), _i(q.stride(0)), _ptr(k_cache), _ptr(v_cache), _ptr(page_table), _i(Hq),
                  _i(Hkv), _i(S), _i(past), _ptr(out), _i(out.stride(0)), _ptr(part_o) if split else None,
                  _ptr(part_ml) if split else None, C.c_int64(part_rows), _i(impl), self.stream_ptr())
        return out

    def attn_decode(self, qkv, k_cache, v_cache, page_table, scalars, inv_freq, Hq, Hkv, nsplit):
        import torch

        dev = qkv.device
        part_o = torch.empty((nsplit, Hq, 128), dtype=torch.float32, device=dev)
        part_ml = torch.empty((nsplit, Hq, 2), dtype=torch.float32, device=dev)
        out = torch.empty((Hq * 128,), dtype=torch.bfloat16, device=dev)
        if getattr(self, "_attn_counters", None) is None:
            self._attn_counters = torch.zeros(64, dtype=torch.int32, device=dev)
        self.call("lcc_attn_decode", _ptr(qkv), _ptr(k_cache), _ptr(v_cache), _ptr(page_table), _ptr(scalars),
                  _ptr(inv_freq), _i(Hq), _i(Hkv), _i(nsplit), _ptr(part_o), _ptr(part_ml), _ptr(self._attn_counters),
                  _ptr(out), self.stream_ptr())
        return out

    def gemv_norm_bias(self, W, x, norm_w, eps, bias, scalars=None):
        import torch

        out = torch.empty((W.shape[0],), dtype=torch.bfloat16, device=W.device)
        self.call("lcc_gemv_norm_bias", _ptr(W), _i(W.stride(0)), _ptr(x), _ptr(norm_w), C.c_float(eps), _ptr(bias),
                  _ptr(out), _i(W.shape[0]), _i(W.shape[1]), _ptr(scalars), self.stream_ptr())
        return out

    def gemv_residual(self, W, x, h_inout, scalars=None):
        self.call("lcc_gemv_residual", _ptr(W), _i(W.stride(0)), _ptr(x), _ptr(h_inout), _i(W.shape[0]), _i(W.shape[1]),
                  _ptr(scalars), self.stream_ptr())
        return h_inout

    def gemv_norm_swiglu(self, W_gu, x, norm_w, eps, scalars=None):
        import torch

        act = torch.empty((W_gu.shape[0] // 2,), dtype=torch.bfloat16, device=W_gu.device)
        self.call("lcc_gemv_norm_swiglu", _ptr(W_gu), _i(W_gu.stride(0)), _ptr(x), _ptr(norm_w), C.c_float(eps),
                  _ptr(act), _i(W_gu.shape[0]), _i(W_gu.shape[1]), _ptr(scalars), self.stream_ptr())
        return act

    def gemv_norm_logits(self, W, x, norm_w, eps, scalars=None):
        import torch

        lg = torch.empty((W.shape[0],), dtype=torch.float32, device=W.device)
        lg2 = torch.empty_like(lg)
        self.call("lcc_gemv_norm_logits", _ptr(W), _i(W.stride(0)), _ptr(x), _ptr(norm_w), C.c_float(eps), _ptr(lg),
                  _ptr(lg2), _i(W.shape[0]), _i(W.shape[1]), _ptr(scalars), self.stream_ptr())
        return lg, lg2

    def sample_greedy(self, logits_raw, logits_proc, seq, scalars, sampling: Sampling, advance_kv, embed, h):
        self.call("lcc_sample_greedy", _ptr(logits_raw), _ptr(logits_proc), _i(logits_raw.numel()), _ptr(seq),
                  _ptr(scalars), C.byref(sampling), _i(advance_kv), _ptr(embed), _ptr(h), _i(embed.shape[1]),
                  self.stream_ptr())


class NativeModel:
    """lcc_model: weight pointers + config + bound workspace; phase-level launches."""

    def __init__(self, ctx: Context, cfg: ModelConfig, weights: ModelWeights, keepalive):
        self.ctx = ctx
        self.lib = ctx.lib
        self._keep = keepalive  # python objects that own the memory the C structs point to
        self.handle = self.lib.lcc_model_create(C.c_void_p(ctx.handle), C.byref(cfg), C.byref(weights))
        if not self.handle:
            raise LiveCCNativeError("lcc_model_create failed: " + self.lib.lcc_last_error(C.c_void_p(ctx.handle)).decode())
        self.workspace = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lcc_model_destroy(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(C.c_void_p(self.handle), *args)
        self.ctx.check(rc, name)

    def bind_workspace(self, max_patches: int, max_tokens: int, device):
        import torch

        nbytes = self.lib.lcc_workspace_bytes(C.c_void_p(self.handle), int(max_patches), int(max_tokens))
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        self.cap_patches, self.cap_tokens = int(max_patches), int(max_tokens)
        self._call("lcc_model_bind_workspace", _ptr(self.workspace), C.c_size_t(nbytes), _i(max_patches), _i(max_tokens))
        self.hidden_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 0)
        self.logits_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 1)
        self.decode_hidden_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 2)
        self.decode_qkv_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 3)
        self.decode_attn_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 4)
        self.decode_act_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 5)
        self.logits_proc_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 6)
        self.mega_error_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 7)
        self.mega_trace_offset = self.lib.lcc_ws_offset(C.c_void_p(self.handle), 8)

    def mega_error(self) -> int:
        """Sticky error flag of the persistent decode kernel (synchronises the device)."""
        import torch

        off = self.mega_error_offset
        return int(self.workspace[off:off + 4].view(torch.int32).item())

    def vit_forward(self, pixel_values, t, h, w, out):
        self._call("lcc_vit_forward", _ptr(pixel_values), _i(t), _i(h), _i(w), _ptr(out), Context.stream_ptr())

    def vit_forward_frames(self, frames_u8, mean255, std255, out):
        T, _, H, W = frames_u8.shape
        m = (C.c_float * 3)(*mean255)
        sd = (C.c_float * 3)(*std255)
        self._call("lcc_vit_forward_frames", _ptr(frames_u8), _i(T), _i(H), _i(W), m, sd, _ptr(out), Context.stream_ptr())

    def prefill(self, st: StreamState, ids, pos3, S, past, video_embeds, sampling: Sampling, slot: int = 0):
        n_rows = video_embeds.shape[0] if video_embeds is not None else 0
        self._call("lcc_prefill", C.byref(st), _ptr(ids), _ptr(pos3), _i(S), _i(past), _ptr(video_embeds), _i(n_rows),
                   C.byref(sampling), _i(slot), Context.stream_ptr())

    def decode_steps(self, st: StreamState, n_steps, nsplit, sampling: Sampling):
        self._call("lcc_decode_steps", C.byref(st), _i(n_steps), _i(nsplit), C.byref(sampling), Context.stream_ptr())

    def decode_batch(self, states, n_steps, sampling: Sampling):
        """states: list of StreamState (<= MAX_BATCH), stream b prefilled with slot=b."""
        arr = (StreamState * len(states))(*states)
        self._call("lcc_decode_batch", arr, _i(len(states)), _i(n_steps), C.byref(sampling), Context.stream_ptr())

    def decode_mega_debug(self, states, layer_begin, layer_end, phase_mask, do_head):
        arr = (StreamState * len(states))(*states)
        self._call("lcc_decode_mega_debug", arr, _i(len(states)), _i(layer_begin), _i(layer_end), _i(phase_mask),
                   _i(do_head), Context.stream_ptr())
