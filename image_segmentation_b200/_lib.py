"""ctypes binding of libunetk.so (the C ABI declared in include/unetk.h).

Host-side plumbing only: device memory comes from PyTorch's allocator and is passed as raw pointers,
kernels are enqueued on torch's current CUDA stream.  There is no CPU path and no fallback: if the
shared library is missing or the tensors are not CUDA tensors the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libunetk.so")

F32, BF16 = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0, 1, 2
# experiment switches of the tcgen05 path (include/unetk.h UNETK_TC_*), OR-ed into `algo`; the shared library itself
# reads no environment variables -- tools set TC_FLAGS (or UNETK_TC_FLAGS="no_pair,no_halo,...") on the Python side
TC_DETERMINISTIC = 1 << 14      # unetk_wgrad: ordered two-pass split reduction instead of fp32 atomics
TC_FLAG_BITS = {"no_pair": 1 << 8, "no_halo": 1 << 9, "no_even_groups": 1 << 10, "no_halo_n256": 1 << 11,
                "no_halo_pair": 1 << 12, "no_wgrad_c64": 1 << 13}
TC_FLAGS = 0
for _f in filter(None, os.environ.get("UNETK_TC_FLAGS", "").split(",")):
    TC_FLAGS |= TC_FLAG_BITS[_f.strip()]
MODE_1X1, MODE_3X3, MODE_CONVT, MODE_CONVT_GATHER = 0, 1, 2, 3

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}


class Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("ld", C.c_int32), ("dtype", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("x", Tensor), ("w", C.c_void_p), ("y", Tensor), ("mode", C.c_int32), ("algo", C.c_int32),
                ("bias", C.c_void_p), ("stat_sum", C.c_void_p), ("stat_sumsq", C.c_void_p),
                ("bn_z", Tensor), ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("bn_mean", C.c_void_p),
                ("bn_invstd", C.c_void_p), ("bn_sums", C.c_void_p)]


class WgradArgs(C.Structure):
    _fields_ = [("u", Tensor), ("s", Tensor), ("dw", C.c_void_p), ("mode", C.c_int32), ("algo", C.c_int32),
                ("partial", C.c_void_p), ("partial_bytes", C.c_int64)]


class BnFinalizeArgs(C.Structure):
    _fields_ = [("sum", C.c_void_p), ("sumsq", C.c_void_p), ("count", C.c_int64), ("c", C.c_int32),
                ("training", C.c_int32), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("conv_bias", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("momentum", C.c_float), ("eps", C.c_float), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p)]


class BnBwdArgs(C.Structure):
    _fields_ = [("z", Tensor), ("dy", Tensor), ("dpool", Tensor), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p), ("sums", C.c_void_p), ("dz", Tensor),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("pool_idx", C.c_void_p)]


class WJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst0", C.c_void_p), ("dst1", C.c_void_p), ("kind", C.c_int32),
                ("cout", C.c_int32), ("cin", C.c_int32), ("kpad", C.c_int32)]


class DiceCeArgs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("target", C.c_void_p), ("n", C.c_int32), ("c", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("class_weights", C.c_void_p), ("has_ignore", C.c_int32), ("ignore_index", C.c_int64),
                ("dice_weight", C.c_float), ("ce_weight", C.c_float), ("smooth", C.c_float), ("accum", C.c_void_p),
                ("coef", C.c_void_p), ("loss", C.c_void_p), ("status", C.c_void_p), ("grad_out", C.c_void_p),
                ("dlogits", C.c_void_p), ("input_kind", C.c_int32), ("nll_eps", C.c_float)]


LOSS_LOGITS, LOSS_PROBS_LOG, LOSS_PROBS_RAW = 0, 1, 2


UNETK_U8, UNETK_I64 = 2, 3      # label dtypes (include/unetk.h)


class HeadBnBwdArgs(C.Structure):
    _fields_ = [("z", Tensor), ("dlogits", C.c_void_p), ("w_head", C.c_void_p), ("dout", C.c_int32),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("sums", C.c_void_p), ("dz", Tensor), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
                ("dw_head", C.c_void_p), ("db_head", C.c_void_p)]


class EvalImage(C.Structure):
    _fields_ = [("crop_top", C.c_int32), ("crop_left", C.c_int32), ("crop_h", C.c_int32), ("crop_w", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32), ("offset", C.c_int64)]


class EvalArgs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("n", C.c_int32), ("c", C.c_int32), ("th", C.c_int32), ("tw", C.c_int32),
                ("images", C.c_void_p), ("max_out_pixels", C.c_int32), ("labels", C.c_void_p),
                ("label_dtype", C.c_int32), ("class_weights", C.c_void_p), ("has_ignore", C.c_int32),
                ("ignore_index", C.c_int64), ("dice_weight", C.c_float), ("ce_weight", C.c_float),
                ("smooth", C.c_float), ("accum", C.c_void_p), ("loss_per_image", C.c_void_p),
                ("loss_sum", C.c_void_p), ("counts", C.c_void_p), ("status", C.c_void_p)]


_lib = None


def lib():
    """The loaded shared library; raises loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -m image_segmentation_b200._build, or __graft_entry__.build()). There is no fallback path.")
        l = C.CDLL(LIB_PATH)
        l.unetk_version.restype = C.c_int
        l.unetk_query_workspace.restype = C.c_int64
        l.unetk_query_workspace.argtypes = [C.c_int32] * 5
        l.unetk_wgrad_partial_bytes.restype = C.c_int64
        l.unetk_wgrad_partial_bytes.argtypes = [C.POINTER(WgradArgs)]
        l.unetk_last_error.restype = C.c_char_p
        vp = C.c_void_p
        P = C.POINTER
        sigs = {
            "unetk_device_query": [P(C.c_int32), P(C.c_int32), P(C.c_int32)],
            "unetk_im2col3x3_first": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P(Tensor), vp],
            "unetk_permute3": [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [C.c_int64] * 6 + [vp],
            "unetk_weights_pack": [vp, vp, C.c_int32, C.c_int32, vp],
            "unetk_weights_unpack": [vp, vp, C.c_int32, vp, vp],
            "unetk_conv": [P(ConvArgs), vp],
            "unetk_wgrad": [P(WgradArgs), vp],
            "unetk_channel_sum": [P(Tensor), vp, vp],
            "unetk_channel_sum_ordered": [P(Tensor), vp, vp, C.c_int64, vp],
            "unetk_bn_stats": [P(Tensor), vp, vp, vp],
            "unetk_bn_finalize": [P(BnFinalizeArgs), vp],
            "unetk_bn_relu_apply": [P(Tensor), vp, vp, P(Tensor), P(Tensor), vp, vp],
            "unetk_bn_relu_bwd_reduce": [P(BnBwdArgs), vp],
            "unetk_bn_relu_bwd_apply": [P(BnBwdArgs), vp],
            "unetk_head_fprop": [P(Tensor), vp, vp, C.c_int32, vp, vp],
            "unetk_head_bwd": [vp, P(Tensor), vp, C.c_int32, P(Tensor), vp, vp, vp],
            "unetk_dice_ce_fwd": [P(DiceCeArgs), vp],
            "unetk_dice_ce_bwd": [P(DiceCeArgs), vp],
            "unetk_argmax_confusion": [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp],
            "unetk_bn_relu_head_fprop": [P(Tensor), vp, vp, P(Tensor), vp, vp, C.c_int32, vp, vp],
            "unetk_head_bn_bwd_reduce": [P(HeadBnBwdArgs), vp],
            "unetk_head_bn_bwd_apply": [P(HeadBnBwdArgs), vp],
            "unetk_crop_resize": [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, vp, vp],
            "unetk_eval_loss_metrics": [P(EvalArgs), vp],
            "unetk_bias_sigmoid_fwd": [P(Tensor), vp, C.c_int32, vp, vp],
            "unetk_bias_sigmoid_bwd": [vp, vp, C.c_int32, P(Tensor), vp, vp],
            "unetk_bilinear_up_fwd": [P(Tensor), P(Tensor), vp],
            "unetk_bilinear_up_bwd": [P(Tensor), P(Tensor), vp],
            "unetk_prompt_compose_fwd": [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp],
            "unetk_prompt_compose_bwd": [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp],
            "unetk_nvls_allreduce_f32": [vp, C.c_int64, C.c_int32, C.c_int32, C.c_float, vp],
            "unetk_nchw_to_nhwc": [vp, P(Tensor), vp],
            "unetk_nhwc_to_nchw": [P(Tensor), vp, vp],
        }
        for name, argtypes in sigs.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = l
    return _lib


EXPORTED_SYMBOLS = (
    "unetk_version", "unetk_last_error", "unetk_device_query", "unetk_query_workspace", "unetk_struct_size", "unetk_im2col3x3_first", "unetk_permute3",
    "unetk_weights_pack", "unetk_weights_unpack",
    "unetk_conv", "unetk_wgrad", "unetk_wgrad_partial_bytes", "unetk_channel_sum", "unetk_channel_sum_ordered", "unetk_bn_stats", "unetk_bn_finalize", "unetk_bn_relu_apply",
    "unetk_bn_relu_bwd_reduce", "unetk_bn_relu_bwd_apply", "unetk_head_fprop", "unetk_head_bwd",
    "unetk_dice_ce_fwd", "unetk_dice_ce_bwd", "unetk_argmax_confusion", "unetk_crop_resize", "unetk_eval_loss_metrics",
    "unetk_head_bn_bwd_reduce", "unetk_head_bn_bwd_apply", "unetk_bn_relu_head_fprop",
    "unetk_bias_sigmoid_fwd", "unetk_bias_sigmoid_bwd", "unetk_bilinear_up_fwd", "unetk_bilinear_up_bwd",
    "unetk_prompt_compose_fwd", "unetk_prompt_compose_bwd", "unetk_nchw_to_nhwc", "unetk_nhwc_to_nchw",
    "unetk_nvls_allreduce_f32",
)


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libunetk: " + lib().unetk_last_error().decode(errors="replace"))


WS_BN_STATS, WS_BN_BWD_SUMS, WS_HEAD_BN_SUMS, WS_POOL_IDX, WS_WGRAD, WS_DICE_ACCUM, WS_DICE_COEF, WS_EVAL_ACCUM, \
    WS_CONFUSION = range(9)


def query_workspace(what: int, a: int, b: int = 0, c: int = 0, d: int = 0) -> int:
    """Bytes of caller-provided scratch `what` (include/unetk.h UNETK_WS_*)."""
    n = lib().unetk_query_workspace(what, a, b, c, d)
    if n < 0:
        check(int(n))
    return int(n)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("image_segmentation_b200 runs on CUDA tensors only (no CPU fallback); got device "
                               + str(t.device))


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def nhwc(t: Optional[torch.Tensor]) -> Tensor:
    """unetk_tensor for a [N,H,W,C] torch view whose pixels are `ld` elements apart (channel slices allowed)."""
    if t is None:
        return Tensor(None, 0, 0, 0, 0, 0, 0)
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    if c > 1 and sc != 1:
        raise ValueError("channels must be contiguous")
    ld = sw if w > 1 else (sh // w if h > 1 else (sn // (h * w) if n > 1 else c))
    if (w > 1 and sw != ld) or (h > 1 and sh != ld * w) or (n > 1 and sn != ld * w * h):
        raise ValueError(f"not a pixel-major NHWC view: shape {tuple(t.shape)} stride {t.stride()}")
    return Tensor(t.data_ptr(), n, h, w, c, ld, _DTYPES[t.dtype])


# ---- instrumentation (bench.py): launch counter and optional per-call CUDA-event records ----------
COUNTERS = {"launches": 0}
PROFILE_HOOK = None   # set to a list to collect (kind, algorithmic_flops, start_event, end_event, label, tag)
LABEL = ""            # free-form tag of the layer being launched (set by the engine, read by the hook)


class Call:
    """One prepared C-ABI call: the ctypes argument structs are built ONCE (device pointers inside a plan are stable)
    and the call is then replayed every step with only the stream changing -- the host cost of a launch drops from a
    struct build + stride checks to one foreign-function call.  `keep` pins the tensors the structs point into.
    Pointer fields that change per step (gradients living in a per-backward flat buffer) are re-pointed with ``patch``."""
    __slots__ = ("kind", "nlaunch", "flops", "fn", "args", "tag", "label", "keep", "patches")

    def __init__(self, kind, nlaunch, flops, fn, args, tag=None, keep=None, label=None):
        self.kind, self.nlaunch, self.flops, self.fn, self.args = kind, nlaunch, flops or 0, fn, tuple(args)
        self.tag, self.keep, self.label, self.patches = tag, keep, label, None

    def patch_ptr(self, struct, field: str, byte_offset: int):
        """Before every call, ``struct.field = base_ptr + byte_offset`` (base_ptr is passed to __call__)."""
        if self.patches is None:
            self.patches = []
        self.patches.append((struct, field, byte_offset))
        return self

    def __call__(self, stream: int, base_ptr: int = 0):
        COUNTERS["launches"] += self.nlaunch
        if self.patches:
            for struct, field, off in self.patches:
                setattr(struct, field, base_ptr + off)
        hook = PROFILE_HOOK
        if hook is None:
            rc = self.fn(*self.args, stream)
            if rc:
                check(rc)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = self.fn(*self.args, stream)
        e1.record()
        if rc:
            check(rc)
        hook.append((self.kind, self.flops, e0, e1, self.label if self.label is not None else LABEL, self.tag))


def _run(kind, nlaunch, flops, fn, *args, tag=None):
    """Immediate call (tests, tools): `args` includes the stream as its last element."""
    Call(kind, nlaunch, flops, fn, args[:-1], tag=tag)(args[-1])


# ---- thin wrappers: prep_*() builds a Call, the plain name runs it on the current stream ---------------------------
def im2col3x3_first(x_nchw: torch.Tensor, out: torch.Tensor):
    prep_im2col3x3_first(x_nchw, out)(stream_ptr())


def prep_im2col3x3_first(x_nchw: torch.Tensor, out: torch.Tensor) -> Call:
    n, cin, h, w = x_nchw.shape
    t = nhwc(out)
    return Call("layout", 1, 0, lib().unetk_im2col3x3_first, (x_nchw.data_ptr(), n, cin, h, w, C.byref(t)), keep=(t, x_nchw, out))


def permute3(src: torch.Tensor, dst: torch.Tensor, dims, src_strides, dst_strides, dst_offset_elems: int = 0):
    _run("layout", 1, 0, lib().unetk_permute3, src.data_ptr(), dst.data_ptr() + dst_offset_elems * dst.element_size(),
         _DTYPES[dst.dtype], dims[0], dims[1], dims[2], src_strides[0], src_strides[1], src_strides[2],
         dst_strides[0], dst_strides[1], dst_strides[2], stream_ptr())


class WeightJobs:
    """Device-resident job + tile tables for unetk_weights_pack / unetk_weights_unpack."""

    def __init__(self, jobs, device):
        # jobs: list of (src_ptr_or_offset, dst0, dst1, kind, cout, cin, kpad)
        arr = (WJob * len(jobs))(*[WJob(*j) for j in jobs])
        tiles = []
        for ji, (_, _, _, kind, cout, cin, _) in enumerate(jobs):
            if kind in (2, 3):
                tiles.append((ji, 0, 0, 0))
                continue
            na, nb = (cout, cin) if kind == 0 else (cin, cout)
            if na % 32 or nb % 32:
                raise ValueError("channel counts must be multiples of 32 for the batched weight kernels")
            tiles += [(ji, a0, b0, 0) for a0 in range(0, na, 32) for b0 in range(0, nb, 32)]
        self.ntiles = len(tiles)
        self.jobs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        self.tiles = torch.tensor(tiles, dtype=torch.int32).to(device)


def weights_pack(wj: "WeightJobs", dtype):
    _run("layout", 1, 0, lib().unetk_weights_pack, wj.jobs.data_ptr(), wj.tiles.data_ptr(), wj.ntiles, _DTYPES[dtype], stream_ptr())


def weights_unpack(wj: "WeightJobs", dst_base: torch.Tensor):
    _run("layout", 1, 0, lib().unetk_weights_unpack, wj.jobs.data_ptr(), wj.tiles.data_ptr(), wj.ntiles, dst_base.data_ptr(),
         stream_ptr())


def conv_flops(x, y, mode):
    if mode in (MODE_1X1, MODE_3X3):
        return 2 * y.shape[0] * y.shape[1] * y.shape[2] * (9 if mode == MODE_3X3 else 1) * x.shape[3] * y.shape[3]
    if mode == MODE_CONVT:
        return 2 * x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] * 4 * y.shape[3]
    return 2 * y.shape[0] * y.shape[1] * y.shape[2] * 4 * x.shape[3] * y.shape[3]


def prep_conv(x, w, y, mode, bias=None, stat_sum=None, stat_sumsq=None, algo=ALGO_AUTO, algo_flops=None, bn_reduce=None,
              label=None) -> Call:
    """bn_reduce = (z, scale, shift, mean, invstd, sums): fuse the BatchNorm-backward reduction of the layer whose
    activated-output gradient this launch produces (see unetk.h)."""
    if bn_reduce is None:
        bnz, bsc, bsh, bmu, bis, bsum = nhwc(None), None, None, None, None, None
    else:
        bnz = nhwc(bn_reduce[0])
        bsc, bsh, bmu, bis, bsum = (t.data_ptr() for t in bn_reduce[1:])
    a = ConvArgs(nhwc(x), w.data_ptr(), nhwc(y), mode, algo | TC_FLAGS, ptr(bias), ptr(stat_sum), ptr(stat_sumsq),
                 bnz, bsc, bsh, bmu, bis, bsum)
    if algo_flops is None:
        algo_flops = conv_flops(x, y, mode)
    simt = (algo & 0xff) == ALGO_SIMT or ((algo & 0xff) == ALGO_AUTO and (x.dtype != torch.bfloat16 or x.shape[3] % 64
                                                                            or y.shape[3] % 64))
    return Call("conv", 2 if (simt and stat_sum is not None) else 1, algo_flops, lib().unetk_conv, (C.byref(a),), tag=mode,
                keep=(a, x, w, y, bias, stat_sum, stat_sumsq, bn_reduce), label=label)


def conv(x, w, y, mode, bias=None, stat_sum=None, stat_sumsq=None, algo=ALGO_AUTO, algo_flops=None, bn_reduce=None):
    prep_conv(x, w, y, mode, bias, stat_sum, stat_sumsq, algo, algo_flops, bn_reduce)(stream_ptr())


def wgrad_partial_bytes(u, s, mode, algo=ALGO_AUTO) -> int:
    """Scratch bytes unetk_wgrad needs in deterministic mode for this problem (needs a CUDA device: SM count)."""
    a = WgradArgs(nhwc(u), nhwc(s), None, mode, algo | TC_FLAGS | TC_DETERMINISTIC, None, 0)
    n = lib().unetk_wgrad_partial_bytes(C.byref(a))
    if n < 0:
        check(int(n))
    return int(n)


def prep_wgrad(u, s, dw, mode, algo=ALGO_AUTO, algo_flops=None, label=None, partial=None) -> Call:
    """partial: scratch tensor (float32) => deterministic split reduction (UNETK_TC_DETERMINISTIC)."""
    a = WgradArgs(nhwc(u), nhwc(s), dw.data_ptr(), mode, algo | TC_FLAGS | (TC_DETERMINISTIC if partial is not None else 0),
                  ptr(partial), 0 if partial is None else partial.numel() * partial.element_size())
    if algo_flops is None:
        taps = (1, 9, 4)[mode]
        algo_flops = 2 * u.shape[0] * u.shape[1] * u.shape[2] * taps * u.shape[3] * s.shape[3]
    return Call("wgrad", 1, algo_flops, lib().unetk_wgrad, (C.byref(a),), tag=mode, keep=(a, u, s, dw, partial), label=label)


def wgrad(u, s, dw, mode, algo=ALGO_AUTO, algo_flops=None, partial=None):
    prep_wgrad(u, s, dw, mode, algo, algo_flops, partial=partial)(stream_ptr())


CHANNEL_SUM_BLOCKS = 512      # include/unetk.h UNETK_CHANNEL_SUM_BLOCKS


def prep_channel_sum(t, out=None, label=None, scratch=None) -> Call:
    """out=None: the destination pointer is patched per call (argument index 1).  scratch (float32, at least
    CHANNEL_SUM_BLOCKS * C elements): ordered, run-to-run reproducible summation."""
    tt = nhwc(t)
    if scratch is not None:
        return Call("reduce", 2, 0, lib().unetk_channel_sum_ordered,
                    (C.byref(tt), ptr(out), scratch.data_ptr(), scratch.numel() * scratch.element_size()),
                    keep=(tt, t, out, scratch), label=label)
    return Call("reduce", 1, 0, lib().unetk_channel_sum, (C.byref(tt), ptr(out)), keep=(tt, t, out), label=label)


def channel_sum(t, out):
    prep_channel_sum(t, out)(stream_ptr())


def bn_stats(z, s, ss):
    _run("bn_stats", 1, 0, lib().unetk_bn_stats, C.byref(nhwc(z)), s.data_ptr(), ss.data_ptr(), stream_ptr())


def prep_bn_finalize(s, ss, count, c, training, gamma, beta, conv_bias, running_mean, running_var, nbt, momentum, eps,
                     scale, shift, mean, invstd, label=None) -> Call:
    a = BnFinalizeArgs(ptr(s), ptr(ss), count, c, 1 if training else 0, ptr(gamma), ptr(beta), ptr(conv_bias),
                       ptr(running_mean), ptr(running_var), ptr(nbt), momentum, eps, ptr(scale), ptr(shift),
                       ptr(mean), ptr(invstd))
    return Call("bn_finalize", 1, 0, lib().unetk_bn_finalize, (C.byref(a),), label=label,
                keep=(a, s, ss, gamma, beta, conv_bias, running_mean, running_var, nbt, scale, shift, mean, invstd))


def bn_finalize(*args):
    prep_bn_finalize(*args)(stream_ptr())


def prep_bn_relu_apply(z, scale, shift, a, pooled=None, pool_idx=None, label=None) -> Call:
    tz, ta, tp = nhwc(z), nhwc(a), nhwc(pooled)
    return Call("bn_apply", 1, 0, lib().unetk_bn_relu_apply,
                (C.byref(tz), scale.data_ptr(), shift.data_ptr(), C.byref(ta), C.byref(tp), ptr(pool_idx)), label=label,
                keep=(tz, ta, tp, z, scale, shift, a, pooled, pool_idx))


def bn_relu_apply(z, scale, shift, a, pooled=None, pool_idx=None):
    prep_bn_relu_apply(z, scale, shift, a, pooled, pool_idx)(stream_ptr())


def prep_bn_relu_bwd(z, dy, dpool, scale, shift, mean, invstd, sums, dz, dgamma=None, dbeta=None, pool_idx=None, reduced=False,
                     label=None):
    """Returns (args struct, [Call, ...]): the reduce pass (unless `reduced`: the sums were already accumulated by the
    producing unetk_conv launch) and the apply pass.  dgamma / dbeta may be None and patched per call."""
    a = BnBwdArgs(nhwc(z), nhwc(dy), nhwc(dpool), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(sums), nhwc(dz),
                  ptr(dgamma), ptr(dbeta), ptr(pool_idx))
    keep = (a, z, dy, dpool, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, pool_idx)
    calls = []
    if not reduced:
        calls.append(Call("bn_bwd_reduce", 1, 0, lib().unetk_bn_relu_bwd_reduce, (C.byref(a),), keep=keep, label=label))
    calls.append(Call("bn_bwd_apply", 1, 0, lib().unetk_bn_relu_bwd_apply, (C.byref(a),), keep=keep, label=label))
    return a, calls


def bn_relu_bwd(z, dy, dpool, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, pool_idx=None, reduced=False):
    """reduced=True: `sums` were already accumulated by the producing unetk_conv launch (bn_reduce=...)."""
    s = stream_ptr()
    for c in prep_bn_relu_bwd(z, dy, dpool, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, pool_idx, reduced)[1]:
        c(s)


def prep_head_fprop(a, w, b, dout, logits, label=None) -> Call:
    ta = nhwc(a)
    return Call("head", 1, 0, lib().unetk_head_fprop, (C.byref(ta), w.data_ptr(), ptr(b), dout, logits.data_ptr()),
                keep=(ta, a, w, b, logits), label=label)


def head_fprop(a, w, b, dout, logits):
    prep_head_fprop(a, w, b, dout, logits)(stream_ptr())


def prep_head_bwd(dlogits, a, w, dout, da, dw=None, db=None, label=None) -> Call:
    """dw / db None: patched per call (arguments 5 and 6 are plain pointers, so the Call carries a tiny struct)."""
    ta, tda = nhwc(a), nhwc(da)
    box = HeadBwdPtrs(ptr(dw), ptr(db))
    call = Call("head", 1, 0, _head_bwd_boxed, (dlogits.data_ptr(), C.byref(ta), w.data_ptr(), dout, C.byref(tda), box),
                keep=(ta, tda, box, dlogits, a, w, da, dw, db), label=label)
    return call, box


class HeadBwdPtrs(C.Structure):
    _fields_ = [("dw", C.c_void_p), ("db", C.c_void_p)]


def _head_bwd_boxed(dlogits, ta, w, dout, tda, box, stream):
    return lib().unetk_head_bwd(dlogits, ta, w, dout, tda, box.dw, box.db, stream)


def head_bwd(dlogits, a, w, dout, da, dw, db):
    _run("head", 1, 0, lib().unetk_head_bwd, dlogits.data_ptr(), C.byref(nhwc(a)), w.data_ptr(), dout, C.byref(nhwc(da)),
         dw.data_ptr(), ptr(db), stream_ptr())


def prep_bn_relu_head_fprop(z, scale, shift, a, w_head, b_head, dout, logits, label=None) -> Call:
    tz, ta = nhwc(z), nhwc(a)
    return Call("bn_apply", 1, 0, lib().unetk_bn_relu_head_fprop,
                (C.byref(tz), scale.data_ptr(), shift.data_ptr(), C.byref(ta), w_head.data_ptr(), ptr(b_head), dout,
                 logits.data_ptr()), keep=(tz, ta, z, scale, shift, a, w_head, b_head, logits), label=label)


def bn_relu_head_fprop(z, scale, shift, a, w_head, b_head, dout, logits):
    """BatchNorm apply + ReLU of the last block fused with the head forward; ``a=None`` skips storing the activation."""
    prep_bn_relu_head_fprop(z, scale, shift, a, w_head, b_head, dout, logits)(stream_ptr())


def prep_head_bn_bwd(dlogits, z, w_head, dout, scale, shift, mean, invstd, sums, dz, dgamma=None, dbeta=None, dw_head=None,
                     db_head=None, label=None):
    """Head backward fused with the BatchNorm backward of the block that feeds the head (two launches).
    Returns (args struct, [reduce Call, apply Call]); the four gradient pointers may be None and patched per call."""
    a = HeadBnBwdArgs(nhwc(z), dlogits.data_ptr(), w_head.data_ptr(), dout, ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                      ptr(sums), nhwc(dz), ptr(dgamma), ptr(dbeta), ptr(dw_head), ptr(db_head))
    keep = (a, dlogits, z, w_head, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, dw_head, db_head)
    return a, [Call("bn_bwd_reduce", 1, 0, lib().unetk_head_bn_bwd_reduce, (C.byref(a),), keep=keep, label=label),
               Call("bn_bwd_apply", 1, 0, lib().unetk_head_bn_bwd_apply, (C.byref(a),), keep=keep, label=label)]


def head_bn_bwd(dlogits, z, w_head, dout, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, dw_head, db_head):
    s = stream_ptr()
    for c in prep_head_bn_bwd(dlogits, z, w_head, dout, scale, shift, mean, invstd, sums, dz, dgamma, dbeta, dw_head,
                              db_head)[1]:
        c(s)


def prep_bias_sigmoid_fwd(z, bias, dout, out, label=None) -> Call:
    tz = nhwc(z)
    return Call("recon", 1, 0, lib().unetk_bias_sigmoid_fwd, (C.byref(tz), ptr(bias), dout, out.data_ptr()),
                keep=(tz, z, bias, out), label=label)


def prep_bias_sigmoid_bwd(dy, out, dout, dz, dbias, label=None) -> Call:
    tz = nhwc(dz)
    return Call("recon", 1, 0, lib().unetk_bias_sigmoid_bwd, (dy.data_ptr(), out.data_ptr(), dout, C.byref(tz), ptr(dbias)),
                keep=(tz, dy, out, dz, dbias), label=label)


def prep_bilinear_up(src, dst, backward=False, label=None) -> Call:
    """forward: dst = interpolate(src); backward: `src` receives the gradient gathered from `dst` (= d dst)."""
    ts, td = nhwc(src), nhwc(dst)
    if backward:
        return Call("bilinear", 1, 0, lib().unetk_bilinear_up_bwd, (C.byref(td), C.byref(ts)), keep=(ts, td, src, dst), label=label)
    return Call("bilinear", 1, 0, lib().unetk_bilinear_up_fwd, (C.byref(ts), C.byref(td)), keep=(ts, td, src, dst), label=label)


def bilinear_up(src, dst, backward=False):
    prep_bilinear_up(src, dst, backward)(stream_ptr())


def prompt_compose_fwd(clip_logits, mask_logits, out):
    n, _, h, w = clip_logits.shape
    _run("prompt", 1, 0, lib().unetk_prompt_compose_fwd, clip_logits.data_ptr(), mask_logits.data_ptr(), n, h, w, out.data_ptr(),
         stream_ptr())


def prompt_compose_bwd(clip_logits, mask_logits, dfinal, dmask):
    n, _, h, w = clip_logits.shape
    _run("prompt", 1, 0, lib().unetk_prompt_compose_bwd, clip_logits.data_ptr(), mask_logits.data_ptr(), dfinal.data_ptr(), n, h, w,
         dmask.data_ptr(), stream_ptr())


def nvls_allreduce(multicast_ptr: int, n_elems: int, rank: int, world: int, scale: float = 1.0):
    """Two-shot NVLS all-reduce of a symmetric fp32 buffer (the caller brackets it with cross-rank barriers)."""
    _run("allreduce", 1, 0, lib().unetk_nvls_allreduce_f32, multicast_ptr, n_elems, rank, world, scale, stream_ptr())


def prep_nchw_to_nhwc(src, dst, label=None) -> Call:
    td = nhwc(dst)
    return Call("layout", 1, 0, lib().unetk_nchw_to_nhwc, (src.data_ptr(), C.byref(td)), keep=(td, src, dst), label=label)


def prep_nhwc_to_nchw(src, dst, label=None) -> Call:
    ts = nhwc(src)
    return Call("layout", 1, 0, lib().unetk_nhwc_to_nchw, (C.byref(ts), dst.data_ptr()), keep=(ts, src, dst), label=label)


def dice_ce_fwd(args):
    _run("loss", 2, 0, lib().unetk_dice_ce_fwd, C.byref(args), stream_ptr())


def dice_ce_bwd(args):
    _run("loss", 1, 0, lib().unetk_dice_ce_bwd, C.byref(args), stream_ptr())


def eval_image_table(metas, device):
    """Device array of ``unetk_eval_image`` from the metadata dictionaries of ``resize_with_padding``
    (utils/utils.py:42-47).  Returns (table tensor, total output pixels, max output pixels)."""
    rows, off, mx = [], 0, 0
    for m in metas:
        left, top, _, _ = m["pad"]
        new_h, new_w = m["new_size"]
        oh, ow = m["original_size"]
        rows.append((int(top) | (int(left) << 32), int(new_h) | (int(new_w) << 32), int(oh) | (int(ow) << 32), off))
        off += oh * ow
        mx = max(mx, oh * ow)
    # four int64 words per image = the 32-byte struct (little endian: low word first)
    table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(device, non_blocking=True)
    return table, off, mx


def crop_resize(src, table, max_out_pixels, mode, out):
    n, c, th, tw = src.shape
    _run("eval", 1, 0, lib().unetk_crop_resize, src.data_ptr(), n, c, th, tw, table.data_ptr(), max_out_pixels, mode,
         out.data_ptr(), stream_ptr())


def eval_loss_metrics(args):
    _run("eval", 2, 0, lib().unetk_eval_loss_metrics, C.byref(args), stream_ptr())


def argmax_confusion(pred, label, n, c, h, w, counts, argmax_out, status):
    _run("metrics", 1, 0, lib().unetk_argmax_confusion, pred.data_ptr(), label.data_ptr(), n, c, h, w, counts.data_ptr(),
         ptr(argmax_out), status.data_ptr(), stream_ptr())
