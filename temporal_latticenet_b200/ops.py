"""Dense / reduction primitives shared by the lattice modules: fp32 matmul, GroupNorm(+ReLU) over
lattice vertices, segmented max / add / mean.  GPU only."""
import ctypes
import os
import threading
import weakref

import torch
from torch.autograd import Function

from . import _lib


def matmul(a, b):
    """fp32 GEMM.  TF32 stays off so the fp32 parity tolerance (rel 1e-4 on accumulations) holds; the
    reference stack (torch 1.7.1 on Ampere) defaulted to TF32 here, so this is >= its precision."""
    return torch.mm(a, b)


def no_grad_path(*tensors):
    """True when nothing will be differentiated: the fused inference kernels may be used."""
    return not (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors))


def linear(x, weight, bias=None, relu_out=False, rows_dev=None):
    """x @ weight.T + bias.  Without autograd and for tensor-core friendly shapes this is the S = 1 case
    of the fused tcgen05 kernel (nn.Linear weights are K-major already); otherwise cuBLAS fp32.
    rows_dev (static-capacity mode): device pointer of the live row count of x when it is NOT the count of its
    capacity class -- a hidden state carries the previous frame's count; rows beyond it are never read."""
    if (x.dim() == 2 and x.is_cuda and no_grad_path(x, weight, bias) and conv_tc_supported(x.shape[1], weight.shape[0], False)
            and x.shape[0] > 0):
        out = conv_tc(x, None, k_major(weight, transposed=True), bias=None if bias is None else bias.detach(), rows_dev=rows_dev)
        return torch.relu_(out) if relu_out else out
    out = torch.nn.functional.linear(x, weight, bias)
    return torch.relu(out) if relu_out else out


def gn_groups(nr_channels):
    """32 groups, or C/2 when C is not divisible by 32."""
    return 32 if nr_channels % 32 == 0 else int(nr_channels / 2)


class _GroupNormRelu(Function):
    """GroupNorm over [1,C,V] (statistics over C/G channels x ALL vertices), optionally fused with
    the ReLU that always follows it in GnRelu* modules.  Two kernels: double-precision group sums,
    then a single read-modify-write pass."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, relu):
        x = x.contiguous().float()
        V, C = x.shape
        lib = _lib.load()
        p = _lib.ptr
        sums = torch.empty(groups, 2, dtype=torch.float64, device=x.device)
        y = torch.empty_like(x)
        vd = _lib.rows_dev(V)
        _lib.check(lib.ltn_gn_stats(p(x), V, vd, C, groups, p(sums), _lib.stream()), "ltn_gn_stats")
        _lib.check(lib.ltn_gn_apply(p(x), V, vd, C, groups, p(sums), p(gamma), p(beta), float(eps), 1 if relu else 0,
                                    p(y), _lib.stream()), "ltn_gn_apply")
        ctx.save_for_backward(x, gamma, sums, y)
        ctx.cfg = (groups, eps, relu)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, sums, y = ctx.saved_tensors
        groups, eps, relu = ctx.cfg
        V, C = x.shape
        gy = gy.contiguous().float()
        chan = torch.empty(C, 2, dtype=torch.float64, device=x.device)
        gx = torch.empty_like(x)
        p = _lib.ptr
        _lib.check(_lib.load().ltn_gn_bwd(p(x), p(gy), p(y) if relu else None, V, C, groups, p(sums), p(gamma), float(eps), p(chan), p(gx),
                                          _lib.stream()), "ltn_gn_bwd")
        chan = chan.float()
        return gx, chan[:, 1].contiguous(), chan[:, 0].contiguous(), None, None, None


def group_norm(x, gamma, beta, groups, eps=1e-5, relu=False):
    return _GroupNormRelu.apply(x, gamma, beta, groups, eps, relu)


def scatter_max(src, index, dim_size=None):
    """torch_scatter.scatter_max(src[R,C], index[R], dim=0) semantics: out 0 and argmax = R for empty
    segments, ids < 0 folded to 0 by the caller, smallest row wins ties.  Differentiable: the
    value is re-gathered from src at argmax so autograd routes the gradient to the arg-max rows."""
    R, C = src.shape
    idx32 = index.to(torch.int32).contiguous()
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
    V = int(dim_size)
    srcd = src.detach().contiguous().float()
    packed = torch.empty(V, C, dtype=torch.int64, device=src.device)
    out = torch.empty(V, C, dtype=torch.float32, device=src.device)
    arg = torch.empty(V, C, dtype=torch.int64, device=src.device)
    p = _lib.ptr
    _lib.check(_lib.load().ltn_scatter_max(p(srcd), p(idx32), R, C, V, p(packed), p(out), p(arg), _lib.stream()),
               "ltn_scatter_max")
    if torch.is_grad_enabled() and src.requires_grad:
        empty = arg >= R
        val = src.gather(0, arg.clamp(max=max(R - 1, 0)))
        out = torch.where(empty, torch.zeros_like(val), val)
    return out, arg


def scatter_add(src, index, dim_size=None, out=None):
    squeeze = src.dim() == 1
    s2 = src.reshape(src.shape[0], -1).contiguous().float()
    idx32 = index.to(torch.int32).contiguous()
    if out is None:
        if dim_size is None:
            dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
        out = torch.zeros(int(dim_size), s2.shape[1], dtype=torch.float32, device=src.device)
    else:
        out = out.reshape(out.shape[0], -1)
    p = _lib.ptr
    _lib.check(_lib.load().ltn_scatter_add(p(s2), p(idx32), s2.shape[0], s2.shape[1], p(out), out.shape[0],
                                           _lib.stream()), "ltn_scatter_add")
    return out.reshape(-1) if squeeze else out


def scatter_mean(src, index, dim_size=None, out=None):
    total = scatter_add(src, index, dim_size, out)
    ones = torch.ones(index.shape[0], dtype=torch.float32, device=src.device)
    count = scatter_add(ones, index, total.shape[0]).clamp(min=1)
    return total / (count if total.dim() == 1 else count.unsqueeze(1))


# ---------------------------------------------------------------------------------------------------
# fused gather + GEMM on the tensor cores (csrc/ltn_conv.cu)
# ---------------------------------------------------------------------------------------------------
_WT_CACHE = {}
TC_PASSES = 3  # 3 = fp32-parity split (default); 1 = single-pass TF32 (stated separately wherever used)
A_LOG2 = 5     # fp16 operands: activations are staged as x * 2^5 (overflow flag beyond |x| >= 2047, hi/lo floor 1e-9)

# Operand type of the fp32-parity convolution.  "tf32": hi/lo split in tf32 (4 bytes / element) -- always safe.
# "f16": the same 11 + 11 significant bits as fp16 hi/lo (2 bytes / element: half the shared-memory traffic, twice
# the tensor rate).  fp16's exponent range is handled by exact power-of-two scaling plus a device-side range flag,
# so "f16" is only selected by callers that CHECK the flag and redo the window with "tf32" when it is raised
# (runner.py / engine.py); modules called directly stay on "tf32".
class _ThreadLocalDict:
    """dict-like view of per-thread state (the lock-step engine runs one host thread per window in flight while capturing)"""

    def __init__(self, **defaults):
        self._defaults = defaults
        self._tl = threading.local()

    def _d(self):
        d = getattr(self._tl, "d", None)
        if d is None:
            d = self._tl.d = dict(self._defaults)
        return d

    def __getitem__(self, k): return self._d()[k]
    def __setitem__(self, k, v): self._d()[k] = v
    def update(self, other): self._d().update(other)
    def keys(self): return self._d().keys()
    def __iter__(self): return iter(self._d())


_TC = _ThreadLocalDict(mode="tf32", flag=None)


class tc_operands:
    """context manager: with ops.tc_operands("f16", flag_tensor): ...  (flag_tensor: int32 [1] on the device)"""

    def __init__(self, mode, flag=None):
        if mode not in ("tf32", "f16"):
            raise ValueError("operand mode must be 'tf32' or 'f16'")
        if mode == "f16" and flag is None:
            raise ValueError("fp16 operands need a range flag the caller checks")
        self.new = {"mode": mode, "flag": flag}

    def __enter__(self):
        self.old = {k: _TC[k] for k in ("mode", "flag")}
        _TC.update(self.new)
        return self

    def __exit__(self, *exc):
        _TC.update(self.old)
        return False


def conv_tc_supported(C, F, folded_norm=True):
    return C % 32 == 0 and C > 0 and (C <= 256 or not folded_norm) and F % 8 == 0 and F > 0


class SplitWeight:
    """K-major ([F,K]) operand copies of one weight for the tensor-core kernel, built on demand per operand type:
    tf32() -> (hi, lo) fp32 tensors; f16() -> (hi, lo, w_log2) fp16 tensors holding w * 2^w_log2."""

    def __init__(self, w, transposed):
        self.w = w                      # detached, contiguous fp32; [F,K] if transposed else [K,F]
        self.transposed = bool(transposed)
        self.F, self.K = (w.shape if transposed else (w.shape[1], w.shape[0]))
        self._tf32 = None
        self._f16 = None

    def tf32(self):
        if self._tf32 is None:
            hi = torch.empty(self.F, self.K, dtype=torch.float32, device=self.w.device)
            lo = torch.empty(self.F, self.K, dtype=torch.float32, device=self.w.device)
            _lib.check(_lib.load().ltn_split_tf32(_lib.ptr(self.w), self.K, self.F, 1 if self.transposed else 0, _lib.ptr(hi),
                                                  _lib.ptr(lo), _lib.stream()), "ltn_split_tf32")
            self._tf32 = (hi, lo)
        return self._tf32

    def f16(self):
        if self._f16 is None:
            # largest |w| -> just below 2^14 (half overflows at 65504): one host read per weight version
            m = float(self.w.abs().max().item()) if self.w.numel() else 0.0
            log2 = 0
            if m > 0.0 and m != float("inf") and m == m:
                import math
                log2 = max(-60, min(60, 13 - math.frexp(m)[1] + 1))
            hi = torch.empty(self.F, self.K, dtype=torch.float16, device=self.w.device)
            lo = torch.empty(self.F, self.K, dtype=torch.float16, device=self.w.device)
            _lib.check(_lib.load().ltn_split_f16(_lib.ptr(self.w), self.K, self.F, 1 if self.transposed else 0, log2, _lib.ptr(hi),
                                                 _lib.ptr(lo), _lib.stream()), "ltn_split_f16")
            self._f16 = (hi, lo, log2)
        return self._f16

    def __iter__(self):   # `hi, lo = k_major(w)` keeps working
        return iter(self.tf32())


def _cached_split(weight, key, make):
    hit = _WT_CACHE.get(key)
    ver = weight._version
    # the weak reference guards against a recycled id(): only the very same live tensor object hits
    if hit is not None and hit[0] == ver and hit[1]() is weight and hit[3] == weight.data_ptr():
        return hit[2]
    sw = make()
    if len(_WT_CACHE) > 1024:
        for k in [k for k, v in _WT_CACHE.items() if v[1]() is None]:
            del _WT_CACHE[k]
    _WT_CACHE[key] = (ver, weakref.ref(weight), sw, weight.data_ptr())
    return sw


def k_major(weight, transposed=False):
    """SplitWeight of a weight, cached per parameter version.  `transposed=True`: the tensor is [F,K] already
    (nn.Linear layout); otherwise the reference's conv layout [K,F]."""
    return _cached_split(weight, (id(weight), bool(transposed)),
                         lambda: SplitWeight(weight.detach().contiguous().float(), transposed))


def k_major_padded(weight, rows):
    """k_major of an nn.Linear weight [F,K] zero-padded to `rows` output rows (F is not a multiple of 8)"""
    def make():
        F, K = weight.shape
        wp = torch.zeros(rows, K, dtype=torch.float32, device=weight.device)
        wp[:F] = weight.detach()
        return SplitWeight(wp, True)
    return _cached_split(weight, (id(weight), "pad", int(rows)), make)


def gn_sums(x, groups):
    """[G,2] double: per-group sum and sum of squares of x [V,C] (GroupNorm over [1,C,V])"""
    x = x.contiguous()
    sums = torch.empty(groups, 2, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().ltn_gn_stats(_lib.ptr(x), x.shape[0], _lib.rows_dev(x.shape[0]), x.shape[1], groups, _lib.ptr(sums), _lib.stream()),
               "ltn_gn_stats")
    return sums


def conv_tc(x, nbr, wt, nr_rows=None, a_scale=None, a_shift=None, gn=None, relu=False, bias=None, res=None, out=None,
            out_sums=None, passes=None, operands=None, flag=None, rows_dev=None):
    """out[v,:] = sum_s act(x[nbr[v,s],:]) @ W[s] (+bias) (+res); nbr None = plain row-wise GEMM.
    wt = k_major(weight).  gn = (sums [G,2], gamma, beta, eps): GroupNorm of x folded into the gather.
    out_sums [Gout,2] (zeroed): receives the GroupNorm statistics of the output.
    operands: "tf32" / "f16" (default: the mode set by ops.tc_operands); f16 needs C % 64 == 0 and `flag`."""
    x = x.contiguous()
    C = x.shape[1]
    S = 1 if nbr is None else nbr.shape[1]
    F = wt.F
    if wt.K != S * C:
        raise RuntimeError("weight is [%d,%d], expected [%d,%d]" % (F, wt.K, F, S * C))
    Vq = (x.shape[0] if nbr is None else nbr.shape[0]) if nr_rows is None else nr_rows
    if out is None:
        out = torch.empty(Vq, F, dtype=torch.float32, device=x.device)
    p = _lib.ptr
    g_sums = g_gamma = g_beta = None
    g_eps, g_groups = 0.0, 0
    if gn is not None:
        g_sums, g_gamma, g_beta, g_eps = gn
        g_groups = g_sums.shape[0]
    passes = TC_PASSES if passes is None else passes
    vx_dev = _lib.rows_dev(x.shape[0]) if rows_dev is None else rows_dev
    vq_dev = _lib.rows_dev(Vq) if rows_dev is None else rows_dev
    mode = _TC["mode"] if operands is None else operands
    flag = _TC["flag"] if flag is None else flag
    if mode == "f16" and passes == 3 and C % 64 == 0:
        if flag is None:
            raise RuntimeError("fp16 operands need a range flag")
        hi, lo, w_log2 = wt.f16()
        ctx = getattr(_BATCH, "ctx", None)
        if ctx is not None and a_scale is None and x.shape[0] > 0 and Vq > 0:
            # lock-step capture (engine.LockstepRunner): this window's request joins the same layer's requests of the
            # other windows in flight; the coordinator issues ONE launch for all of them (conv_tc_batched) and resumes us
            ctx.request(dict(x=x, Vx=x.shape[0], vx_dev=vx_dev, nbr=nbr, Vq=Vq, vq_dev=vq_dev, C=C, S=S, hi=hi, lo=lo, w_log2=int(w_log2),
                             F=F, g_sums=g_sums, g_gamma=g_gamma, g_beta=g_beta, g_eps=float(g_eps), g_groups=int(g_groups),
                             relu=1 if relu else 0, bias=bias, res=res, out=out, ldo=out.stride(0), out_sums=out_sums, flag=flag))
            if out_sums is not None:
                out._ltn_gn_sums = (_FRAME["id"], out_sums)
            return out
        rc = _lib.load().ltn_conv_tc_f16(p(x), x.shape[0], vx_dev, p(nbr), Vq, vq_dev, C, S, p(hi), p(lo),
                                         int(w_log2), int(A_LOG2), F, p(a_scale), p(a_shift), p(g_sums), p(g_gamma), p(g_beta),
                                         float(g_eps), int(g_groups), 1 if relu else 0, p(bias), p(res), p(out), out.stride(0),
                                         p(out_sums), 0 if out_sums is None else out_sums.shape[0], p(flag), _lib.stream())
        _lib.check(rc, "ltn_conv_tc_f16")
    else:
        hi, lo = wt.tf32()
        rc = _lib.load().ltn_conv_tc(p(x), x.shape[0], vx_dev, p(nbr), Vq, vq_dev, C, S, p(hi), p(lo), F,
                                     p(a_scale), p(a_shift),
                                     p(g_sums), p(g_gamma), p(g_beta), float(g_eps), int(g_groups), 1 if relu else 0, p(bias), p(res),
                                     p(out), out.stride(0), p(out_sums), 0 if out_sums is None else out_sums.shape[0],
                                     passes, _lib.stream())
        _lib.check(rc, "ltn_conv_tc")
    if out_sums is not None:
        out._ltn_gn_sums = (_FRAME["id"], out_sums)   # travels with the tensor to the next layer's GroupNorm
    return out


STAGE_GATHERED = os.environ.get("LTN_CONVB_STAGE", "1") != "0"   # pre-stage the A operand of gathering layers (csrc k_stage_a)
_BATCH = threading.local()   # .ctx: coordinator of a lock-step capture (engine.LockstepRunner), absent otherwise
MAX_BATCH = 8


def _ptr_array(items):
    return (ctypes.c_void_p * len(items))(*[None if t is None else (t.value if isinstance(t, ctypes.c_void_p) else t.data_ptr()) for t in items])


def conv_tc_batched(reqs):
    """ONE launch of the fused fp16-operand convolution for the same layer of several independent windows
    (csrc/ltn_conv_batched.cu).  reqs: the dicts conv_tc hands to the lock-step coordinator, one per window; weights,
    bias and GroupNorm parameters are taken from the first (the windows' models hold identical values)."""
    nb = len(reqs)
    if nb < 1 or nb > MAX_BATCH:
        raise RuntimeError("between 1 and %d problems per batched launch" % MAX_BATCH)
    r0 = reqs[0]
    same = ("C", "S", "F", "w_log2", "g_eps", "g_groups", "relu", "ldo")
    for r in reqs[1:]:
        if any(r[k] != r0[k] for k in same) or (r["nbr"] is None) != (r0["nbr"] is None) or (r["g_sums"] is None) != (r0["g_sums"] is None) \
                or (r["out_sums"] is None) != (r0["out_sums"] is None) or (r["res"] is None) != (r0["res"] is None) \
                or (r["bias"] is None) != (r0["bias"] is None):
            raise RuntimeError("the windows of a lock-step group reached different layers")
    ints = lambda k: (ctypes.c_int * nb)(*[int(r[k]) for r in reqs])   # noqa: E731
    arr = lambda k: _ptr_array([r[k] for r in reqs])                    # noqa: E731
    p = _lib.ptr
    lib = _lib.load()
    x_arr, gn_arr, staged = arr("x"), arr("g_sums"), 0
    if STAGE_GATHERED and r0["nbr"] is not None and r0["C"] <= 256:
        # gathering layer: every row is used by ~9 tiles, so its GroupNorm / ReLU / fp16 hi-lo split is done ONCE here
        # (k_stage_a) instead of nine times in the gather loop; the convolution then only moves the staged quads
        bufs = [torch.empty_like(r["x"]) for r in reqs]
        rc = lib.ltn_stage_a_batched(nb, x_arr, ints("Vx"), arr("vx_dev"), r0["C"], gn_arr, p(r0["g_gamma"]), p(r0["g_beta"]), r0["g_eps"],
                                     r0["g_groups"], r0["relu"], int(A_LOG2), _ptr_array(bufs), arr("flag"), _lib.stream())
        _lib.check(rc, "ltn_stage_a_batched")
        x_arr, staged = _ptr_array(bufs), 1
    rc = lib.ltn_conv_tc_f16_batched(nb, x_arr, ints("Vx"), arr("vx_dev"), arr("nbr"), ints("Vq"), arr("vq_dev"), r0["C"], r0["S"],
                                     p(r0["hi"]), p(r0["lo"]), r0["w_log2"], int(A_LOG2), r0["F"], gn_arr, p(r0["g_gamma"]),
                                     p(r0["g_beta"]), r0["g_eps"], r0["g_groups"], r0["relu"], p(r0["bias"]), arr("res"), arr("out"),
                                     r0["ldo"], arr("out_sums"), 0 if r0["out_sums"] is None else r0["out_sums"].shape[0],
                                     arr("flag"), staged, _lib.stream())
    _lib.check(rc, "ltn_conv_tc_f16_batched")


_SUMS_SLOTS = 96   # accumulators handed out per frame before falling back to individual allocations
_FRAME = _ThreadLocalDict(arena=None, used=0, id=0)
_FRAME_IDS = [0]   # frame ids are unique across threads: statistics attached to a tensor are trusted within their frame only


def begin_frame(owner, device):
    """One memset for all of a frame's [G,2] statistics accumulators.  The arena belongs to `owner` (the
    model instance: every window runner / graph lane has its own, so concurrently replayed graphs never
    share accumulators); statistics attached to tensors are only trusted within the frame that made them."""
    buf = getattr(owner, "_ltn_sums_arena", None)
    if buf is None or buf.device != device:
        buf = torch.zeros(_SUMS_SLOTS, 32, 2, dtype=torch.float64, device=device)
        owner._ltn_sums_arena = buf
    else:
        buf.zero_()
    _FRAME["arena"], _FRAME["used"] = buf, 0
    _FRAME_IDS[0] += 1
    _FRAME["id"] = _FRAME_IDS[0]


def new_sums(nr_channels, device):
    """zeroed [G,2] accumulator for the epilogue statistics of a layer with `nr_channels` outputs"""
    g = gn_groups(nr_channels)
    buf = _FRAME["arena"]
    if buf is not None and buf.device == device and _FRAME["used"] < _SUMS_SLOTS and g <= 32:
        t = buf[_FRAME["used"], :g]
        _FRAME["used"] += 1
        return t
    return torch.zeros(g, 2, dtype=torch.float64, device=device)


def sums_of(x, groups):
    """GroupNorm statistics of x: the ones its producing kernel left behind, else one pass over x"""
    s = getattr(x, "_ltn_gn_sums", None)
    if s is not None and s[0] == _FRAME["id"] and s[1].shape[0] == groups:
        return s[1]
    return gn_sums(x, groups)
