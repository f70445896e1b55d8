"""ctypes binding of libnemar_b200.so (include/nemar_b200.h).

The engine has no CPU fallback: if the shared library is missing, or a call fails, this module raises.
"""
import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG, "libnemar_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "nemar_b200.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
PAD_ZERO, PAD_REFLECT = 0, 1


class NemarTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("pad", C.c_int32), ("cs", C.c_int32), ("coff", C.c_int32), ("dtype", C.c_int32)]


class ConvGeom(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("transposed", C.c_int32)]


class EngineError(RuntimeError):
    pass


def declared_symbols():
    """Every function name the public header declares (used by the export test)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nemar_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("libnemar_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` — the engine has no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.nemar_last_error.restype = C.c_char_p
        _lib.nemar_last_conv_kernel.restype = C.c_char_p
        _lib.nemar_conv2d_wgrad_workspace.restype = C.c_int64
        _lib.nemar_conv2d_fprop_stats_workspace.restype = C.c_int64
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().nemar_last_error()
        raise EngineError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise EngineError("unsupported dtype %s" % t.dtype)


def torch_dtype(code):
    return torch.float32 if code == F32 else torch.bfloat16


def view(t, pad=0, coff=0, c=None):
    """Describe an engine tensor: torch tensor [N, H+2*pad, W+2*pad, Cs] (contiguous) -> nemar_tensor."""
    assert t.dim() == 4 and t.is_contiguous(), "engine tensors are contiguous [N,Hp,Wp,C]"
    n, hp, wp, cs = t.shape
    c = cs - coff if c is None else c
    return NemarTensor(t.data_ptr(), n, hp - 2 * pad, wp - 2 * pad, c, pad, cs, coff, dtype_code(t))


def fptr(t):
    """float* of a contiguous fp32 tensor (or NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.dtype == torch.float32 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def vptr(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


COUNTERS = {"launches": 0}


class KernelTimer:
    """CUDA-event timing of the conv launches on the launching stream (bench.py's live roofline numbers)."""
    CONV = {"nemar_conv2d_fprop": (0, 6, 4, 8), "nemar_conv2d_fprop_ws": (0, 6, 4, 10), "nemar_conv2d_dgrad": (4, 0, 3, 5),
            "nemar_conv2d_wgrad": (0, 1, 2, 6)}

    def __init__(self):
        self.on = False
        self.records = []
        self.min_flops = 2.0e10      # >= 20 GFLOP per launch: the layers that carry ~85 % of the step's FLOPs

    def enable(self, flag):
        self.on = int(flag)          # 1: conv launches (roofline), 2: every engine call (per-op breakdown)
        self.records = []

    def key_and_flops(self, name, args):
        ix, iy, ig, itc = self.CONV[name]        # positions of: layer input side, output side, geometry, use_tc
        x, y, g = args[ix], args[iy], args[ig]
        pix = (x.n * x.h * x.w) if g.transposed else (y.n * y.h * y.w)
        flops = 2.0 * pix * g.cin * g.cout * g.kh * g.kw
        eng = "tc" if int(args[itc]) else "generic"
        op = name.replace("nemar_conv2d_", "").replace("fprop_ws", "fprop")
        return "%s[%s]" % (op, eng), "%s %d->%d k%d s%d%s @%dx%d" % (
            op, g.cin, g.cout, g.kh, g.stride, "T" if g.transposed else "", y.h, y.w), flops

    def collect(self):
        """-> {kernel: {ms, n, flops, top: {geometry: ms}, ops: {pass: {ms, n, flops}}}} (synchronises)"""
        if not self.records:
            return {}
        torch.cuda.synchronize()
        out = {}
        for key, geo, flops, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0, "top": {}, "ops": {}})
            d["ms"] += ms
            d["n"] += 1
            d["flops"] += flops
            d["top"][geo] = d["top"].get(geo, 0.0) + ms
            o = d["ops"].setdefault(geo.split(" ")[0], {"ms": 0.0, "n": 0, "flops": 0.0})      # the pass: fprop / dgrad / wgrad
            o["ms"] += ms
            o["n"] += 1
            o["flops"] += flops
        self.records = []
        return out


TIMER = KernelTimer()


def call(name, *args):
    COUNTERS["launches"] += 1
    if TIMER.on and (name in KernelTimer.CONV or TIMER.on >= 2):
        if name in KernelTimer.CONV:
            key, geo, flops = TIMER.key_and_flops(name, args)
            if TIMER.on == 1 and flops < TIMER.min_flops:      # mode 1 times only the heavy launches (event overhead)
                _call(name, *args)
                return
        else:
            key, flops = name.replace("nemar_", ""), 0.0
            t0 = next((a for a in args if isinstance(a, NemarTensor)), None)
            geo = key if t0 is None else "%s n%d c%d @%dx%d" % (key, t0.n, t0.c, t0.h, t0.w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _call(name, *args)
        e1.record()
        if name in KernelTimer.CONV:
            # key the record on the kernel INSTANCE the library chose (e.g. tc_gather_kernel<256,64,bf16> serves the
            # forward and the data-gradient passes of every 256-channel layer), not on the op class
            kname = lib().nemar_last_conv_kernel()
            if kname:
                key = kname.decode()
        TIMER.records.append((key, geo, flops, e0, e1))
        return
    _call(name, *args)


def _call(name, *args):
    fn = getattr(lib(), name)
    conv = []
    for a in args:
        if isinstance(a, (NemarTensor, ConvGeom)):
            conv.append(C.byref(a))
        elif isinstance(a, float):
            conv.append(C.c_float(a))
        elif isinstance(a, bool):
            conv.append(C.c_int(int(a)))
        elif isinstance(a, int):
            conv.append(C.c_int64(a) if abs(a) > 0x7FFFFFFF else C.c_int(a))
        elif a is None:
            conv.append(C.c_void_p(0))
        else:
            conv.append(a)
    check(fn(*conv), name)


def i64(v):
    return C.c_int64(int(v))


def u64(v):
    return C.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)
