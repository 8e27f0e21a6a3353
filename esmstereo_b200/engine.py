"""Serialised engines: the forward of a model at one input shape as a flat list of C-ABI calls plus a memory plan.

The reference deploys ESMStereo by exporting ONNX (`onnx_transformed.py:48-51`), building a TensorRT `.plan` offline and
running it from C++ (`kitti_publisher/src/kitti_publisher_cuda_node.cpp:177-263,364-383`: load engine, bind three device
buffers, `enqueueV3`).  The B200-native equivalent here: every device operation of the forward already goes through
`libesm_b200.so`'s C ABI, so one recorded forward IS the engine --

    export_engine(model, (1, 3, 384, 1248), "esm_L_384x1248.esmeng", train_status=False)

records each call (function name + arguments, device pointers rewritten as (segment, offset) of the allocator's
segments), snapshots the live device blocks before the forward (weights, packed weights, constants) and writes one
file; `host/esm_host.cpp` maps the segments, uploads the blocks and replays the calls with no Python, no torch and no
allocation at run time.

File layout (little endian):  "ESMENG01" | u32 nseg, u64 size[nseg] | u32 nblk, {u32 seg, u64 off, u64 n, bytes}[nblk] |
u32 nfn, {u16 len, name}[nfn] | u32 ncall, {u16 fn, u16 nargs, arg[nargs]}[ncall] | io: left, right {u32 seg, u64 off, u64 n},
u32 nout, {u32 seg, u64 off, u32 ndim, u32 dim[ndim]}[nout] | u32 len, utf-8 json (shape, model, version) |
u32 len, engine-plan text (esm_conv_plans_import: the replay takes the engines this forward took)
arg: tag byte 'i' i32 | 'q' i64 | 'f' f32 | 'p' u32 seg (0xffffffff = NULL), u64 off | 'S' (the stream) |
     's' u32 n, bytes, u16 nfix, {u32 at, u32 seg, u64 off}[nfix] (a struct with its pointer fields listed) | 'h' u32 n, bytes (host array)
"""
from __future__ import annotations

import bisect
import ctypes as C
import json
import struct
from typing import Dict, List, Tuple

import torch

from . import _lib

MAGIC = b"ESMENG01"
NULL_SEG = 0xFFFFFFFF


class _Segments:
    def __init__(self, device) -> None:
        snap = [s for s in torch.cuda.memory_snapshot() if s["device"] == device.index]
        self.segs = sorted((s["address"], s["total_size"]) for s in snap)
        self.starts = [a for a, _ in self.segs]
        self.blocks = []  # live blocks (address, size)
        for s in snap:
            addr = s["address"]
            for b in s["blocks"]:
                if b["state"] == "active_allocated":
                    self.blocks.append((addr, b["size"]))
                addr += b["size"]

    def locate(self, ptr: int) -> Tuple[int, int]:
        i = bisect.bisect_right(self.starts, ptr) - 1
        if i < 0 or ptr >= self.segs[i][0] + self.segs[i][1]:
            raise RuntimeError("engine export: device pointer 0x%x is outside every allocator segment" % ptr)
        return i, ptr - self.segs[i][0]


class _Recorder:
    """Stands in for the ctypes library handle: records every call, then forwards it."""

    def __init__(self, real) -> None:
        self._real = real
        self.calls: List[Tuple[str, tuple]] = []

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if name not in _lib.SIGNATURES or not name.endswith(("_f32", "_u16")):
            return fn

        def wrapped(*args):
            self.calls.append((name, tuple(self._freeze(a) for a in args)))
            return fn(*args)

        return wrapped

    @staticmethod
    def _freeze(a):
        obj = getattr(a, "_obj", None)  # C.byref(struct)
        if obj is not None:
            return ("struct", type(obj), bytes(obj))
        if isinstance(a, C.Array):
            return ("host", bytes(a))
        return a


def _struct_pointer_fields(cls, base: int = 0):
    """(byte offset) of every pointer field of a ctypes Structure, nested structures and arrays included."""
    out = []
    for name, typ in cls._fields_:
        off = base + getattr(cls, name).offset
        if typ is C.c_void_p:
            out.append(off)
        elif isinstance(typ, type) and issubclass(typ, C.Structure):
            out += _struct_pointer_fields(typ, off)
        elif isinstance(typ, type) and issubclass(typ, C.Array) and issubclass(typ._type_, C.Structure):
            for i in range(typ._length_):
                out += _struct_pointer_fields(typ._type_, off + i * C.sizeof(typ._type_))
    return out


def _memcpy_d2h(ptr: int, n: int) -> bytes:
    buf = (C.c_char * n)()
    _lib.check(_lib.lib().esm_download(C.addressof(buf), ptr, n), "download")
    return bytes(buf)


def export_engine(model, shape, path: str, **fwd_kw) -> Dict[str, object]:
    """Record one forward of `model` on inputs of `shape` and write the engine file.  Returns a summary."""
    dev = next(model.parameters()).device
    left = torch.zeros(shape, device=dev)
    right = torch.zeros(shape, device=dev)
    with torch.no_grad():
        for _ in range(2):  # packs the weights, pins the plans, grows the allocator's segments to their final size
            model(left, right, **fwd_kw)
    torch.cuda.synchronize(dev)
    _lib.lib()
    real = _lib._lib
    rec = _Recorder(real)
    _lib._lib = rec
    try:
        with torch.no_grad():
            out = model(left, right, **fwd_kw)
        torch.cuda.synchronize(dev)
    finally:
        _lib._lib = real
    outs = [out] if isinstance(out, torch.Tensor) else [t for t in out]
    segs = _Segments(dev)  # after the forward: every segment the replay touches exists; `out` and the inputs are live
    keep = {t.data_ptr() for t in outs}
    io_ptrs = {left.data_ptr(), right.data_ptr()}

    w = bytearray()
    w += MAGIC
    w += struct.pack("<I", len(segs.segs))
    for _, size in segs.segs:
        w += struct.pack("<Q", size)
    # live blocks = parameters, packed weights, cached constants (+ whatever else Python still holds: harmless);
    # the inputs and outputs themselves carry no state
    blocks = [(a, n) for a, n in segs.blocks if a not in keep and a not in io_ptrs]
    w += struct.pack("<I", len(blocks))
    total_const = 0
    for a, n in blocks:
        s, off = segs.locate(a)
        w += struct.pack("<IQQ", s, off, n)
        w += _memcpy_d2h(a, n)
        total_const += n
    names = sorted({c[0] for c in rec.calls})
    w += struct.pack("<I", len(names))
    for nme in names:
        b = nme.encode()
        w += struct.pack("<H", len(b)) + b
    w += struct.pack("<I", len(rec.calls))
    for nme, args in rec.calls:
        argtypes = _lib.SIGNATURES[nme][1]
        w += struct.pack("<HH", names.index(nme), len(args))
        for i, (a, t) in enumerate(zip(args, argtypes)):
            if isinstance(a, tuple) and a[0] == "struct":
                _, cls, raw = a
                fixes = []
                for off in _struct_pointer_fields(cls):
                    ptr = struct.unpack_from("<Q", raw, off)[0]
                    if ptr:
                        fixes.append((off,) + segs.locate(ptr))
                w += b"s" + struct.pack("<I", len(raw)) + raw + struct.pack("<H", len(fixes))
                for off, s, o in fixes:
                    w += struct.pack("<IIQ", off, s, o)
            elif isinstance(a, tuple) and a[0] == "host":
                w += b"h" + struct.pack("<I", len(a[1])) + a[1]
            elif t is C.c_void_p:
                if i == len(args) - 1:
                    w += b"S"
                elif a is None or a == 0:
                    w += b"p" + struct.pack("<IQ", NULL_SEG, 0)
                else:
                    w += b"p" + struct.pack("<IQ", *segs.locate(int(a)))
            elif t is C.c_int:
                w += b"i" + struct.pack("<i", int(a))
            elif t is C.c_longlong:
                w += b"q" + struct.pack("<q", int(a))
            elif t is C.c_float:
                w += b"f" + struct.pack("<f", float(a))
            else:
                raise RuntimeError("engine export: argument %d of %s has an unsupported type %r" % (i, nme, t))
    for t in (left, right):
        w += struct.pack("<IQQ", *segs.locate(t.data_ptr()), t.numel() * 4)
    w += struct.pack("<I", len(outs))
    for t in outs:
        assert t.is_contiguous(), "engine export: outputs must be contiguous"
        w += struct.pack("<IQI", *segs.locate(t.data_ptr()), t.dim())
        for d in t.shape:
            w += struct.pack("<I", int(d))
    meta = json.dumps({"shape": list(shape), "model": type(model).__name__, "abi_version": int(real.esm_version()), "calls": len(rec.calls)}).encode()
    w += struct.pack("<I", len(meta)) + meta
    plans = _lib.export_plans().encode()
    w += struct.pack("<I", len(plans)) + plans
    with open(path, "wb") as f:
        f.write(w)
    return {"calls": len(rec.calls), "functions": names, "segments": len(segs.segs), "reserved_bytes": sum(n for _, n in segs.segs),
            "state_bytes": total_const, "file_bytes": len(w)}
