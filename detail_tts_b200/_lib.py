"""ctypes binding of the C-ABI kernel library (include/dtts.h -> detail_tts_b200/libdtts.so).

The struct layouts are parsed from the header itself and cross-checked against `dtts_sizeof()` in
the loaded library, so the binding cannot drift from the ABI.  There is no CPU fallback: a missing
library or a failed launch raises.

A `Plan` records launches (function pointer + filled parameter struct) once and replays them with
one ctypes call each: the diffusion eval and the GPT decode step use fixed device buffers, so the
per-step Python cost is a loop over prepared structs rather than re-marshalling ~150 calls.
"""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "dtts.h")
LIB_PATH = os.path.join(HERE, "libdtts.so")

_CT = {"int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64}

# enums of dtts.h
ACT_NONE, ACT_GELU_NEW, ACT_RELU, ACT_SILU, ACT_MISH, ACT_LRELU, ACT_TANH, ACT_LOG_CLAMP = 0, 1, 2, 3, 4, 5, 6, 7
ACT_PAIR_TANH_SIGMOID, ACT_PAIR_GLU = 16, 17
BIAS_NONE, BIAS_RELPOS_TABLE, BIAS_WINDOW_REL = 0, 1, 2


def parse_header(path=HEADER):
    """Return ({struct_name: [(field, ctype, is_ptr)]}, [function names]) from dtts.h."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            m = re.match(r"^(const\s+)?(\w+)\s*(\*?)\s*(.*)$", stmt)
            base, star, rest = m.group(2), m.group(3), m.group(4)
            for decl in rest.split(","):
                decl = decl.strip()
                ptr = bool(star) or decl.startswith("*")
                fname = decl.lstrip("*").strip()
                if ptr:
                    fields.append((fname, ctypes.c_void_p, True))
                else:
                    fields.append((fname, _CT[base], False))
        structs[name] = fields
    funcs = re.findall(r"\bint\s+(dtts_\w+)\s*\(\s*const\s+(\w+)\s*\*\s*\w+\s*,\s*void\s*\*\s*stream\s*\)", src)
    return structs, funcs


class DttsError(RuntimeError):
    pass


class Plan:
    """A recorded launch sequence over fixed device buffers."""

    def __init__(self, lib):
        self.lib = lib
        self.calls = []      # (cfunc, struct)
        self.keep = []       # tensors kept alive
        self.n_kernels = 0

    def run(self, stream=None):
        st = ctypes.c_void_p(stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        n0 = self.lib.cdll.dtts_kernel_launches()
        for fn, s in self.calls:
            rc = fn(ctypes.byref(s), st)
            if rc != 0:
                raise DttsError(f"{fn.__name__} failed ({rc}): {self.lib.last_error()}")
        self.n_kernels = self.lib.cdll.dtts_kernel_launches() - n0     # kernels one run (or one graph replay) launches

    def replayed(self):
        """Account for one CUDA-graph replay of this plan (its kernels do not pass through the library's counter)."""
        self.lib.graph_launches += self.n_kernels

    def profile(self, stream=None):
        """Run once with a CUDA-event pair around every launch; returns [(entry point, struct, ms)]."""
        st = torch.cuda.current_stream()
        sp = ctypes.c_void_p(st.cuda_stream)
        evs = []
        for fn, s in self.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = fn(ctypes.byref(s), sp)
            if rc != 0:
                raise DttsError(f"{fn.__name__} failed ({rc}): {self.lib.last_error()}")
            e1.record(st)
            evs.append((fn.__name__, s, e0, e1))
        torch.cuda.synchronize()
        return [(n, s, a.elapsed_time(b)) for n, s, a, b in evs]

    def __len__(self):
        return len(self.calls)


class Lib:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise DttsError(f"{path} not found: build it with `python -m detail_tts_b200.build` "
                            "(there is no CPU fallback for the synthesis path)")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        structs, funcs = parse_header()
        self.struct_fields = structs
        self.structs = {}
        for name, fields in structs.items():
            self.structs[name] = type(name, (ctypes.Structure,), {"_fields_": [(f, t) for f, t, _ in fields]})
        self.cdll.dtts_last_error.restype = ctypes.c_char_p
        self.cdll.dtts_sizeof.argtypes = [ctypes.c_char_p]
        self.funcs = {}
        for fname, sname in funcs:
            fn = getattr(self.cdll, fname)
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            fn.restype = ctypes.c_int
            fn.__name__ = fname
            self.funcs[fname] = (fn, sname)
        for name, cls in self.structs.items():
            n = self.cdll.dtts_sizeof(name.encode())
            if n != ctypes.sizeof(cls):
                raise DttsError(f"ABI mismatch for {name}: header says {ctypes.sizeof(cls)}, library says {n}")
        if self.cdll.dtts_abi_version() != 2:
            raise DttsError("ABI version mismatch")
        self._plan = None
        self.graph_launches = 0

    # -- info ------------------------------------------------------------------------------
    def last_error(self):
        return (self.cdll.dtts_last_error() or b"").decode()

    def launches(self):
        """Kernels of this library launched so far: direct launches + the kernels of replayed CUDA graphs."""
        return int(self.cdll.dtts_kernel_launches()) + self.graph_launches

    def device_info(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = self.cdll.dtts_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        if rc != 0:
            raise DttsError(self.last_error())
        return a.value, b.value, c.value

    # -- call machinery ----------------------------------------------------------------------
    def record(self):
        """Context manager: launches issued inside are recorded into the returned Plan, not run."""
        lib = self

        class _Rec:
            def __enter__(self_):
                assert lib._plan is None, "nested Plan recording"
                lib._plan = Plan(lib)
                return lib._plan

            def __exit__(self_, *exc):
                lib._plan = None
                return False
        return _Rec()

    def call(self, fname, **kw):
        fn, sname = self.funcs[fname]
        s = self.structs[sname]()
        keep = []
        known = set()
        for f, _, is_ptr in self.struct_fields[sname]:
            known.add(f)
            if f not in kw:
                continue
            v = kw[f]
            if is_ptr:
                if v is None:
                    setattr(s, f, None)
                elif isinstance(v, torch.Tensor):
                    if not v.is_cuda:
                        raise DttsError(f"{fname}: field {f} must be a CUDA tensor")
                    setattr(s, f, v.data_ptr())
                    keep.append(v)
                else:
                    setattr(s, f, int(v))
            else:
                setattr(s, f, v)
        extra = set(kw) - known
        if extra:
            raise DttsError(f"{fname}: unknown fields {sorted(extra)}")
        if self._plan is not None:
            self._plan.calls.append((fn, s))
            self._plan.keep.extend(keep)
            return
        rc = fn(ctypes.byref(s), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise DttsError(f"{fname} failed ({rc}): {self.last_error()}")


_LIB = None


def lib():
    """The process-wide library handle (loads on first use; raises if the .so is missing)."""
    global _LIB
    if _LIB is None:
        _LIB = Lib(os.environ.get("DTTS_LIB") or LIB_PATH)   # DTTS_LIB: an A/B build of the same ABI (development only)
    return _LIB
