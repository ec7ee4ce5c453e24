"""ctypes binding of libgapart_b200.so.

The prototypes are read from include/gapart_b200.h, so the header is the single source of truth
for the C ABI; loading fails loudly if the shared library or any declared symbol is missing
(there is no CPU fallback: the product path is the CUDA library or nothing).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "gapart_b200.h")
# GAPART_LIB: perf tooling only (tools/perf_wgrad.py compares build variants of one kernel)
LIB_PATH = os.environ.get("GAPART_LIB") or os.path.join(_HERE, "libgapart_b200.so")

_lib = None
_protos: Dict[str, Tuple[str, List[str]]] = {}


class GapartError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """-> {symbol: (return type, [parameter types])} for every `gp_*` prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?[\w ]+?\*?)\s*(gp_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        ptypes: List[str] = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    ptypes.append("ptr")
                else:
                    toks = p.split()
                    ptypes.append(" ".join(toks[:-1]))  # drop the parameter name
        protos[name] = (ret, ptypes)
    return protos


_CT = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "uint32_t": ctypes.c_uint32,
    "int64_t": ctypes.c_int64,
    "ptr": ctypes.c_void_p,
}


def load():
    """dlopen the library and bind every prototype of the header. Needs no GPU."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GapartError(
            f"{LIB_PATH} is missing - run `python __graft_entry__.py` (nvcc, sm_100a). "
            "gapartnet_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    missing = []
    for name, (ret, ptypes) in _protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.argtypes = [_CT[t] for t in ptypes]
        if "char" in ret:
            fn.restype = ctypes.c_char_p
        elif ret == "long long":
            fn.restype = ctypes.c_longlong
        else:
            fn.restype = ctypes.c_int
    if missing:
        raise GapartError(f"{LIB_PATH} does not export: {missing}")
    ver = lib.gp_version()
    m = re.search(r"#define\s+GP_ABI_VERSION\s+(\d+)", open(HEADER).read())
    if m and int(m.group(1)) != ver:
        raise GapartError(f"ABI mismatch: header {m.group(1)} vs library {ver} (stale .so?)")
    _lib = lib
    return lib


def symbols() -> List[str]:
    return sorted(parse_header().keys())


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gp_last_error()
        raise GapartError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


# int-returning queries (their value is an answer, not a status code)
_QUERIES = ("gp_version", "gp_device_sms", "gp_conv_tc_supported", "gp_conv_wgrad_tc_supported",
            "gp_conv_tc_ksplit", "gp_bn_cluster_ok", "gp_conv_win_supported", "gp_conv_wgrad_win_supported")


class _Caller:
    """lib.gp_xxx(...) with automatic status check."""

    def __getattr__(self, name):
        lib = load()
        fn = getattr(lib, name)
        ret = _protos[name][0]
        if ret == "int" and name not in _QUERIES:
            def call(*args, _fn=fn, _name=name):
                rc = _fn(*args)
                if rc != 0:
                    check(rc, _name)
            setattr(self, name, call)
            return call
        setattr(self, name, fn)
        return fn


C = _Caller()
