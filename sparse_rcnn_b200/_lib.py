"""ctypes binding of the C-ABI library (include/scn_b200.h <-> csrc/libscn_b200.so).

The prototypes are parsed from the public header, so the Python binding cannot drift
from the declared ABI.  There is NO fallback: if the shared library is missing the
import of any compute path raises (the product must fail loudly without its CUDA
extension).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "scn_b200.h")
# SCN_B200_LIB: load another build of the same ABI (kernel experiments, scripts/build_variant.py); never a fallback
LIB_PATH = os.environ.get("SCN_B200_LIB") or os.path.join(_HERE, "csrc", "libscn_b200.so")

_PROTO = re.compile(r"^(const char\*|int64_t|int)\s+(scn_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for ret, name, args in _PROTO.findall(src):
        restype = {"const char*": ctypes.c_char_p, "int64_t": ctypes.c_int64, "int": ctypes.c_int}[ret]
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or a.startswith("scn_stream_t"):
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("int64_t"):
                    argtypes.append(ctypes.c_int64)
                elif a.startswith("uint32_t"):
                    argtypes.append(ctypes.c_uint32)
                elif a.startswith("float"):
                    argtypes.append(ctypes.c_float)
                elif a.startswith("int"):
                    argtypes.append(ctypes.c_int)
                else:
                    raise RuntimeError("unparsed argument %r of %s" % (a, name))
        out[name] = (restype, argtypes)
    return out


class _Lib:
    def __init__(self):
        self._dll = None
        self._protos = None
        self._fns = {}

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "sparse_rcnn_b200: CUDA extension %s not built; run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH)
        self._protos = parse_header()
        for name, (restype, argtypes) in self._protos.items():
            fn = getattr(dll, name)          # AttributeError => header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        self._dll = dll
        return dll

    def fn(self, name):
        f = self._fns.get(name)
        if f is None:
            f = getattr(self.load(), name)
            self._fns[name] = f
        return f

    def call(self, name, *args):
        rc = self.fn(name)(*args)
        if rc != 0:
            msg = self.load().scn_last_error()
            raise RuntimeError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))

    def raw(self, name):
        return self.fn(name)


LIB = _Lib()
_FNS = LIB._fns


def call(name, *args):
    f = _FNS.get(name)
    if f is None:
        f = LIB.fn(name)
    rc = f(*args)
    if rc != 0:
        msg = LIB.load().scn_last_error()
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))


def raw(name):
    return LIB.fn(name)
