"""ctypes binding of libfreqair.so.

The prototypes are read from ``include/freqair.h`` (the single source of truth for the C ABI), so
the Python side cannot drift from the header.  There is no fallback: if the library is missing or
fails to load, every op raises ``RuntimeError`` (the product never computes on the CPU).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'freqair.h')
LIB_PATH = os.path.join(HERE, 'libfreqair.so')


class FaGemmEpilogue(ctypes.Structure):
    _fields_ = [('bias', ctypes.c_void_p), ('act', ctypes.c_int), ('act_param', ctypes.c_float),
                ('aux', ctypes.c_void_p), ('ldaux', ctypes.c_int64), ('aux_act', ctypes.c_int),
                ('aux_param', ctypes.c_float), ('rowscale', ctypes.c_void_p), ('rows_per_scale', ctypes.c_int),
                ('residual', ctypes.c_void_p), ('ldr', ctypes.c_int64), ('accumulate', ctypes.c_int),
                ('alpha', ctypes.c_float), ('preact', ctypes.c_void_p), ('ldpre', ctypes.c_int64),
                ('a_rowsum', ctypes.c_void_p), ('a_kscale', ctypes.c_void_p), ('a_k_rows_per_scale', ctypes.c_int),
                ('b_is_tf32', ctypes.c_int)]


_SCALARS = {'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'float': ctypes.c_float, 'double': ctypes.c_double,
            'size_t': ctypes.c_size_t, 'fa_stream_t': ctypes.c_void_p}


def _ctype(decl: str):
    decl = decl.replace('const', '').strip()
    if '*' in decl:
        base = decl.split('*')[0].strip()
        if base == 'FaGemmEpilogue':
            return ctypes.POINTER(FaGemmEpilogue)
        if base == 'char':
            return ctypes.c_char_p
        return ctypes.c_void_p
    base = decl.split()[0]
    return _SCALARS[base]


def parse_header(path=HEADER):
    """{name: (restype, [argtypes])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'typedef struct.*?}\s*\w+;', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'([\w\s\*]+?)\b(fa_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith('typedef'):
            continue
        if 'char' in ret:
            restype = ctypes.c_char_p
        elif ret == 'void':
            restype = None
        else:
            restype = _SCALARS[ret.split()[-1]]
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                # drop the parameter name
                mm = re.match(r'(.*?[\*\s])(\w+)$', a)
                argtypes.append(_ctype(mm.group(1) if mm else a))
        out[name] = (restype, argtypes)
    return out


_lib = None
_protos = None


def _stale():
    """True when a source of the library is newer than the built .so (only meaningful where the sources are)."""
    csrc = os.path.join(HERE, 'csrc')
    if not os.path.isdir(csrc):
        return False
    built = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [HEADER]
    return any(os.path.getmtime(f) > built for f in srcs)


def abi_version_of_header(path=HEADER):
    m = re.search(r'#define\s+FREQAIR_ABI_VERSION\s+(\d+)', open(path).read())
    return int(m.group(1)) if m else None


def load():
    """Load libfreqair.so.  If a source file is newer than the .so and nvcc is present the library is rebuilt in-tree
    first (``build.build()``); without nvcc a stale library is an error, not something to load silently.  After loading,
    every prototype of ``include/freqair.h`` must resolve and ``fa_abi_version()`` must equal the header's
    ``FREQAIR_ABI_VERSION`` (a library built from another revision of the header is refused)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if os.path.exists(LIB_PATH) and _stale():
        from . import build as _build
        nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
        if os.path.exists(nvcc):
            _build.build()
        elif os.environ.get('FREQAIR_ALLOW_STALE') != '1':
            raise RuntimeError(f'{LIB_PATH} is older than its sources and nvcc ({nvcc}) is not available to rebuild it '
                               '(set FREQAIR_ALLOW_STALE=1 to load it anyway)')
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). '
                           'freqair has no CPU or PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argtypes) in _protos.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    want, got = abi_version_of_header(), lib.fa_abi_version()
    if want != got:
        raise RuntimeError(f'{LIB_PATH} implements ABI version {got}, include/freqair.h declares {want}: rebuild the library')
    _lib = lib
    return lib


def check(rc: int, name: str = ''):
    if rc != 0:
        msg = load().fa_last_error_string()
        raise RuntimeError(f'{name}: error {rc}: {msg.decode() if msg else ""}')
