"""Host-logic dry run: a stub shared library exporting every C-ABI symbol (all return IPN_OK) lets
the Python orchestration (arena, engine, autograd wiring, trainer) run end to end on CPU tensors.
Numerics are garbage by construction -- this only checks plumbing: argument counts/types through
ctypes, buffer shapes, state_dict handling.  The real library is never replaced on a GPU box."""
import contextlib
import ctypes as C
import os
import subprocess
import tempfile

from inpaintnet_b200 import _lib, ops


def build_stub():
    d = tempfile.mkdtemp(prefix="ipn_stub_")
    src = os.path.join(d, "stub.c")
    with open(src, "w") as f:
        f.write("#include <string.h>\n")
        for name, (res, args) in _lib.SYMBOLS.items():
            if name == "ipn_last_error":
                f.write('const char* ipn_last_error(void){return "stub";}\n')
            elif name == "ipn_struct_sizes":
                sizes = ",".join(str(C.sizeof(s)) for s in _lib.STRUCTS_IN_ORDER)
                f.write("int ipn_struct_sizes(int* o,int n){int s[]={%s};int m=%d;for(int i=0;i<n&&i<m;i++)o[i]=s[i];return m;}\n"
                        % (sizes, len(_lib.STRUCTS_IN_ORDER)))
            elif name == "ipn_launch_count":
                f.write("long long ipn_launch_count(void){return 0;}\n")
            else:
                f.write("int %s(){return 0;}\n" % name)
    so = os.path.join(d, "libstub.so")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-w", "-o", so, src])
    return so


@contextlib.contextmanager
def stubbed():
    so = build_stub()
    lib = C.CDLL(so)
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    old = (_lib._lib, ops.stream, ops.require_cuda)
    _lib._lib = lib
    ops.stream = lambda: 0
    ops.require_cuda = lambda t, what="tensor": None
    try:
        yield
    finally:
        _lib._lib, ops.stream, ops.require_cuda = old
