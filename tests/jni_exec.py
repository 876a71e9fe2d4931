"""Builds rafting_b200/csrc/jni/rafting_jni.c against the stand-in jni.h + JNIEnv (tests/jni_stub/) and exposes its natives to
Python: what a JVM would do with the glue, minus the JVM (there is no JDK in this image)."""
import ctypes as C
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "rafting_b200", "csrc", "jni", "rafting_jni.c")
PFX = "Java_io_lubricant_consensus_raft_gpu_NativeEngine_"


class Buf(C.Structure):                       # the fake VM's direct ByteBuffer
    _fields_ = [("addr", C.c_void_p), ("cap", C.c_int64)]


class Longs(C.Structure):                     # the fake VM's long[]
    _fields_ = [("n", C.c_int32), ("p", C.POINTER(C.c_int64))]


def buf_of_array(arr):
    b = Buf(arr.ctypes.data, arr.nbytes)
    b._keep = arr
    return b


def buf_of_struct(st):
    b = Buf(C.addressof(st), C.sizeof(st))
    b._keep = st
    return b


def build(out_dir: str):
    if shutil.which("gcc") is None:
        return None
    from rafting_b200 import _build
    _build.build(), _build.build_durable(), _build.build_ingest()
    libdir = os.path.join(ROOT, "rafting_b200")
    out = os.path.join(out_dir, "librafting_jni_exec.so")
    res = subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "jni_stub"),
                          "-I", os.path.join(ROOT, "include"), GLUE, os.path.join(ROOT, "tests", "jni_stub", "fake_env.c"),
                          "-L", libdir, "-lrafting_b200", "-lrafting_durable", "-lrafting_ingest", f"-Wl,-rpath,{libdir}", "-o", out],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    L = C.CDLL(out)
    L.fake_env.restype = C.c_void_p
    L.fake_thrown_class.restype = C.c_char_p
    L.fake_thrown_message.restype = C.c_char_p
    L.env = C.c_void_p(L.fake_env())
    return L


def fn(L, name, restype, *argtypes):
    f = getattr(L, PFX + name)
    f.restype = restype
    f.argtypes = [C.c_void_p, C.c_void_p] + list(argtypes)
    return lambda *a: f(L.env, None, *a)
