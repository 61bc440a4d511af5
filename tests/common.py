"""Shared helpers for the tests: golden fixtures, the oracle port (ctypes), parity statistics."""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def manifest():
    return json.load(open(os.path.join(GOLD, "manifest.json")))


def golden(name):
    f = os.path.join(GOLD, name + ".out.npy")
    out = np.load(f) if os.path.exists(f) else np.load(os.path.join(GOLD, name + ".out.npz"))["out"]  # large outputs are compressed
    return (np.load(os.path.join(GOLD, name + ".in.npy")), out)


def model_paths(model):
    d = os.path.join(ROOT, "models", model)
    return os.path.join(d, "x4.param"), os.path.join(d, "x4.bin")


def diff_stats(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return dict(max=int(d.max()), frac_diff=float((d > 0).mean()), n_gt1=int((d > 1).sum()), n=int(d.size))


_ORACLE = None


def oracle_lib():
    """The CPU restatement (oracle/realsr_oracle.c) as a ctypes library; built on demand with make."""
    global _ORACLE
    if _ORACLE is None:
        so = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
        lib = ctypes.CDLL(so)
        lib.rsr_oracle_load.restype = ctypes.c_void_p
        lib.rsr_oracle_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.rsr_oracle_free.argtypes = [ctypes.c_void_p]
        lib.rsr_oracle_process.restype = ctypes.c_int
        lib.rsr_oracle_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.rsr_oracle_num_conv.restype = ctypes.c_int
        lib.rsr_oracle_num_conv.argtypes = [ctypes.c_void_p]
        lib.rsr_oracle_net.restype = ctypes.c_int
        lib.rsr_oracle_net.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.rsr_oracle_conv3x3.restype = None
        lib.rsr_oracle_conv3x3.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
        _ORACLE = lib
    return _ORACLE


class Oracle:
    def __init__(self, model):
        self.lib = oracle_lib()
        p, b = model_paths(model)
        self.h = self.lib.rsr_oracle_load(p.encode(), b.encode())
        assert self.h, "oracle failed to load " + model

    def process(self, img, tilesize, tta=False, prepadding=10, threads=0):
        h, w, c = img.shape
        img = np.ascontiguousarray(img)
        out = np.zeros((h * 4, w * 4, c), dtype=np.uint8)
        r = self.lib.rsr_oracle_process(self.h, img.ctypes.data, w, h, c, out.ctypes.data, tilesize, prepadding, int(tta), threads)
        assert r == 0, "oracle process failed: %d" % r
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.rsr_oracle_free(self.h)
            self.h = None


def read_convs(model):
    """independent numpy reader of x4.param / x4.bin (formats: net.cpp:1374-1589, modelbin.cpp:88-296)"""
    p, b = model_paths(model)
    convs = []
    for line in open(p):
        t = line.split()
        if t and t[0] == "Convolution":
            kv = dict(x.split("=", 1) for x in t if "=" in x)
            convs.append(dict(name=t[1], cout=int(kv["0"]), wsize=int(kv["6"]), act=int(kv.get("9", 0))))
    buf = open(b, "rb").read()
    off = 0
    for c in convs:
        tag = int.from_bytes(buf[off:off + 4], "little")
        off += 4
        assert tag == 0x01306B47
        n = c["wsize"]
        c["cin"] = n // (c["cout"] * 9)
        c["w"] = np.frombuffer(buf, dtype=np.float16, count=n, offset=off).reshape(c["cout"], c["cin"], 3, 3)
        off += (n * 2 + 3) & ~3
        c["b"] = np.frombuffer(buf, dtype=np.float32, count=c["cout"], offset=off)
        off += 4 * c["cout"]
    assert off == len(buf)
    return convs
