"""TEST / BENCH INFRASTRUCTURE: ctypes wrapper of oracle/cpu_ref.cpp -- the C++ multi-threaded CPU restatement of the reference's
Client.Review loop (the `cpu_baseline` / `--impl reference` arm of bench.py, kind "cpp-restatement").  Never imported by the
product package."""
import ctypes as C
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libgk_cpuref.so")


class CpuRef:
    def __init__(self):
        if not os.path.exists(LIB):
            from gatekeeper_b200 import build
            build.build()
        self.lib = C.CDLL(LIB)
        L = self.lib
        L.gk_cpuref_create.restype = C.c_void_p
        L.gk_cpuref_destroy.argtypes = [C.c_void_p]
        L.gk_cpuref_add_template.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.gk_cpuref_add_constraint.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.gk_cpuref_put_namespace.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.gk_cpuref_constraint_count.argtypes = [C.c_void_p]
        L.gk_cpuref_constraint_count.restype = C.c_uint32
        L.gk_cpuref_constraint_key.argtypes = [C.c_void_p, C.c_uint32]
        L.gk_cpuref_constraint_key.restype = C.c_char_p
        L.gk_cpuref_review_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_int,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                            C.POINTER(C.c_char_p)]
        self.h = C.c_void_p(L.gk_cpuref_create())

    def _err(self, rc, err):
        if rc != 0:
            raise RuntimeError(err.value.decode() if err.value else "cpu_ref error")

    def add_template(self, kind, rego):
        err = C.c_char_p()
        b = rego.encode()
        self._err(self.lib.gk_cpuref_add_template(self.h, kind.encode(), b, len(b), C.byref(err)), err)

    def add_constraint(self, c):
        err = C.c_char_p()
        b = json.dumps(c).encode()
        self._err(self.lib.gk_cpuref_add_constraint(self.h, b, len(b), C.byref(err)), err)

    def add_namespace(self, ns):
        b = json.dumps(ns).encode()
        self.lib.gk_cpuref_put_namespace(self.h, ns["metadata"]["name"].encode(), b, len(b))

    def constraints(self):
        n = self.lib.gk_cpuref_constraint_count(self.h)
        return [self.lib.gk_cpuref_constraint_key(self.h, i).decode() for i in range(n)]

    def review_blob(self, blob, ep, threads, source="Original"):
        """-> (pair totals per constraint key, number of results, matcher errors, seconds)"""
        keys = self.constraints()
        tot = (C.c_uint64 * max(1, len(keys)))()
        nres, nerr, secs = C.c_uint64(), C.c_uint64(), C.c_double()
        err = C.c_char_p()
        self._err(self.lib.gk_cpuref_review_blob(self.h, blob.buf, C.cast(blob.offsets, C.c_void_p), len(blob), source.encode(), ep.encode(), threads,
                                                 tot, C.byref(nres), C.byref(nerr), C.byref(secs), C.byref(err)), err)
        return {k: int(tot[i]) for i, k in enumerate(keys)}, int(nres.value), int(nerr.value), float(secs.value)

    def close(self):
        if self.h:
            self.lib.gk_cpuref_destroy(self.h)
            self.h = None
