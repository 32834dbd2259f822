"""TEST INFRASTRUCTURE — ctypes access to oracle/_build/libdcsim_oracle.so (the C restatement).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdcsim_oracle.so")
SUMMARY_K = 24 + 8 * 8
RNG_PHILOX, RNG_MT19937 = 0, 1
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dcsim_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "dcsim_b200.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.dcoracle_sizeof_spec.restype = C.c_size_t
        L.dcoracle_set_test_time_quantum.argtypes = [C.c_double]
        L.dcoracle_run_batch.restype = C.c_longlong
        L.dcoracle_run_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                         C.c_void_p]
        L.dcoracle_run_batch_hist.restype = C.c_longlong
        L.dcoracle_run_batch_hist.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p]
        L.dcoracle_open.restype = C.c_void_p
        L.dcoracle_open.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.dcoracle_advance.restype = C.c_int
        L.dcoracle_advance.argtypes = [C.c_void_p, C.c_uint64]
        L.dcoracle_summary.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("dcoracle_trace", "dcoracle_job_log", "dcoracle_cluster_log"):
            getattr(L, name).restype = C.c_uint32
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.dcoracle_close.argtypes = [C.c_void_p]
        L.dcoracle_rng_words.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_void_p]
        L.dcoracle_rng_samples.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int]
        _lib = L
    return _lib


def set_test_time_quantum(q: float):
    """TEST HOOK (dcsim_oracle.c g_test_time_quantum): 0 = off.  Rounds arrival / xfer_done instants up to multiples of q."""
    lib().dcoracle_set_test_time_quantum(float(q))


def run_batch(spec_bytes: bytes, n_replicas: int, base_seed: int, first_replica_id: int = 0,
              rng_kind: int = RNG_PHILOX, n_threads: int = 1):
    """-> (summaries [n, SUMMARY_K] float64, total events)."""
    out = np.zeros((n_replicas, SUMMARY_K), dtype=np.float64)
    buf = C.create_string_buffer(spec_bytes, len(spec_bytes))
    total = lib().dcoracle_run_batch(buf, len(spec_bytes), n_replicas, base_seed & (2**64 - 1), first_replica_id,
                                     rng_kind, n_threads, out.ctypes.data)
    if total < 0:
        raise ValueError("oracle rejected the spec blob")
    return out, int(total)


def run_batch_hist(spec_bytes: bytes, n_replicas: int, base_seed: int, first_replica_id: int = 0, n_threads: int = 1):
    """-> (summaries, per-replica latency histograms [n, 2, 128] uint32)."""
    out = np.zeros((n_replicas, SUMMARY_K), dtype=np.float64)
    hist = np.zeros((n_replicas, 2, 128), dtype=np.uint32)
    buf = C.create_string_buffer(spec_bytes, len(spec_bytes))
    total = lib().dcoracle_run_batch_hist(buf, len(spec_bytes), n_replicas, base_seed & (2**64 - 1), first_replica_id,
                                          RNG_PHILOX, n_threads, out.ctypes.data, hist.ctypes.data)
    if total < 0:
        raise ValueError("oracle rejected the spec blob")
    return out, hist


class OracleSim:
    """One replica with chunked advance and recorders."""

    def __init__(self, spec_bytes: bytes, seed: int, rng_kind: int = RNG_PHILOX, trace_cap: int = 0,
                 joblog_cap: int = 0, clog_cap: int = 0):
        self._buf = C.create_string_buffer(spec_bytes, len(spec_bytes))
        self._caps = (trace_cap, joblog_cap, clog_cap)
        self._h = lib().dcoracle_open(self._buf, len(spec_bytes), seed & (2**64 - 1), rng_kind, trace_cap, joblog_cap,
                                      clog_cap)
        if not self._h:
            raise ValueError("oracle rejected the spec blob")

    def advance(self, max_events: int = 0) -> bool:
        return bool(lib().dcoracle_advance(self._h, max_events))

    def summary(self):
        out = np.zeros(SUMMARY_K, dtype=np.float64)
        lib().dcoracle_summary(self._h, out.ctypes.data)
        return out

    def _records(self, fn, dtype, cap):
        arr = np.zeros(max(cap, 1), dtype=dtype)
        n = fn(self._h, arr.ctypes.data, cap)
        return arr[:n]

    def trace(self):
        dt = np.dtype([("t", "<f8"), ("seq", "<u4"), ("kind", "<u4")])
        return self._records(lib().dcoracle_trace, dt, self._caps[0])

    def job_log(self):
        return self._records(lib().dcoracle_job_log, JOB_DTYPE, self._caps[1])

    def cluster_log(self):
        return self._records(lib().dcoracle_cluster_log, CLUSTER_DTYPE, self._caps[2])

    def close(self):
        if self._h:
            lib().dcoracle_close(self._h)
            self._h = None

    def __del__(self):
        self.close()


JOB_DTYPE = np.dtype([("jid", "<u4"), ("n_gpus", "<u4"), ("ingress", "u1"), ("jtype", "u1"), ("dc", "u1"), ("_pad0", "u1"),
                      ("_pad1", "<u4"), ("size", "<f8"), ("f_used", "<f8"), ("start_s", "<f8"), ("finish_s", "<f8")], align=True)
CLUSTER_DTYPE = np.dtype([("time_s", "<f8"), ("freq", "<f8"), ("util_gpu_time", "<f8"), ("util_begin_ts", "<f8"),
                          ("acc_job_unit", "<f8"), ("power_w", "<f8"), ("energy_j", "<f8"), ("dc", "<i4"),
                          ("busy", "<i4"), ("run_total", "<i4"), ("run_inf", "<i4"), ("q_inf", "<i4"),
                          ("q_train", "<i4")], align=True)


def rng_words(rng_kind: int, seed: int, n: int):
    out = np.zeros(n, dtype=np.uint32)
    lib().dcoracle_rng_words(rng_kind, seed & (2**64 - 1), n, out.ctypes.data)
    return out


def rng_samples(rng_kind: int, seed: int, n: int, choice_n: int):
    r, e, l = (np.zeros(n) for _ in range(3))
    c = np.zeros(n, dtype=np.int32)
    lib().dcoracle_rng_samples(rng_kind, seed & (2**64 - 1), n, r.ctypes.data, e.ctypes.data, l.ctypes.data,
                               c.ctypes.data, choice_n)
    return r, e, l, c
