"""TEST-ONLY ctypes access to the single-lane host build of the device core (tests/hostemu/hostemu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "hostemu")
_SO = os.path.join(_DIR, "_build", "libdcsim_hostemu.so")
_CORE = os.path.join(_HERE, "..", "distributed_cluster_gpus_b200", "csrc", "dcsim_core.cuh")
_HDR = os.path.join(_HERE, "..", "include", "dcsim_b200.h")
SUMMARY_K = 24 + 8 * 8
TRACE_DTYPE = np.dtype([("t", "<f8"), ("seq", "<u4"), ("kind", "<u4")])
_SO_PERTURBED = os.path.join(_DIR, "_build", "libdcsim_hostemu_perturbed.so")
_SO_SMALLRING = os.path.join(_DIR, "_build", "libdcsim_hostemu_smallring.so")
_SO_UNIFORM = os.path.join(_DIR, "_build", "libdcsim_hostemu_uniform.so")
_lib_smallring = None
_lib_uniform = None
_lib = None
_lib_perturbed = None


def _bind(path):
    L = C.CDLL(path)
    L.hostemu_sizeof_spec.restype = C.c_size_t
    L.hostemu_set_test_time_quantum.argtypes = [C.c_double]
    L.hostemu_run_batch.restype = C.c_longlong
    L.hostemu_run_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                                    C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


def _ensure_built():
    srcs = (os.path.join(_DIR, "hostemu.cpp"), os.path.join(_DIR, "build.sh"), _CORE, _HDR)
    for so in (_SO, _SO_PERTURBED, _SO_SMALLRING, _SO_UNIFORM):
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run([os.path.join(_DIR, "build.sh")], check=True, capture_output=True)
            return


def lib(perturbed=False, smallring=False, uniform=False):
    """perturbed=True: the conditioning probe (every 5th pow() result moved by one ulp, see hostemu.cpp);
    smallring=True: the list merge built with a one-chunk ring (its HBM fall-back paths do all the work);
    uniform=True: the event-loop skeleton of the lane-group GPU builds (replicas switched off instead of leaving)."""
    global _lib, _lib_perturbed, _lib_smallring, _lib_uniform
    if _lib is None:
        _ensure_built()
        _lib, _lib_perturbed, _lib_smallring, _lib_uniform = _bind(_SO), _bind(_SO_PERTURBED), _bind(_SO_SMALLRING), _bind(_SO_UNIFORM)
    return _lib_uniform if uniform else (_lib_smallring if smallring else (_lib_perturbed if perturbed else _lib))


def set_test_time_quantum(q):
    """TEST HOOK (dcsim_core.cuh dcsim_test_quantize): 0 = off."""
    lib().hostemu_set_test_time_quantum(float(q))
    lib(perturbed=True).hostemu_set_test_time_quantum(float(q))
    lib(smallring=True).hostemu_set_test_time_quantum(float(q))
    lib(uniform=True).hostemu_set_test_time_quantum(float(q))


def run_batch(spec_bytes, n_replicas, seed0, chunk_events=0, trace_cap=0, rec_replica=-1, job_dtype=None,
              jobs_cap=0, cluster_dtype=None, cluster_cap=0, rng_kind=0, perturbed=False, smallring=False, uniform=False):
    out = np.zeros((n_replicas, SUMMARY_K))
    buf = C.create_string_buffer(spec_bytes, len(spec_bytes))
    trace = np.zeros(max(trace_cap, 1), dtype=TRACE_DTYPE)
    jobs = np.zeros(max(jobs_cap, 1), dtype=job_dtype) if job_dtype is not None else None
    cluster = np.zeros(max(cluster_cap, 1), dtype=cluster_dtype) if cluster_dtype is not None else None
    counts = np.zeros(4, dtype=np.uint32)
    layout = np.zeros(8, dtype=np.int32)
    hist = np.zeros((n_replicas, 2, 128), dtype=np.uint32)
    total = lib(perturbed, smallring, uniform).hostemu_run_batch(buf, len(spec_bytes), n_replicas, seed0 & (2**64 - 1), chunk_events,
                                    out.ctypes.data, rec_replica,
                                    trace.ctypes.data if trace_cap else None, trace_cap,
                                    jobs.ctypes.data if jobs is not None and jobs_cap else None, jobs_cap,
                                    cluster.ctypes.data if cluster is not None and cluster_cap else None, cluster_cap,
                                    counts.ctypes.data, layout.ctypes.data, hist.ctypes.data, rng_kind)
    if total < 0:
        raise ValueError("hostemu rejected the spec blob")
    res = {"summary": out, "events": int(total), "lat_hist": hist, "trace": trace[:min(int(counts[0]), trace_cap)],
           "layout": {"total_bytes": int(layout[0]), "cap_xfer": int(layout[1]), "cap_run": int(layout[2]),
                      "cap_q_inf": int(layout[3]), "cap_q_trn": int(layout[4])}}
    if jobs is not None:
        res["jobs"] = jobs[:min(int(counts[1]), jobs_cap)]
    if cluster is not None:
        res["cluster"] = cluster[:min(int(counts[2]), cluster_cap)]
    return res
