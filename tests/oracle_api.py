"""ctypes binding of oracle/liboracle.so (CPU restatement; test infrastructure only)."""
import ctypes as C
import os

import numpy as np


class OrcRg(C.Structure):
    _fields_ = [("sample", C.c_uint32), ("median", C.c_uint32), ("read_length", C.c_uint32),
                ("stddev", C.c_double), ("offset", C.c_int32), ("len", C.c_uint32),
                ("values", C.POINTER(C.c_double)), ("min_prob", C.c_double),
                ("lower_quantile_dist", C.c_uint32), ("upper_quantile_dist", C.c_uint32),
                ("max_load", C.c_uint32), ("min_init_del_len", C.c_uint32)]


class OrcParams(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("min_len", C.c_uint32), ("min_lr", C.c_double),
                ("min_sample_fraction", C.c_double), ("window_size", C.c_uint32),
                ("window_buffer", C.c_uint32), ("somatic", C.c_int32), ("window_wise", C.c_int32)]


class OrcCall(C.Structure):
    _fields_ = [("initial_length", C.c_uint32), ("iterations", C.c_uint32), ("deletion_length", C.c_uint32),
                ("lr", C.c_double), ("frequency", C.c_double), ("window_position", C.c_uint32),
                ("position", C.c_uint32), ("end_position", C.c_uint32), ("filter", C.c_uint32),
                ("segment", C.c_uint32)]


class Oracle:
    def __init__(self, so):
        self.lib = C.CDLL(so)
        self.lib.orc_process_histogram.restype = C.c_double
        self.lib.orc_process_histogram.argtypes = [C.POINTER(C.c_double), C.c_uint32, C.c_int32, C.c_uint32,
                                                   C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_uint32),
                                                   C.POINTER(C.c_uint32)]
        self.lib.orc_call_files.restype = C.c_int
        self.lib.orc_call_files.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.c_char_p, C.c_int, C.c_int,
                                            C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int64)]
        self.lib.orc_scan_contig.restype = C.c_int64
        self.lib.orc_scan_contig.argtypes = [C.POINTER(OrcParams), C.c_uint32, C.c_uint32, C.POINTER(OrcRg),
                                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32),
                                             C.c_uint32, C.c_uint32, C.POINTER(OrcCall), C.POINTER(C.c_uint32),
                                             C.c_int64, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]

    def unify_segments(self, calls, per_sample, mean_stddev, min_cover=0.5, output_failed=False):
        """unifyCalls (utils_popdel.h:567-654) on in-memory window calls (popdel_b200.api.CALL_DTYPE records with
        `segment` tags, per_sample [n, N, 13]). Returns (variants in the same dtype, per_sample, significant windows)."""
        n, N = len(calls), per_sample.shape[1]
        inp = (OrcCall * max(n, 1))()
        for i, c in enumerate(calls):
            inp[i] = OrcCall(initial_length=int(c["initial_length"]), iterations=int(c["iterations"]),
                             deletion_length=int(c["deletion_length"]), lr=float(c["lr"]), frequency=float(c["frequency"]),
                             window_position=int(c["window_position"]), position=int(c["position"]),
                             end_position=int(c["end_position"]), filter=int(c["filter"]), segment=int(c["segment"]))
        ps = np.ascontiguousarray(per_sample, dtype=np.uint32)
        out = (OrcCall * max(n, 1))()
        out_ps = np.zeros((max(n, 1), N, 13), dtype=np.uint32)
        out_sig = np.zeros(max(n, 1), dtype=np.uint32)
        self.lib.orc_unify_segments.restype = C.c_int64
        self.lib.orc_unify_segments.argtypes = [C.POINTER(OrcCall), C.POINTER(C.c_uint32), C.c_int64, C.c_uint32, C.c_double,
                                                C.c_double, C.c_int, C.POINTER(OrcCall), C.POINTER(C.c_uint32),
                                                C.POINTER(C.c_uint32)]
        m = self.lib.orc_unify_segments(inp, ps.ctypes.data_as(C.POINTER(C.c_uint32)), n, N, float(mean_stddev),
                                        float(min_cover), int(bool(output_failed)), out,
                                        out_ps.ctypes.data_as(C.POINTER(C.c_uint32)), out_sig.ctypes.data_as(C.POINTER(C.c_uint32)))
        res = np.zeros(m, dtype=calls.dtype)
        for i in range(m):
            for f in ("initial_length", "iterations", "deletion_length", "lr", "frequency", "window_position", "position",
                      "end_position", "filter", "segment"):
                res[f][i] = getattr(out[i], f)
        return res, out_ps[:m].copy(), out_sig[:m].copy()

    def process_histogram(self, counts, offset, median, read_length, smoothing=True, pseudo=500):
        v = np.ascontiguousarray(counts, dtype=np.float64).copy()
        lq, uq = C.c_uint32(0), C.c_uint32(0)
        mp = self.lib.orc_process_histogram(v.ctypes.data_as(C.POINTER(C.c_double)), v.size, int(offset),
                                            int(median), int(read_length), int(smoothing), int(pseudo),
                                            C.byref(lq), C.byref(uq))
        return v, mp, lq.value, uq.value

    def call_files(self, files, dump_path, window_wise=False, dump_windows=False, uncompressed=False,
                   min_init_len=0, max_load=100):
        arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        nw = C.c_int64(0)
        rc = self.lib.orc_call_files(arr, len(files), os.fsencode(dump_path), int(window_wise), int(dump_windows),
                                     int(uncompressed), int(min_init_len), int(max_load), C.byref(nw))
        if rc != 0:
            raise RuntimeError(f"orc_call_files failed ({rc})")
        return nw.value

    def scan_contig(self, params, rgs, rg_off, pos, dev, n_samples, max_calls=100000, window_sums=False, max_windows=0):
        """params: dict; rgs: list of dicts (with 'values' ndarray). Returns (calls ndarray, per_sample, n_windows)
        or, with window_sums=True, additionally an int64 array [n_windows, 1 + R, 3] of per-window
        {pos, ncalls, 0} and per read group {active count, sum dev, sum pos}."""
        p = OrcParams(**params)
        keep = []
        arr = (OrcRg * len(rgs))()
        for i, r in enumerate(rgs):
            v = np.ascontiguousarray(r["values"], dtype=np.float64)
            keep.append(v)
            arr[i] = OrcRg(r["sample"], r["median"], r["read_length"], r["stddev"], r["offset"], v.size,
                           v.ctypes.data_as(C.POINTER(C.c_double)), r["min_prob"], r["lower_quantile_dist"],
                           r["upper_quantile_dist"], r["max_load"], r["min_init_del_len"])
        rg_off = np.ascontiguousarray(rg_off, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        dev = np.ascontiguousarray(dev, dtype=np.int32)
        calls = (OrcCall * max_calls)()
        per = np.zeros(13 * n_samples * max_calls, dtype=np.uint32)
        nw = C.c_int64(0)
        wd, wcap = None, 0
        if window_sums:
            wcap = int(max_windows) * (1 + len(rgs))
            wbuf = np.zeros(3 * wcap, dtype=np.int64)
            wd = wbuf.ctypes.data_as(C.POINTER(C.c_int64))
        n = self.lib.orc_scan_contig(C.byref(p), n_samples, len(rgs), arr, rg_off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                     pos.ctypes.data_as(C.POINTER(C.c_uint32)), dev.ctypes.data_as(C.POINTER(C.c_int32)),
                                     0, 0xFFFFFFFF, calls, per.ctypes.data_as(C.POINTER(C.c_uint32)), max_calls,
                                     wd, wcap, C.byref(nw))
        if n < 0:
            raise RuntimeError("call buffer too small")
        dt = np.dtype([("initial_length", "u4"), ("iterations", "u4"), ("deletion_length", "u4"), ("_pad", "u4"),
                       ("lr", "f8"), ("frequency", "f8"), ("window_position", "u4"), ("position", "u4"),
                       ("end_position", "u4"), ("filter", "u4"), ("segment", "u4"), ("_pad2", "u4")])
        assert dt.itemsize == C.sizeof(OrcCall), (dt.itemsize, C.sizeof(OrcCall))
        out = np.frombuffer(calls, dtype=dt, count=n).copy()
        ps = per[:13 * n_samples * n].reshape(n, n_samples, 13).copy()
        if window_sums:
            return out, ps, nw.value, wbuf[:3 * nw.value * (1 + len(rgs))].reshape(nw.value, 1 + len(rgs), 3).copy()
        return out, ps, nw.value


def load(so):
    return Oracle(so)
