"""ctypes front-end of the CPU oracle (oracle/dsp_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module.  It deliberately shares no code with dsp_stuff_b200/ (the product): the method names
match dsp_stuff_b200.engine.Engine so the same graph description can be applied to both.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dsp_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_engine_create.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_void_p)]
        L.orc_engine_destroy.argtypes = [ctypes.c_void_p]
        L.orc_engine_destroy.restype = None
        L.orc_node_add.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
        L.orc_node_set_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_float]
        L.orc_node_set_enum.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_char_p]
        L.orc_node_set_taps.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        L.orc_node_get_i64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p,
                                       ctypes.POINTER(ctypes.c_int64)]
        L.orc_node_port_index.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_int32)]
        L.orc_link.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_char_p]
        L.orc_compile.argtypes = [ctypes.c_void_p]
        L.orc_reset_state.argtypes = [ctypes.c_void_p]
        L.orc_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.orc_node_process.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int64]
        L.orc_resampler_create.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]
        L.orc_resampler_destroy.argtypes = [ctypes.c_void_p]
        L.orc_resampler_destroy.restype = None
        L.orc_resampler_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_int64)]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class Oracle:
    """CPU oracle engine: same graph-building surface as the product Engine."""

    def __init__(self, channels: int, sample_rate: int = 48000, ring_granule: int = 1024, threads: int = 0):
        self._L = lib()
        self.channels = channels
        h = ctypes.c_void_p()
        self._ck(self._L.orc_engine_create(channels, sample_rate, ring_granule, threads, ctypes.byref(h)))
        self._h = h
        self._n_in = 0
        self._n_out = 0

    def _ck(self, rc):
        if rc != 0:
            raise OracleError(rc, self._L.orc_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.orc_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def add_node(self, typename: str, node_id: int):
        self._ck(self._L.orc_node_add(self._h, typename.encode(), node_id))
        if typename == "input":
            self._n_in += 1
        if typename == "output":
            self._n_out += 1

    def set_f32(self, node_id: int, field: str, value: float):
        self._ck(self._L.orc_node_set_f32(self._h, node_id, field.encode(), value))

    def set_enum(self, node_id: int, field: str, variant: str):
        self._ck(self._L.orc_node_set_enum(self._h, node_id, field.encode(), variant.encode()))

    def set_taps(self, node_id: int, taps):
        t = np.ascontiguousarray(taps, dtype=np.float64)
        self._ck(self._L.orc_node_set_taps(self._h, node_id, t.ctypes.data, t.size))

    def get_i64(self, node_id: int, key: str) -> int:
        v = ctypes.c_int64()
        self._ck(self._L.orc_node_get_i64(self._h, node_id, key.encode(), ctypes.byref(v)))
        return v.value

    def port_index(self, node_id: int, port: str, is_output: bool = False) -> int:
        v = ctypes.c_int32()
        self._ck(self._L.orc_node_port_index(self._h, node_id, port.encode(), int(is_output), ctypes.byref(v)))
        return v.value

    def link(self, src: int, out_port: str, dst: int, in_port: str):
        self._ck(self._L.orc_link(self._h, src, out_port.encode(), dst, in_port.encode()))

    def compile(self):
        self._ck(self._L.orc_compile(self._h))

    def reset_state(self):
        self._ck(self._L.orc_reset_state(self._h))

    def process(self, inputs):
        """inputs: list of [C x n] float32 arrays (one per `input` terminal) -> list of outputs."""
        if isinstance(inputs, np.ndarray):
            inputs = [inputs]
        ins = [np.ascontiguousarray(x, dtype=np.float32) for x in inputs]
        assert len(ins) == self._n_in, f"graph has {self._n_in} input terminals"
        n = ins[0].shape[1] if ins else self._n_hint
        for x in ins:
            assert x.shape == (self.channels, n)
        outs = [np.zeros((self.channels, n), dtype=np.float32) for _ in range(self._n_out)]
        ip = (ctypes.c_void_p * max(1, len(ins)))(*[x.ctypes.data for x in ins])
        op = (ctypes.c_void_p * max(1, len(outs)))(*[y.ctypes.data for y in outs])
        self._ck(self._L.orc_process(self._h, ip, op, n))
        return outs

    def process_n(self, n: int):
        """For graphs without input terminals (SignalGen sources)."""
        self._n_hint = n
        return self.process([])

    def node_process(self, node_id: int, port_inputs, n_outputs: int = 1):
        """One node on pre-averaged port buffers; None = unconnected port."""
        arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.float32) for x in port_inputs]
        n = next(a.shape[1] for a in arrs if a is not None)
        outs = [np.zeros((self.channels, n), dtype=np.float32) for _ in range(n_outputs)]
        ip = (ctypes.c_void_p * len(arrs))(*[None if a is None else a.ctypes.data for a in arrs])
        pres = (ctypes.c_uint8 * len(arrs))(*[0 if a is None else 1 for a in arrs])
        op = (ctypes.c_void_p * n_outputs)(*[y.ctypes.data for y in outs])
        self._ck(self._L.orc_node_process(self._h, node_id, ip, pres, op, n))
        return outs


class Resampler:
    """Playback-side converter of the reference (devices.rs:443-500, 550-556): 48 kHz mono -> target-rate stereo through
    dasp's Converter + Sinc<[f32; 16]> (restated from the published crates: parity UNPINNED, see dsp_oracle.cpp)."""

    def __init__(self, channels: int, target_hz: float, source_hz: float = 48000.0):
        self._L = lib()
        self.channels = channels
        h = ctypes.c_void_p()
        rc = self._L.orc_resampler_create(channels, float(source_hz), float(target_hz), ctypes.byref(h))
        if rc:
            raise OracleError(rc, self._L.orc_last_error().decode())
        self._h = h

    def process(self, mono, n_out: int):
        """-> (interleaved [C, n_out, 2] float32, consumed input samples)"""
        x = np.ascontiguousarray(mono, dtype=np.float32)
        assert x.ndim == 2 and x.shape[0] == self.channels
        out = np.zeros((self.channels, n_out, 2), dtype=np.float32)
        used = ctypes.c_int64()
        rc = self._L.orc_resampler_process(self._h, x.ctypes.data, x.shape[1], out.ctypes.data, n_out, ctypes.byref(used))
        if rc:
            raise OracleError(rc, self._L.orc_last_error().decode())
        return out, used.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.orc_resampler_destroy(self._h)
            self._h = None

    __del__ = close
