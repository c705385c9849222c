"""Host mirror of the wire-format entry points (include/solvempc_b200.h): the reference's ASCII frames
(src/SerialPort.cpp:106-166) and the asynchronous per-device state feed.  Off the hot path; needs no GPU."""
import ctypes as C

import numpy as np

from . import _lib as L


def parse_frame(buf):
    """(dt, X[4]) of one frame as SerialPort::getDataFromSerial reads it, or None when readPort would reject it."""
    if isinstance(buf, str):
        buf = buf.encode()
    dt, X = C.c_double(), np.zeros(4)
    ok = L.lib().smpc_wire_parse_frame(buf, len(buf), C.byref(dt), X.ctypes.data)
    return (dt.value, X) if ok else None


def format_control(U, max_chars=0):
    """std::to_string(U) as writePort sends it; max_chars=8 reproduces the reference's sizeof(char*) truncation."""
    out = C.create_string_buffer(64)
    n = L.lib().smpc_wire_format_control(float(U), out, 64, int(max_chars))
    return out.raw[:n].decode()


class StateFeed:
    """Reader thread on a file descriptor; latest() never blocks."""

    def __init__(self, fd):
        h = C.c_void_p()
        L.check(L.lib().smpc_feed_open(C.byref(h), int(fd)))
        self._h, self._seq = h, C.c_longlong(0)

    def latest(self):
        """(dt, X[4]) of a frame newer than the last one returned, else None."""
        dt, X = C.c_double(), np.zeros(4)
        if L.lib().smpc_feed_latest(self._h, C.byref(self._seq), C.byref(dt), X.ctypes.data):
            return dt.value, X
        return None

    def stats(self):
        a, r = C.c_longlong(), C.c_longlong()
        L.check(L.lib().smpc_feed_stats(self._h, C.byref(a), C.byref(r)))
        return a.value, r.value

    def close(self):
        if getattr(self, "_h", None):
            L.lib().smpc_feed_close(self._h)
        self._h = None

    __del__ = close
