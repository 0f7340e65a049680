"""benchlib -- measurement-only helpers for bench.py and the sweep scripts (never imported by the product package).

libsmb_peaks.so (benchlib/peaks.cu): the INT32 issue-rate microbenchmark that supplies the roofline denominator
MEASURED_PEAKS.json does not carry, and the pinned host<->device copy peak that bounds every host-buffer figure.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _peaks():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libsmb_peaks.so")
        if not os.path.exists(path):
            raise RuntimeError("%s not found: run `make benchlib`" % path)
        _lib = C.CDLL(path)
        _lib.smb_measure_int_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.smb_measure_copy_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.smb_copy_peak_setup.argtypes = [C.c_int]
        _lib.smb_copy_peak_measure.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.smb_copy_peak_mix.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    return _lib


def measure_int_peak(device: int = 0, mode: int = 2) -> float:
    """1e9 thread-instructions/s of mode 0 IADD3, 1 LOP3, 2 IADD3+IMAD, 3 LOP3+IMAD."""
    g = C.c_double()
    if _peaks().smb_measure_int_peak(device, mode, C.byref(g)) != 0:
        raise RuntimeError("smb_measure_int_peak failed")
    return g.value


def measure_copy_peak(device: int = 0, mode: int = 2):
    """(H2D GB/s, D2H GB/s) from pinned memory; mode 0 H2D alone, 1 D2H alone, 2 both at once."""
    g = (C.c_double * 2)()
    if _peaks().smb_measure_copy_peak(device, mode, g) != 0:
        raise RuntimeError("smb_measure_copy_peak failed")
    return g[0], g[1]


def measure_copy_concurrent(device: int, barrier, mixes=(), reps: int = 6):
    """This rank's link while every other rank measures its own at the same time: buffers are set up first,
    `barrier()` lines the ranks up, then 256 MB copies are timed (mean rate, not the best repetition: contention
    is the point).  Returns (H2D GB/s, D2H GB/s) with both directions saturated, and for every (up, down) byte
    ratio in `mixes` the seconds one round of that mix takes (up + down = at most 256 MB each)."""
    L = _peaks()
    if L.smb_copy_peak_setup(device) != 0:
        raise RuntimeError("smb_copy_peak_setup failed")
    try:
        barrier()
        g = (C.c_double * 2)()
        if L.smb_copy_peak_measure(2, reps, 0, g) != 0:
            raise RuntimeError("smb_copy_peak_measure failed")
        out = []
        for up, down in mixes:
            barrier()
            unit = (256 << 20) // max(up, down)
            sec = C.c_double()
            if L.smb_copy_peak_mix(up * unit, down * unit, reps, C.byref(sec)) != 0:
                raise RuntimeError("smb_copy_peak_mix failed")
            out.append((sec.value, up * unit, down * unit))
        return g[0], g[1], out
    finally:
        L.smb_copy_peak_teardown()
