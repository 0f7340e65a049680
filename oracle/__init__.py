"""ctypes front-end of the parity oracle.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(stereomatching_b200) never imports this.

Two back-ends:
  * ``Oracle``  -- oracle/liboracle.so, the C restatement (stereo_oracle.c).
  * ``RefLib``  -- oracle/_ref/libref_{wrap,ghost}_D<N>.so, the UNMODIFIED reference
    (src/stereo.c / src/stereo-ghost.c) compiled by oracle/Makefile; its own global
    stage functions are called directly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
WRAP, GHOST = 0, 1

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def crc32(a: np.ndarray) -> str:
    """zlib CRC32 of the raw row-major bytes, as 8 hex digits (SURVEY 8c convention)."""
    return "%08x" % (zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF)


def build(ref: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Oracle:
    def __init__(self) -> None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_edges.argtypes = [_u8p, C.c_int, C.c_int, C.c_double, C.c_int, _u8p]
        L.oracle_match_plane.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.oracle_box_direct.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.oracle_box_fast.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.oracle_shift_planes.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _u8p, _i32p, _i32p]
        L.oracle_match_wta.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, _i32p, _i32p]
        L.oracle_fill_web_holes.argtypes = [_i32p, C.c_int, C.c_int, C.c_int]
        L.oracle_draw_contour_map.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, _u8p]
        L.oracle_draw_contour_map.restype = C.c_int
        L.oracle_synth_pair.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _i32p]
        L.oracle_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_crc32.restype = C.c_uint32
        self.L = L

    def edges(self, img: np.ndarray, threshold: float, variant: int) -> np.ndarray:
        h, w = img.shape
        out = np.empty((h, w), np.uint8)
        self.L.oracle_edges(np.ascontiguousarray(img), w, h, threshold, variant, out)
        return out

    def match_wta(self, le, re, num_shifts, sw, variant, direct=False):
        h, w = le.shape
        best = np.empty((h, w), np.int32)
        web = np.empty((h, w), np.int32)
        self.L.oracle_match_wta(np.ascontiguousarray(le), np.ascontiguousarray(re), w, h,
                                num_shifts, sw, variant, int(direct), best, web)
        return best, web

    def shift_planes(self, le, re, sw, i, variant, direct=False):
        h, w = le.shape
        m = np.empty((h, w), np.uint8)
        a = np.empty((h, w), np.int32)
        s = np.empty((h, w), np.int32)
        self.L.oracle_shift_planes(np.ascontiguousarray(le), np.ascontiguousarray(re), w, h, sw, i,
                                   variant, int(direct), m, a, s)
        return m, a, s

    def box(self, m, sw, variant, direct):
        h, w = m.shape
        t = np.empty((h, w), np.int32)
        (self.L.oracle_box_direct if direct else self.L.oracle_box_fast)(
            np.ascontiguousarray(m), w, h, sw, variant, t)
        return t

    def fill_web_holes(self, web, times):
        web = np.ascontiguousarray(web.copy())
        h, w = web.shape
        self.L.oracle_fill_web_holes(web, w, h, times)
        return web

    def draw_contour_map(self, web, lines):
        h, w = web.shape
        out = np.zeros((h, w), np.uint8)
        rc = self.L.oracle_draw_contour_map(np.ascontiguousarray(web), w, h, lines, out)
        return rc, out

    def synth_pair(self, seed, w, h, num_shifts):
        left = np.empty((h, w), np.uint8)
        right = np.empty((h, w), np.uint8)
        disp = np.empty((h, w), np.int32)
        self.L.oracle_synth_pair(seed, w, h, num_shifts, left, right, disp)
        return left, right, disp


def ref_available(variant: int, num_shifts: int) -> bool:
    return os.path.exists(_ref_path(variant, num_shifts))


def _ref_path(variant: int, num_shifts: int) -> str:
    return os.path.join(HERE, "_ref", "libref_%s_D%d.so" % ("ghost" if variant else "wrap", num_shifts))


class RefLib:
    """The reference's own stage functions (non-static globals of src/stereo.c and
    src/stereo-ghost.c), called on numpy buffers.  NUM_SHIFTS is baked into each .so."""

    def __init__(self, variant: int, num_shifts: int) -> None:
        self.variant, self.D = variant, num_shifts
        # RTLD_LOCAL: every libref_* exports the same global names (matches, scores, ...)
        self.L = C.CDLL(_ref_path(variant, num_shifts), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        self.L.find_all_edges.restype = None
        self.L.fillup_matches.restype = None
        self.L.fillup_scores.restype = None
        self.L.find_highest_scoring_shifts.restype = None

    # -- step 1 -------------------------------------------------------------
    def edges(self, img_u8: np.ndarray, threshold: float) -> np.ndarray:
        """find_all_edges (stereo.c:72 / stereo-ghost.c:74) on u8/256.0 doubles (image.c:13)."""
        h, w = img_u8.shape
        b = img_u8.astype(np.float64) / 256.0
        if self.variant == WRAP:
            out = np.zeros((h, w), np.uint8)
            self.L.find_all_edges(b.ctypes.data_as(C.c_void_p), w, h, C.c_double(threshold),
                                  out.ctypes.data_as(C.c_void_p))
            return out
        g = self.D  # GHOST_SIZE_EDGES == NUM_SHIFTS, stereo-ghost.c:12
        bp = np.full((h + 2, w + 2), 128.0)  # ghost_add_double(..., 1, 128.0), stereo-ghost.c:384
        bp[1:-1, 1:-1] = b
        ep = np.zeros((h + 2 * g, w + 2 * g), np.uint8)
        b0 = bp.ctypes.data + ((w + 2) + 1) * 8
        e0 = ep.ctypes.data + (g * (w + 2 * g) + g)
        self.L.find_all_edges(C.c_void_p(b0), w, h, C.c_double(threshold), C.c_void_p(e0))
        return np.ascontiguousarray(ep[g:g + h, g:g + w])

    # -- step 2 -------------------------------------------------------------
    def match_wta(self, le: np.ndarray, re: np.ndarray, sw: int, planes: bool = False):
        """fillup_matches + fillup_scores + find_highest_scoring_shifts (stereo.c:306-312)."""
        h, w = le.shape
        L, D = self.L, self.D
        best = np.zeros((h, w), np.int32)
        web = np.zeros((h, w), np.int32)
        buf = np.zeros((h, w), np.int32)
        out = {}
        if self.variant == WRAP:
            le = np.ascontiguousarray(le)
            re = np.ascontiguousarray(re)
            L.allocate_matches(w, h)
            L.allocate_scores(w, h)
            L.fillup_matches(le.ctypes.data_as(C.c_void_p), re.ctypes.data_as(C.c_void_p), w, h)
            L.fillup_scores(w, h, sw, buf.ctypes.data_as(C.c_void_p))
            L.find_highest_scoring_shifts(best.ctypes.data_as(C.c_void_p),
                                          web.ctypes.data_as(C.c_void_p), w, h)
            if planes:
                mp = (C.c_void_p * D).in_dll(L, "matches")
                sp = (C.c_void_p * D).in_dll(L, "scores")
                out["matches"] = [np.ctypeslib.as_array(C.cast(mp[i], C.POINTER(C.c_uint8)), (h, w)).copy() for i in range(D)]
                out["scores"] = [np.ctypeslib.as_array(C.cast(sp[i], C.POINTER(C.c_int32)), (h, w)).copy() for i in range(D)]
            L.free_matches()
            L.free_scores()
        else:
            g = D
            lp = np.zeros((h + 2 * g, w + 2 * g), np.uint8)
            rp = np.zeros((h + 2 * g, w + 2 * g), np.uint8)
            lp[g:g + h, g:g + w] = le
            rp[g:g + h, g:g + w] = re
            off = g * (w + 2 * g) + g
            L.allocate_matches(w, h, sw)
            L.allocate_scores(w, h)
            L.fillup_matches(C.c_void_p(lp.ctypes.data + off), C.c_void_p(rp.ctypes.data + off), w, h, sw)
            L.fillup_scores(w, h, sw, buf.ctypes.data_as(C.c_void_p))
            L.find_highest_scoring_shifts(best.ctypes.data_as(C.c_void_p),
                                          web.ctypes.data_as(C.c_void_p), w, h)
            if planes:
                sp = (C.c_void_p * D).in_dll(L, "scores")
                out["scores"] = [np.ctypeslib.as_array(C.cast(sp[i], C.POINTER(C.c_int32)), (h, w)).copy() for i in range(D)]
            L.free_matches(w, sw)
            L.free_scores()
        if planes:
            return best, web, out
        return best, web
