"""CPU tests of the host C side: PNG reader and PPM writer (host/hostimage.c) and the drivers'
argument handling (host/driver.c), which mirrors the reference's main() (stereo.cu:350-398)."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

from util import FIXTURES, IMGS, ROOT, load_pair


class Image8(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("width", C.c_int), ("height", C.c_int)]


@pytest.fixture(scope="module")
def himg():
    subprocess.check_call(["make", "-s", "-C", ROOT, "host/libhostimage.so"])
    L = C.CDLL(os.path.join(ROOT, "host", "libhostimage.so"))
    L.read_image.argtypes = [C.c_char_p, C.POINTER(Image8)]
    L.make_filename.restype = C.c_void_p
    L.make_filename.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.write_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return L


@pytest.fixture(scope="module")
def drivers():
    subprocess.check_call(["make", "-s", "-C", ROOT, "build=debug"])
    return os.path.join(ROOT, "debug", "stereopar"), os.path.join(ROOT, "debug", "stereopar-ghost")


def _read(himg, path):
    im = Image8()
    rc = himg.read_image(path.encode(), C.byref(im))
    if rc:
        return rc, None
    a = np.ctypeslib.as_array(im.data, (im.height, im.width)).copy()
    C.CDLL(None).free(im.data)
    return 0, a


@pytest.mark.parametrize("name", FIXTURES)
def test_png_reader_matches_pillow_on_the_fixtures(himg, name):
    a, b = load_pair(name)
    for f, ref in (("a.png", a), ("b.png", b)):
        rc, got = _read(himg, os.path.join(IMGS, name, f))
        assert rc == 0 and np.array_equal(got, ref)


def _png(path, arr, mode, bits=8):
    if mode == "L" and bits == 8:
        Image.fromarray(arr.astype(np.uint8), "L").save(path)
    elif mode == "I16":
        Image.fromarray(arr.astype(np.uint16)).save(path)
    else:
        Image.fromarray(arr, mode).save(path)


def test_png_reader_other_layouts(himg, tmp_path):
    rng = np.random.default_rng(3)
    g = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    p = str(tmp_path / "g.png")
    Image.fromarray(g, "L").save(p, compress_level=9)  # exercises all PNG filters through the optimiser
    rc, got = _read(himg, p)
    assert rc == 0 and np.array_equal(got, g)
    # 16-bit gray: stb_image keeps the high byte
    g16 = rng.integers(0, 65536, (9, 11)).astype(np.uint16)
    p = str(tmp_path / "g16.png")
    Image.fromarray(g16).save(p)
    rc, got = _read(himg, p)
    assert rc == 0 and np.array_equal(got, (g16 >> 8).astype(np.uint8))
    # 1-bit gray is scaled to 0 / 255
    g1 = rng.integers(0, 2, (8, 19)).astype(np.uint8) * 255
    p = str(tmp_path / "g1.png")
    Image.fromarray(g1, "L").convert("1").save(p)
    rc, got = _read(himg, p)
    assert rc == 0 and np.array_equal(got, g1)


def test_png_reader_rejects_what_the_reference_rejects(himg, tmp_path, capfd):
    rgb = np.zeros((4, 5, 3), np.uint8)
    p = str(tmp_path / "rgb.png")
    Image.fromarray(rgb, "RGB").save(p)
    rc, _ = _read(himg, p)
    err = capfd.readouterr().err
    assert rc == 1
    # verbatim message of image.c:27-31, without a newline
    assert err == "error reading image %s: wrong number of channels (3) (image must be grayscale)" % p
    la = np.zeros((4, 5, 2), np.uint8)
    p2 = str(tmp_path / "la.png")
    Image.fromarray(la, "LA").save(p2)
    assert _read(himg, p2)[0] == 1 and "(2)" in capfd.readouterr().err
    assert _read(himg, str(tmp_path / "missing.png"))[0] == 1
    assert "error reading image" in capfd.readouterr().err
    junk = tmp_path / "junk.png"
    junk.write_bytes(b"not a png at all")
    assert _read(himg, str(junk))[0] == 1


def _ppm_text(a, binary):
    """The reference's writer (image.c:37-47,71-88) restated in Python."""
    h, w = a.shape
    if binary:
        v = np.where(a == 1, 0, 255)
    else:
        mn, mx = int(a.min()), int(a.max())
        v = (a.astype(np.int64) - mn) * 255 // (mx - mn) if mx != mn else np.zeros_like(a, np.int64)
    body = "".join("%d %d %d\n" % (x, x, x) for x in v.ravel())
    return "P3\n%d %d\n255\n" % (w, h) + body


def test_ppm_writer_is_byte_identical(himg, tmp_path):
    rng = np.random.default_rng(5)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        b = rng.integers(0, 2, (13, 17), dtype=np.uint8)
        himg.write_image(b.ctypes.data_as(C.c_void_p), 17, 13, 0, himg.make_filename(b"edges", 1, 1))
        assert open("edges-1.ppm").read() == _ppm_text(b, True)
        s = rng.integers(0, 442, (13, 17)).astype(np.int32)
        himg.write_image(s.ctypes.data_as(C.c_void_p), 17, 13, 1, himg.make_filename(b"scores", 1, 29))
        assert open("scores-29.ppm").read() == _ppm_text(s, False)
        web = rng.integers(1, 31, (13, 17)).astype(np.int32)
        himg.write_image(web.ctypes.data_as(C.c_void_p), 17, 13, 1, himg.make_filename(b"web", 3, 2))
        assert open("web-2.ppm").read() == _ppm_text(web, False)
    finally:
        os.chdir(cwd)


def test_driver_argument_handling(drivers, tmp_path):
    a = os.path.join(IMGS, "1-240x135", "a.png")
    b = os.path.join(IMGS, "1-240x135", "b.png")
    b2 = os.path.join(IMGS, "2-480x270", "b.png")
    for exe in drivers:
        def run(*args):
            r = subprocess.run([exe, *args], capture_output=True, text=True, cwd=tmp_path)
            return r.returncode, r.stderr
        rc, err = run()
        assert rc == 1 and err.startswith("usage: stereomatch [image 1] [image 2] [threshold = 0.15] "
                                          "[square_width = 21] [times = 32] [lines = 10]")
        assert run(a, b2) == (1, "error: the two images must have equal width and height\n")
        assert run(a, b, "x") == (1, "error: threshold must be a number\n")
        assert run(a, b, "0.15", "y") == (1, "error: square_width must be a number\n")
        assert run(a, b, "0.15", "21", "z") == (1, "error: times must be a number\n")
        assert run(a, b, "0.15", "21", "32", "w") == (1, "error: lines must be a number\n")
        assert run(a, b, "1.5") == (1, "error: threshold must be between 0 and 1\n")
        assert run(a, b, "0.15", "136") == (1, "error: square width must not be higher than image width/height\n")
        rc, err = run(a, "nope.png")
        assert rc == 1 and err.startswith("error reading image nope.png:")


def test_stereobatch_argument_handling():
    """host/batch.c: usage and argument errors come before any GPU work (exit 1, message on stderr)."""
    exe = os.path.join(ROOT, "timing", "stereobatch")
    if not os.path.exists(exe):
        pytest.skip("timing/stereobatch not built (make build=timing)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("usage: ")
    r = subprocess.run([exe, "64", "64", "16"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("usage: ")
