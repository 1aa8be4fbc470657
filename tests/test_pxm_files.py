"""On-disk formats of the stereo applications (extra/SavePPM.h:20-39, stereo/main.cpp:400-410): byte-exact headers and
payloads, and a round trip.  CPU only."""
import numpy as np

from kangaroo_b200 import pxm


def test_save_pxm_bytes(tmp_path):
    img = np.arange(12, dtype=np.uint8).reshape(3, 4)
    p = tmp_path / "a.pgm"
    pxm.SavePXM(str(p), img)
    assert p.read_bytes() == b"P5\n4 3\n255\n" + bytes(range(12))
    assert np.array_equal(pxm.LoadPXM(str(p)), img)
    f = np.linspace(0, 1, 6, dtype=np.float32).reshape(2, 3)
    pxm.SavePXM(str(p), f, "P7", 65535)
    raw = p.read_bytes()
    assert raw.startswith(b"P7\n3 2\n65535\n") and raw[len(b"P7\n3 2\n65535\n"):] == f.tobytes()


def test_save_pdm_bytes_and_round_trip(tmp_path):
    d = np.array([[1.5, np.nan, 3.25], [0.0, -1.0, 127.75]], np.float32)
    p = tmp_path / "SDepth-00001.pdm"
    pxm.SavePDM(str(p), d)
    raw = p.read_bytes()
    hdr = b"P7\n3 2\n4294967295\n"
    assert raw[:len(hdr)] == hdr and raw[len(hdr):] == d.tobytes()
    back = pxm.LoadPXM(str(p))
    assert back.dtype == np.float32 and np.array_equal(np.isnan(back), np.isnan(d))
    assert np.array_equal(back[~np.isnan(d)], d[~np.isnan(d)])


def test_save_volume_pxm(tmp_path):
    v = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)   # (d, h, w)
    p = tmp_path / "v.pxm"
    pxm.SaveVolumePXM(str(p), v)
    assert p.read_bytes() == b"P5\n4 3 2\n255\n" + v.tobytes()
    assert np.array_equal(pxm.LoadPXM(str(p)), v)
