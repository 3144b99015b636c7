"""Torch7 serialisation (next row 4): the reader/writer against the calibration files the
reference ships (tests/golden/ref_calibration.npz holds their bytes; known answers from
radial/generate_calibration_file.lua), and model-file round trips.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-estimation_b200"))
from depthmatch import torch7io as T  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "ref_calibration.npz")


def test_reads_the_reference_calibration_files_and_writes_them_back_byte_for_byte():
    z = np.load(G)
    for name in ("ardrone", "gopro", "rectified_gopro", "rectified_gopro_v2"):
        raw = z[name].tobytes()
        cal = T.loads(raw)
        assert cal["K"].dtype == np.float32 and cal["K"].shape == (3, 3) and cal["distortion"].shape == (5,)
        assert T.dumps(cal) == raw, name
    # known answers: radial/generate_calibration_file.lua:85-103 (gopro), :5-30 (ardrone)
    cal = T.loads(z["gopro"].tobytes())
    assert (cal.wImg, cal.hImg, cal.bad_image_threshold) == (1280, 720, 0.2)
    np.testing.assert_array_equal(cal.K, np.array([[602.663208, 0, 641.4552], [0, 603.193289, 344.950836],
                                                    [0, 0, 1]], np.float32))
    np.testing.assert_array_equal(cal.distortion, np.array([-0.35574, 0.142684, 0.000469, 0.000801, -0.027673],
                                                           np.float32))
    assert cal.sfm == {"max_points": 400, "points_quality": 0.001, "ransac_max_dist": 1}
    v2 = T.loads(z["rectified_gopro_v2"].tobytes())
    assert v2.rectify is True and not v2.distortion.any() and v2.sfm["trackerWinSize"] == 21
    assert T.loads(z["ardrone"].tobytes()).K[0, 0] == np.float32(293.824707)


def test_value_kinds_references_views_and_errors():
    shared = np.arange(12, dtype=np.float32).reshape(3, 4)
    obj = {"n": 1.5, "i": 7, "s": "txt", "b": False, "nil": None, "seq": [1, "two", [3.0]],
           "t": shared, "again": shared, "long": np.arange(5, dtype=np.int64), "d": np.eye(2),
           "fn": T.LuaFunction(b"\x1bLJ\x01junk", {"up": 1}), "mod": T.TorchObject("nn.Tanh", {"output": shared})}
    back = T.loads(T.dumps(obj))
    assert back["n"] == 1.5 and back["i"] == 7 and back["s"] == "txt" and back["b"] is False
    assert "nil" not in back or back["nil"] is None
    assert back["seq"] == [1, "two", [3]]
    np.testing.assert_array_equal(back["t"], shared)
    assert back["again"] is back["t"] and back["mod"].fields["output"] is back["t"]   # references kept
    assert back["long"].dtype == np.int64 and back["d"].dtype == np.float64
    assert back["fn"].dumped == b"\x1bLJ\x01junk" and back["fn"].upvalues == {"up": 1}
    assert back["mod"].typename == "nn.Tanh"
    assert T.dumps(back) == T.dumps(obj)
    # a strided view (select / narrow / transpose in Lua) shares its storage: hand-built file
    w = T._Writer()
    w.put("<ii", T.TYPE_TORCH, 1)
    w.string("V 1")
    w.string("torch.FloatTensor")
    w.put("<i", 2)
    w.put("<qq", 2, 3)          # sizes
    w.put("<qq", 1, 4)          # strides: the transpose of rows 0..2, columns 1..2 of a 3 x 4
    w.put("<q", 2)              # offset 2 (1-based)
    w.put("<ii", T.TYPE_TORCH, 2)
    w.string("V 1")
    w.string("torch.FloatStorage")
    w.put("<q", 12)
    w.out.append(shared.tobytes())
    view = T.loads(b"".join(w.out))
    np.testing.assert_array_equal(view, shared[:, 1:3].T)
    with pytest.raises(T.Torch7FormatError):
        T.loads(T.dumps(obj)[:-3])
    with pytest.raises(T.Torch7FormatError):
        T.loads(b"3\n1\n2\n")     # ascii-mode file
    bad = bytearray(b"".join(w.out))
    bad[bad.index(b"torch.FloatStorage") + 18:bad.index(b"torch.FloatStorage") + 26] = (5).to_bytes(8, "little")
    with pytest.raises(T.Torch7FormatError):
        T.loads(bytes(bad[:bad.index(b"torch.FloatStorage") + 26 + 20]))


def test_model_files_round_trip_without_a_gpu(tmp_path):
    """saveModel / loadWeightsFrom / saveNetwork / loadTesterNetwork move weights through the
    version-9 / version-1 tables (no forward pass: no GPU needed)."""
    import depthmatch as dm
    rng = np.random.default_rng(5)
    g = dm.Geometry(layers=[[3, 5, 5, 8], [4, 16, 16, 10]], maxh=17, maxw=17, maxhHR=17, maxwHR=17, maxhGT=17,
                    maxwGT=17, hImg=180, wImg=320, output_extraction_method="max")
    learning = dict(rate=0.01, rate_decay=0.1, weight_decay=0, first_image=1, delta=2, num_images=10)
    flt = dm.getFilter(g, rng)
    path = dm.saveModel(str(tmp_path), "model_of", g, learning, flt, 12, score=0.25)
    assert path.endswith("3x5x5x8_4x16x16x10-17x17-/17x17-r0.01_rd0.1_wd0/1_2_19/model_of_e000012")
    raw = T.load(path)
    assert raw["version"] == 9 and sorted(raw["weights"]) == ["layer1", "layer2"] and raw["score"] == 0.25
    assert raw["geometry"]["layers"] == [[3, 5, 5, 8], [4, 16, 16, 10]]
    other = dm.getFilter(g, np.random.default_rng(6))
    assert not np.array_equal(other.getWeights()["layer2"], flt.getWeights()["layer2"])
    dm.loadWeightsFrom(other, path)
    for k in ("layer1", "layer2"):
        np.testing.assert_array_equal(other.getWeights()[k], flt.getWeights()[k])
    T.save(str(tmp_path / "old"), {"version": 8, "weights": {}})
    with pytest.raises(dm.DepthMatchError):
        dm.loadWeightsFrom(other, str(tmp_path / "old"))
    # radial, version 1: weights and biases
    netp = dict(layers=[[3, 1, 17, 5], "tanh", [5, 17, 1, 10]], hWin=15, wInput=400, hInput=400)
    rf = dm.getRadialFilter(netp, rng)
    p = dm.saveNetwork(str(tmp_path), 3, netp, rf)
    flt2, matcher, netp2 = dm.loadTesterNetwork(p, np.random.default_rng(9))
    assert netp2["layers"] == netp["layers"] and matcher.hWin == 15
    for a, b in zip(flt2.modules, rf.modules):
        if hasattr(a, "weight"):
            np.testing.assert_array_equal(a.weight, b.weight)
            np.testing.assert_array_equal(a.bias, b.bias)
    cal = dm.loadCalibration.__wrapped__ if hasattr(dm.loadCalibration, "__wrapped__") else dm.loadCalibration
    z = np.load(G)
    (tmp_path / "gopro.cal").write_bytes(z["gopro"].tobytes())
    c = cal(str(tmp_path / "gopro.cal"))
    e2 = (float(c.K[0, 2]) * 640 / c.wImg, float(c.K[1, 2]) * 640 / c.wImg)   # SURVEY 8d, c4's epipole
    assert abs(e2[0] - 320.73) < 0.01 and abs(e2[1] - 172.48) < 0.01
