"""CPU: the synthetic generator is deterministic and produces the shapes SURVEY 8(d) describes."""
import numpy as np

from rdpn6d_b200 import synth


def test_make_batch_deterministic_and_shaped():
    a = synth.make_batch(3, H=32, seed=11)
    b = synth.make_batch(3, H=32, seed=11)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["depth"].shape == (3, 64, 64) and a["coor"].shape == (3, 3, 64, 64)
    assert a["region_idx"].dtype == np.uint8 and a["anchors"].shape == (3, 32, 3)
    assert a["hyp_idx"].shape == (3, 32, 3) and a["hyp_idx"].max() < 4096
    fg = a["depth"] > 0
    assert 0.03 < fg.mean() < 0.6
    assert (a["depth"][fg] > 0.3).all() and (a["depth"][fg] < 1.6).all()


def test_tile_and_models():
    a = synth.make_batch(2, H=8, seed=1, models=synth.make_models(3, 16, seed=2, n_symmetric=1))
    t = synth.tile_batch(a, 5)
    assert t["depth"].shape[0] == 5 and np.array_equal(t["depth"][2], a["depth"][0])
    assert a["anchors"].shape == (2, 16, 3)
    d = synth.make_batch(1, H=8, seed=1, dense=True)
    assert d["anchors"] is None and d["region_idx"] is None
