"""Mixture-weight formats the bundled models do not exercise: 4-bit clustered sendump
(ref: src/ptm_mgau.c:456-609, decode at :375-378) and float mixture_weights
(ref: src/ptm_mgau.c:611-692).  Pinned by tests/golden/loader_variants.json, which
tools/make_golden.py --loaders produced with the compiled reference."""
import json
import os

import numpy as np
import pytest

import model_variants as mv
from conftest import GOLDEN, model_dir


@pytest.fixture(scope="module")
def pinned():
    with open(os.path.join(GOLDEN, "loader_variants.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def variants(tmp_path_factory, oracles):
    root = str(tmp_path_factory.mktemp("variants"))
    src = model_dir("en-us")
    mixw = oracles("en-us").model_arrays()["mixw"]
    book, packed = mv.write_clustered_sendump(src, os.path.join(root, "clustered"), mixw)
    mv.write_float_mixw(src, os.path.join(root, "floatmixw"), mixw)
    return dict(root=root, book=book, packed=packed, mixw=mixw)


def test_clustered_sendump_expansion_rule(variants, pinned):
    """The numpy statement of the reference's nibble rule reproduces the reference's table."""
    want = mv.expand_clustered(variants["book"], variants["packed"], variants["mixw"].shape[2])
    assert mv.sha(want) == pinned["clustered"]["mixw"]
    # the quirk is real: selecting by senone parity would give another table
    code = variants["packed"]
    by_parity = np.stack([code & 15, code >> 4], -1).reshape(code.shape[0], code.shape[1], -1)
    assert mv.sha(variants["book"][by_parity[..., :want.shape[2]]]) != pinned["clustered"]["mixw"]


@pytest.mark.parametrize("name", ["clustered", "floatmixw"])
def test_oracle_loads_variant(variants, pinned, synthetic, name):
    from oracle.oracle import Oracle
    o = Oracle(os.path.join(variants["root"], name))
    assert mv.sha(o.model_arrays()["mixw"]) == pinned[name]["mixw"]
    assert mv.sha(o.score_all(synthetic["feat1"]).astype(np.int16)) == pinned[name]["senscr"]


@pytest.mark.parametrize("name", ["clustered", "floatmixw"])
def test_product_loads_variant(variants, pinned, name):
    import soundswallower_b200 as ssb
    m = ssb.AcousticModel(os.path.join(variants["root"], name), device=-1)
    assert mv.sha(m.arrays()["mixw"]) == pinned[name]["mixw"]
    m.close()


def test_reference_agrees_when_present(variants, pinned, synthetic):
    """In the build container the compiled reference is asked directly."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("oracle/_ref/libssref.so not built here")
    for name in ("clustered", "floatmixw"):
        r = refshim.Ref(os.path.join(variants["root"], name), compallsen=True)
        assert mv.sha(r.model_arrays()["mixw"]) == pinned[name]["mixw"]
        assert mv.sha(r.score_all(synthetic["feat1"]).astype(np.int16)) == pinned[name]["senscr"]
        r.close()


def test_bad_variant_files_are_refused(variants, tmp_path):
    import soundswallower_b200 as ssb
    src = os.path.join(variants["root"], "floatmixw")
    dst = str(tmp_path / "broken")
    mv._link_shared(model_dir("en-us"), dst)
    blob = bytearray(open(os.path.join(src, "mixture_weights"), "rb").read())
    blob[-9] ^= 0x40                                      # flip a data bit: checksum must fail
    with open(os.path.join(dst, "mixture_weights"), "wb") as fh:
        fh.write(blob)
    with pytest.raises(ssb.SsbError):
        ssb.AcousticModel(dst, device=-1)
    os.remove(os.path.join(dst, "mixture_weights"))
    cl = open(os.path.join(variants["root"], "clustered", "sendump"), "rb").read()
    with open(os.path.join(dst, "sendump"), "wb") as fh:
        fh.write(cl[:len(cl) // 2])                       # truncated rows
    with pytest.raises(ssb.SsbError):
        ssb.AcousticModel(dst, device=-1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["clustered", "floatmixw"])
def test_gpu_scores_variant(variants, pinned, synthetic, name):
    """The device image built from a variant model scores like the reference did."""
    import soundswallower_b200 as ssb
    m = ssb.AcousticModel(os.path.join(variants["root"], name), device=0)
    scr = ssb.score_batch(m, [synthetic["feat1"]])[0]
    assert mv.sha(np.ascontiguousarray(scr, np.int16)) == pinned[name]["senscr"]
    m.close()
