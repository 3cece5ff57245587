"""Link-level drop-in proof (scorer boundary).

oracle/_ref/libssref_ssb.so is the UNMODIFIED reference (every source compiled where it lies
under /root/reference) with the one-line change of INTEGRATION.md section 1 applied at compile
time -- acmod.c is built with -Dptm_mgau_init=ssb_ptm_mgau_init -- plus integration/ssb_glue.c,
linked against libssb200.so.  The reference's own decoder_t (decoder.c, acmod.c, fsg_search.c,
state_align_search.c, ps_alignment.c on the host) then scores every frame through our mgau_t
object's vtable on the GPU, and must reproduce its own alignments bit for bit."""
import os

import numpy as np
import pytest

from conftest import DATA, model_dir
from oracle import refshim

TEXT = {"en-us": "go forward ten meters", "fr-fr": "avance de dix mètres"}
RAW = {"en-us": "goforward.raw", "fr-fr": "goforward_fr.raw"}


def test_link_library_exports_the_glue():
    if not os.path.exists(refshim.LIB_SSB):
        pytest.skip("oracle/_ref/libssref_ssb.so not built (needs /root/reference at build time)")
    import subprocess
    syms = subprocess.run(["nm", "-D", refshim.LIB_SSB], capture_output=True, text=True).stdout
    assert " T ssb_ptm_mgau_init" in syms and " T ssb_glue_n_mgau_init" in syms
    assert " U ssb_mgau_init" in syms and " U ssb_model_load" in syms      # resolved by libssb200.so
    assert " T acmod_load_am" in syms or " t acmod_load_am" in syms or "acmod_init" in syms


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_reference_decoder_runs_on_our_scorer(golden, lang):
    if not os.path.exists(refshim.LIB_SSB):
        pytest.skip("oracle/_ref/libssref_ssb.so not built")
    import ctypes as C
    g = golden[lang]
    r = refshim.Ref(model_dir(lang), lib=refshim.LIB_SSB)
    assert r.lib.ssb_glue_n_mgau_init() >= 1          # acmod_load_am took the B200 scorer
    pcm = np.frombuffer(open(os.path.join(DATA, RAW[lang]), "rb").read(), np.int16)
    a = r.align_pcm(pcm, TEXT[lang])
    assert a["hyp_score"] == int(g["hyp_score"])       # -2761 / -4236
    assert a["n_frames"] == int(g["n_frames"])
    assert np.array_equal(a["words"], g["words"])
    assert np.array_equal(a["phones"], g["phones"])
    assert np.array_equal(a["states"], g["states"])
    maps = open("/proc/self/maps").read()
    assert "libssb200.so" in maps
    r.close()
