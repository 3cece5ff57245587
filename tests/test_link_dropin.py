"""Link-level drop-in proof (scorer, search and decoder boundaries).

oracle/_ref/libssref_ssb.so is the UNMODIFIED reference (every source compiled where it lies
under /root/reference) with the changes of INTEGRATION.md applied at compile time:
  * acmod.c is built with -Dptm_mgau_init=ssb_ptm_mgau_init: acmod_load_am takes our scorer;
  * decoder.c is built with -Dfsg_search_init=ssb_glue_fsg_search_init
    -Dstate_align_search_init=ssb_glue_state_align_search_init: decoder_set_fsg (hence
    decoder_set_align_text, decoder_set_jsgf_*) and decoder_alignment create our searches;
plus integration/ssb_glue.c, linked against libssb200.so.  The reference's own decoder_t
(decoder_start_utt / decoder_process_int16 / decoder_end_utt / decoder_hyp / decoder_seg_iter /
decoder_alignment, its frontend, feature buffer, dictionary and alignment_t on the host) then
runs both passes through our objects' vtables on the GPU and must reproduce its own results
bit for bit.  The library exports every decoder_* / config_* / alignment_* symbol the Cython
and JS bindings use -- they are the reference's own."""
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
    assert " T ssb_glue_fsg_search_init" in syms and " T ssb_glue_state_align_search_init" in syms
    for sym in ("decoder_init", "decoder_start_utt", "decoder_process_int16", "decoder_end_utt", "decoder_hyp",
                "decoder_seg_iter", "decoder_alignment", "decoder_result_json", "decoder_set_align_text",
                "decoder_set_fsg", "decoder_set_jsgf_string", "config_init", "alignment_words", "seg_iter_next"):
        assert " T %s\n" % sym in syms, sym                                # the Cython / JS surface
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
    n0 = r.lib.ssb_glue_n_search_init()
    pcm = np.frombuffer(open(os.path.join(DATA, RAW[lang]), "rb").read(), np.int16)
    a = r.align_pcm(pcm, TEXT[lang])
    assert a["hyp_score"] == int(g["hyp_score"])       # -2761 / -4236
    assert a["n_frames"] == int(g["n_frames"])
    assert np.array_equal(a["words"], g["words"])
    assert np.array_equal(a["phones"], g["phones"])
    assert np.array_equal(a["states"], g["states"])
    assert r.lib.ssb_glue_n_search_init() >= n0 + 2    # decoder_set_fsg + decoder_alignment: our searches
    maps = open("/proc/self/maps").read()
    assert "libssb200.so" in maps
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["en-us", "fr-fr"])
def test_reference_decoder_asks_our_search_between_steps(golden, lang):
    """decoder_hyp / decoder_seg_iter of the reference's decoder_t between two
    search_module_step calls, served by our grammar search through the glue: the unmodified
    reference's partial hypotheses (tests/golden/fsg_partial.npz)."""
    if not os.path.exists(refshim.LIB_SSB):
        pytest.skip("oracle/_ref/libssref_ssb.so not built")
    pg = np.load(os.path.join(os.path.dirname(__file__), "golden", "fsg_partial.npz"))
    r = refshim.Ref(model_dir(lang), lib=refshim.LIB_SSB)
    stops = pg["%s_stops" % lang].tolist()
    res = r.fsg_partial(golden[lang]["feat"], TEXT[lang], stops)
    for k, x in enumerate(res):
        assert (x["hyp"] or "") == str(pg["%s_hyp" % lang][k]), stops[k]
        if x["hyp"]:
            assert x["hyp_score"] == int(pg["%s_score" % lang][k]), stops[k]
        n = int(pg["%s_nseg" % lang][k])
        assert np.array_equal(x["segs"], pg["%s_segs" % lang][k, :n]), stops[k]
    r.close()
