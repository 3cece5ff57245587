/* ssb_glue.c -- the reference-side binding of libssb200.so: the few functions a SoundSwallower
 * maintainer adds to the reference tree so that its own acmod.c / decoder.c run on the B200
 * library.  Compiled against the REFERENCE's headers (it is the only file of
 * this repository that includes them) and linked with -lssb200; it contains no algorithm.
 *
 *   ssb_ptm_mgau_init(acmod_t *)      same contract as ptm_mgau_init (ref: src/ptm_mgau.c:722-
 *                                     ; src/acmod.c:101-119): a scorer object whose first member
 *                                     is mgau_t, or NULL to decline (the caller falls through to
 *                                     the next scorer).
 *
 * oracle/Makefile target ref_ssb builds the unmodified reference sources with
 * -Dptm_mgau_init=ssb_ptm_mgau_init on acmod.c only, i.e. with the one-line change of
 * INTEGRATION.md section 1 applied at compile time, into oracle/_ref/libssref_ssb.so;
 * tests/test_link_dropin.py runs the reference's decoder through it on the GPU. */
#include <string.h>

#include <soundswallower/acmod.h>
#include <soundswallower/configuration.h>
#include <soundswallower/err.h>

#include "ssb200.h"

static int g_n_init = 0;

/* how many scorers this glue has handed to acmod (the test checks the B200 path really ran) */
int
ssb_glue_n_mgau_init(void)
{
    return g_n_init;
}

mgau_t *
ssb_ptm_mgau_init(acmod_t *acmod)
{
    ssb_config_t sc;
    ssb_model_t *sm;
    ssb_mgau_t *g;
    const char *hmm = config_str(acmod->config, "hmm");

    if (hmm == NULL)
        return NULL;
    ssb_config_defaults(&sc);
    sc.logbase = config_float(acmod->config, "logbase");
    sc.varfloor = config_float(acmod->config, "varfloor");
    sc.mixwfloor = config_float(acmod->config, "mixwfloor");
    sc.tmatfloor = config_float(acmod->config, "tmatfloor");
    sc.topn = config_int(acmod->config, "topn");
    sc.ds = config_int(acmod->config, "ds");
    sm = ssb_model_load(hmm, &sc);
    if (sm == NULL) {
        E_INFO("B200 scorer declines: %s\n", ssb_last_error());
        return NULL;
    }
    g = ssb_mgau_init(sm);
    if (g == NULL) {
        E_INFO("B200 scorer declines: %s\n", ssb_last_error());
        ssb_model_free(sm);
        return NULL;
    }
    ++g_n_init;
    /* (the model lives as long as the scorer; ssb_mgau_free releases both when the scorer
     * owns it: see ssb_mgau_own_model) */
    ssb_mgau_own_model(g, 1);
    return (mgau_t *)g;
}
