// model.h -- host-side acoustic model tables (parsed from a SoundSwallower model
// directory) and their packed device image.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ssb200.h"

namespace ssb {

void set_error(const char *fmt, ...);
const char *last_error();
// kernels launched by this thread since the last reset (what ssb_batch_n_launches reports)
void note_launch();
int launch_count(bool reset);

// Integer log domain helpers (ref: src/logmath.c:283-302).  shift-0 table-less use only.
struct LogMath {
    double base = 1.0001;
    double inv_log_base = 0;  // 1/ln(base)
    explicit LogMath(double b = 1.0001);
    int32_t zero(int shift) const { return INT32_MIN >> (shift + 2); }
    int32_t log(double p, int shift) const;
    int32_t ln_to_log(double lnp, int shift) const;
    // 8-bit add table for shift `shift` (ref: src/logmath.c:61-163); false if it needs >8 bits
    bool add_table8(int shift, uint8_t out[256]) const;
};

struct HostModel {
    // Gaussian codebooks
    int32_t n_mgau = 0, n_feat = 0, n_density = 0;
    int32_t featlen[SSB_MAX_FEAT] = {0, 0, 0, 0};
    int32_t featoff[SSB_MAX_FEAT + 1] = {0, 0, 0, 0, 0};
    int32_t blk = 0;
    std::vector<float> mean, var, det;   // file order [mgau][feat][density][len]
    std::vector<int64_t> gau_off;        // [mgau][feat] -> offset in mean/var
    // senones
    int32_t n_sen = 0;
    std::vector<uint8_t> mixw;           // [feat][density][n_sen]
    std::vector<uint8_t> sen2cb;         // [n_sen]
    // model definition
    int32_t n_ciphone = 0, n_phone = 0, n_emit = 0, n_ci_sen = 0, n_sseq = 0, sil = -1;
    std::vector<uint16_t> sseq;          // [n_sseq][n_emit]
    std::vector<int32_t> ph_ssid, ph_tmat, ph_ci;
    std::vector<std::string> ciname;
    // triphone lookup (graph preparation, lexicon.cpp): the mdef's context tree
    // {ctx, n_down, pid | down} and the filler flag of every CI phone (ref: bin_mdef.h:92-125)
    struct CdNode {
        int16_t ctx, n_down;
        int32_t c;
    };
    std::vector<CdNode> cd_tree;
    std::vector<uint8_t> ci_filler;
    // transitions
    int32_t n_tmat = 0;
    std::vector<uint8_t> tp;             // [n_tmat][n_emit][n_emit+1]
    uint8_t lut8[256];
    ssb_config_t cfg;
    int32_t kind = SSB_SCORER_PTM;

    bool load(const std::string &dir, const ssb_config_t &cfg);
};

}  // namespace ssb
