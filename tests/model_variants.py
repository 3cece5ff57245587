"""Writers for the mixture-weight file formats the bundled models do not use.

The bundled en-us / fr-fr models ship an 8-bit `sendump`.  The reference also reads a
4-bit clustered `sendump` (ref: src/ptm_mgau.c:456-609) and, without a sendump, the
float `mixture_weights` S3 file (ref: src/ptm_mgau.c:611-692).  These helpers derive such
files deterministically from a bundled model so that the loaders can be pinned:
`tools/make_golden.py --loaders` loads the variants with the compiled reference and
records the SHA-256 of the weight table it ends up with in tests/golden/loader_variants.json.
"""
import hashlib
import os
import struct

import numpy as np

SHARED = ("mdef", "means", "variances", "transition_matrices", "feat_params.json", "noisedict.txt",
          "dict.txt", "phoneset.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _link_shared(src, dst):
    os.makedirs(dst, exist_ok=True)
    for name in SHARED:
        s = os.path.join(src, name)
        if os.path.exists(s) and not os.path.exists(os.path.join(dst, name)):
            os.symlink(os.path.abspath(s), os.path.join(dst, name))


def write_frontend_variant(src, dst, **params):
    """A model directory whose feat_params.json is the bundled one updated with `params`."""
    import json
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(src, "feat_params.json")) as fh:
        d = json.load(fh)
    d.update(params)
    with open(os.path.join(dst, "feat_params.json"), "w") as fh:
        json.dump(d, fh)
    for name in SHARED + ("sendump",):
        s = os.path.join(src, name)
        if name != "feat_params.json" and os.path.exists(s) and not os.path.exists(os.path.join(dst, name)):
            os.symlink(os.path.abspath(s), os.path.join(dst, name))
    return dst


def synthetic_pcm(n, seed, samprate=16000):
    """Speech-like test signal: gated harmonics + noise, int16."""
    rs = np.random.RandomState(seed)
    t = np.arange(n) / float(samprate)
    f0 = 110 + 40 * np.sin(2 * np.pi * 0.7 * t + rs.uniform(0, 6))
    ph = 2 * np.pi * np.cumsum(f0) / samprate
    x = sum(np.sin(k * ph) / k for k in range(1, 12))
    gate = (np.sin(2 * np.pi * 1.3 * t + rs.uniform(0, 6)) > -0.2).astype(float)
    x = 6000 * x * gate / 3 + rs.normal(0, 120, n)
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def _sendump_strings(fh, strings):
    for s in strings:
        b = s.encode() + b"\0"
        fh.write(struct.pack("<i", len(b)) + b)


def write_clustered_sendump(src, dst, mixw, seed=7, n_clust=15):
    """4-bit clustered sendump.  Returns (codebook[16], packed[F][D][(S+1)//2])."""
    _link_shared(src, dst)
    return _write_clustered(dst, mixw, seed, n_clust)


def _write_clustered(dst, mixw, seed=7, n_clust=15):
    n_feat, n_density, n_sen = mixw.shape
    rs = np.random.RandomState(seed)
    book = np.sort(rs.choice(np.arange(0, 160), 16, replace=False)).astype(np.uint8)
    # nearest codeword per weight, two per byte (even senone in the low nibble)
    code = np.abs(mixw.astype(np.int32)[..., None] - book.astype(np.int32)).argmin(-1).astype(np.uint8)
    if n_sen & 1:
        code = np.concatenate([code, np.zeros(code.shape[:2] + (1,), np.uint8)], -1)
    packed = (code[..., 0::2] | (code[..., 1::2] << 4)).astype(np.uint8)
    with open(os.path.join(dst, "sendump"), "wb") as fh:
        _sendump_strings(fh, ["clustered test sendump", "header"])
        _sendump_strings(fh, ["feature_count %d" % n_feat, "mixture_count %d" % n_density,
                              "model_count %d" % n_sen, "cluster_count %d" % n_clust,
                              "cluster_bits 4"])
        fh.write(struct.pack("<i", 0))
        fh.write(book.tobytes())
        fh.write(packed.tobytes())
    return book, packed


def expand_clustered(book, packed, n_sen):
    """What the reference's scoring loop reads out of a clustered sendump
    (ref: src/ptm_mgau.c:375-378): the nibble is chosen by the low bit of the packed
    byte, both senones of a pair get the same codeword."""
    b = np.repeat(packed, 2, axis=-1)[..., :n_sen].astype(np.int32)
    return book[np.where(b & 1, b >> 4, b & 15)]


def write_float_mixw(src, dst, mixw, logbase=1.0001, seed=11, chksum=True):
    """Float `mixture_weights` (and no sendump): un-normalised weights whose logs land near
    the bundled 8-bit values, with some zeros and sub-floor entries."""
    n_feat, n_density, n_sen = mixw.shape
    rs = np.random.RandomState(seed)
    p = np.power(logbase, -(mixw.astype(np.float64) * 1024.0 + rs.uniform(0, 1024, mixw.shape)))
    p[rs.uniform(size=p.shape) < 0.01] = 0.0
    p[rs.uniform(size=p.shape) < 0.01] = 1e-9
    p *= rs.uniform(0.5, 3.0, (n_feat, 1, n_sen))           # the loader has to normalise
    data = np.ascontiguousarray(p.transpose(2, 0, 1)).astype("<f4")   # [sen][feat][density]
    words = np.concatenate([np.array([n_sen, n_feat, n_density, data.size], "<i4").view("<u4"),
                            data.view("<u4").ravel()])
    _link_shared(src, dst)
    with open(os.path.join(dst, "mixture_weights"), "wb") as fh:
        fh.write(b"s3\nversion 1.0\n" + (b"chksum0 yes\n" if chksum else b"") + b"endhdr\n")
        fh.write(struct.pack("<I", 0x11223344))
        fh.write(words.tobytes())
        if chksum:
            s = 0
            for w in words.tolist():                         # ref: s3file.c:365-397
                s = ((((s << 20) | (s >> 12)) & 0xffffffff) + w) & 0xffffffff
            fh.write(struct.pack("<I", s))
    return data


DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
FE_LENGTHS = (0, 1, 100, 409, 410, 411, 569, 570, 571, 730, 5000)
FE_CASES = [  # (tag, feat_params overrides, sample rate, input)
    ("base", {}, 16000, "goforward.raw"),
    ("fr", {}, 16000, "goforward_fr.raw"),
    ("wav8k", {}, 8000, "sense_and_sensibility_01_austen_64kb-0880.wav"),
    ("legacy", dict(transform="legacy"), 16000, "synth:20000:1"),
    ("htk_dc", dict(transform="htk", remove_dc=True), 16000, "synth:12345:2"),
    ("plain", dict(remove_noise=False, lifter=0, cmn="none"), 16000, "synth:9000:3"),
    ("f40_varnorm", dict(nfilt=40, lowerf=133.33334, upperf=6855.4976, varnorm=True), 16000, "synth:16000:4"),
    ("dbw_noround", dict(doublebw=True, lowerf=300, round_filters=False, unit_area=False), 16000, "synth:8000:5"),
    ("nfft1024_wlen", dict(nfft=1024, wlen=0.02, frate=125, alpha=0.0), 16000, "synth:7777:6"),
]


def fe_input(spec, samprate):
    """int16 samples of a frontend test case (file under tests/data, or synth:<n>:<seed>)."""
    if spec.startswith("synth:"):
        _, n, seed = spec.split(":")
        return synthetic_pcm(int(n), int(seed), samprate)
    path = os.path.join(DATA, spec)
    if spec.endswith(".wav"):
        import wave
        w = wave.open(path)
        return np.frombuffer(w.readframes(w.getnframes()), np.int16).copy()
    return np.fromfile(path, np.int16)


# ---------------------------------------------------------------------- semi-continuous models
def read_gauden(path):
    """(n_mgau, n_feat, n_density, featlen, data) of a means / variances S3 file."""
    blob = open(path, "rb").read()
    pos = blob.index(b"endhdr\n") + 7
    assert struct.unpack_from("<I", blob, pos)[0] == 0x11223344
    pos += 4
    n_mgau, n_feat, n_density = struct.unpack_from("<3i", blob, pos)
    pos += 12
    featlen = struct.unpack_from("<%di" % n_feat, blob, pos)
    pos += 4 * n_feat
    n = struct.unpack_from("<i", blob, pos)[0]
    pos += 4
    data = np.frombuffer(blob, "<f4", n, pos).reshape(n_mgau, n_density * sum(featlen))
    return n_mgau, n_feat, n_density, featlen, data


def _s3_checksum(words):
    s = 0
    for w in words.tolist():
        s = ((((s << 20) | (s >> 12)) & 0xffffffff) + w) & 0xffffffff
    return s


def write_gauden(path, arr, featlen):
    """arr [n_mgau][n_feat][n_density][L] (equal stream lengths) -> S3 file with checksum."""
    n_mgau, n_feat, n_density, L = arr.shape
    assert all(x == L for x in featlen)
    words = np.concatenate([np.array([n_mgau, n_feat, n_density] + list(featlen) + [arr.size],
                                     "<i4").view("<u4"),
                            np.ascontiguousarray(arr, "<f4").view("<u4").ravel()])
    with open(path, "wb") as fh:
        fh.write(b"s3\nversion 1.0\nchksum0 yes\nendhdr\n" + struct.pack("<I", 0x11223344))
        fh.write(words.tobytes())
        fh.write(struct.pack("<I", _s3_checksum(words)))


def write_semi_model(src, dst, n_sen, n_density=256, seed=3, clustered=False, topn_beam=None):
    """A semi-continuous (single codebook) model directory for the s2_semi scorer: the bundled
    model's mdef / transitions / feature parameters, `n_density` Gaussians per stream drawn from
    its own codebooks, and seeded mixture weights for every senone (8-bit sendump, or the 4-bit
    clustered one)."""
    import json
    rs = np.random.RandomState(seed)
    os.makedirs(dst, exist_ok=True)
    n_mgau, n_feat, nd, featlen, mean = read_gauden(os.path.join(src, "means"))
    _, _, _, _, var = read_gauden(os.path.join(src, "variances"))
    L = featlen[0]
    mean = mean.reshape(n_mgau, n_feat, nd, L)
    var = var.reshape(n_mgau, n_feat, nd, L)
    pick_c, pick_d = rs.randint(0, n_mgau, (n_feat, n_density)), rs.randint(0, nd, (n_feat, n_density))
    f_idx = np.arange(n_feat)[:, None]
    write_gauden(os.path.join(dst, "means"), mean[pick_c, f_idx, pick_d][None], featlen)
    write_gauden(os.path.join(dst, "variances"), var[pick_c, f_idx, pick_d][None], featlen)
    with open(os.path.join(src, "feat_params.json")) as fh:
        fp = json.load(fh)
    if topn_beam is not None:
        fp["topn_beam"] = topn_beam
    with open(os.path.join(dst, "feat_params.json"), "w") as fh:
        json.dump(fp, fh)
    for name in ("mdef", "transition_matrices", "noisedict.txt", "dict.txt", "phoneset.json"):
        s = os.path.join(src, name)
        if os.path.exists(s) and not os.path.exists(os.path.join(dst, name)):
            os.symlink(os.path.abspath(s), os.path.join(dst, name))
    # senones: the mdef decides how many (n_sen of the source model)
    mixw = np.minimum(159, (rs.gamma(2.0, 18.0, (n_feat, n_density, n_sen))).astype(np.int32)).astype(np.uint8)
    if clustered:
        _write_clustered(dst, mixw, seed + 1)
    else:
        with open(os.path.join(dst, "sendump"), "wb") as fh:
            _sendump_strings(fh, ["semi test sendump", "header"])
            _sendump_strings(fh, ["feature_count %d" % n_feat, "mixture_count %d" % n_density,
                                  "model_count %d" % n_sen])
            fh.write(struct.pack("<i", 0))
            fh.write(struct.pack("<2i", n_density, n_sen))
            fh.write(mixw.tobytes())
    return dict(n_sen=n_sen, n_density=n_density, n_feat=n_feat)


SEMI_CASES = [  # (tag, write_semi_model keywords, topn_beam as a list)
    ("semi256", dict(n_density=256, seed=3), None),
    ("semi4b", dict(n_density=128, seed=4, clustered=True), None),
    ("beam64", dict(n_density=64, seed=5, topn_beam="40,0,25"), [40, 0, 25]),
]


# ---------------------------------------------------------------------- continuous models
def write_cont_model(src, dst, n_sen, n_density=8, seed=9):
    """A fully continuous model directory for the ms_mgau scorer: one codebook per senone with
    `n_density` 39-dimensional Gaussians (a single feature stream: no svspec), drawn from the
    bundled model's codebooks, and seeded float mixture weights."""
    import json
    rs = np.random.RandomState(seed)
    os.makedirs(dst, exist_ok=True)
    n_mgau, n_feat, nd, featlen, mean = read_gauden(os.path.join(src, "means"))
    _, _, _, _, var = read_gauden(os.path.join(src, "variances"))
    L = featlen[0]
    mean = mean.reshape(n_mgau, n_feat, nd, L)
    var = var.reshape(n_mgau, n_feat, nd, L)
    pc = rs.randint(0, n_mgau, (n_sen, n_density, n_feat))
    pd = rs.randint(0, nd, (n_sen, n_density, n_feat))
    fi = np.arange(n_feat)
    mu = mean[pc, fi, pd].reshape(n_sen, 1, n_density, n_feat * L)
    vv = var[pc, fi, pd].reshape(n_sen, 1, n_density, n_feat * L)
    write_gauden(os.path.join(dst, "means"), mu, [n_feat * L])
    write_gauden(os.path.join(dst, "variances"), vv, [n_feat * L])
    w = rs.gamma(0.7, 1.0, (n_sen, 1, n_density)).astype(np.float32)
    w[rs.uniform(size=w.shape) < 0.05] = 0.0
    words = np.concatenate([np.array([n_sen, 1, n_density, w.size], "<i4").view("<u4"),
                            w.astype("<f4").view("<u4").ravel()])
    with open(os.path.join(dst, "mixture_weights"), "wb") as fh:
        fh.write(b"s3\nversion 1.0\nchksum0 yes\nendhdr\n" + struct.pack("<I", 0x11223344))
        fh.write(words.tobytes())
        fh.write(struct.pack("<I", _s3_checksum(words)))
    with open(os.path.join(src, "feat_params.json")) as fh:
        fp = json.load(fh)
    fp.pop("svspec", None)
    with open(os.path.join(dst, "feat_params.json"), "w") as fh:
        json.dump(fp, fh)
    for name in ("mdef", "transition_matrices", "noisedict.txt", "dict.txt", "phoneset.json"):
        s = os.path.join(src, name)
        if os.path.exists(s) and not os.path.exists(os.path.join(dst, name)):
            os.symlink(os.path.abspath(s), os.path.join(dst, name))
    return dict(n_sen=n_sen, n_density=n_density)


CONT_CASES = [  # (tag, write_cont_model keywords)
    ("cont8", dict(n_density=8, seed=9)),
    ("cont3", dict(n_density=3, seed=10)),   # fewer densities than topn: the unsorted "all" list
]


def write_five_state_model(src, dst, seed=21):
    """A model directory whose HMMs have FIVE emitting states (hmm_vit_eval_5st_lr; no bundled
    model has them): the bundled mdef with every senone sequence [a, b, c] stretched to
    [a, a, b, c, c] and `transition_matrices` replaced by seeded 5 x 6 left-to-right matrices
    (self loop, next, skip).  Gaussians and mixture weights are the bundled ones."""
    os.makedirs(dst, exist_ok=True)
    for name in ("means", "variances", "sendump", "feat_params.json", "noisedict.txt", "dict.txt",
                 "phoneset.json"):
        s = os.path.join(src, name)
        if os.path.exists(s) and not os.path.exists(os.path.join(dst, name)):
            os.symlink(os.path.abspath(s), os.path.join(dst, name))
    blob = bytearray(open(os.path.join(src, "mdef"), "rb").read())
    assert blob[:4] == b"BMDF"
    fmt_len = struct.unpack_from("<i", blob, 8)[0]
    hp = 12 + fmt_len
    h = list(struct.unpack_from("<10i", blob, hp))
    n_ci, n_phone, n_emit, n_sseq, n_cd = h[0], h[1], h[2], h[6], h[8]
    assert n_emit == 3
    p = names = hp + 40
    for _ in range(n_ci):
        p = blob.index(b"\0", p) + 1
    p = names + (((p - names) + 3) & ~3)
    p += n_cd * 8 + n_phone * 12
    assert struct.unpack_from("<i", blob, p)[0] == n_sseq * 3
    sseq = np.frombuffer(bytes(blob[p + 4:p + 4 + n_sseq * 6]), "<u2").reshape(n_sseq, 3)
    tail = bytes(blob[p + 4 + n_sseq * 6:])
    s5 = sseq[:, [0, 0, 1, 2, 2]]
    h[2] = 5
    struct.pack_into("<10i", blob, hp, *h)
    out = bytes(blob[:p]) + struct.pack("<i", n_sseq * 5) + np.ascontiguousarray(s5, "<u2").tobytes() + tail
    with open(os.path.join(dst, "mdef"), "wb") as fh:
        fh.write(out)
    # transition matrices [n_tmat][5][6], rows normalised by the loader
    n_tmat = h[5]
    rs = np.random.RandomState(seed)
    tm = np.zeros((n_tmat, 5, 6), "<f4")
    for t in range(n_tmat):
        for a in range(5):
            w = rs.uniform(0.05, 1.0, 3)
            if rs.rand() < 0.3:
                w[2] = 0.0                      # no skip arc
            w /= w.sum()
            for k, b in enumerate((a, a + 1, a + 2)):
                if b <= 5:
                    tm[t, a, b] = w[k]
    words = np.concatenate([np.array([n_tmat, 5, 6, tm.size], "<i4").view("<u4"), tm.view("<u4").ravel()])
    with open(os.path.join(dst, "transition_matrices"), "wb") as fh:
        fh.write(b"s3\nversion 1.0\nchksum0 yes\nendhdr\n" + struct.pack("<I", 0x11223344))
        fh.write(words.tobytes())
        fh.write(struct.pack("<I", _s3_checksum(words)))
    return dst
