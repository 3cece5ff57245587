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


def _sendump_strings(fh, strings):
    for s in strings:
        b = s.encode() + b"\0"
        fh.write(struct.pack("<i", len(b)) + b)


def write_clustered_sendump(src, dst, mixw, seed=7, n_clust=15):
    """4-bit clustered sendump.  Returns (codebook[16], packed[F][D][(S+1)//2])."""
    n_feat, n_density, n_sen = mixw.shape
    rs = np.random.RandomState(seed)
    book = np.sort(rs.choice(np.arange(0, 160), 16, replace=False)).astype(np.uint8)
    # nearest codeword per weight, two per byte (even senone in the low nibble)
    code = np.abs(mixw.astype(np.int32)[..., None] - book.astype(np.int32)).argmin(-1).astype(np.uint8)
    if n_sen & 1:
        code = np.concatenate([code, np.zeros(code.shape[:2] + (1,), np.uint8)], -1)
    packed = (code[..., 0::2] | (code[..., 1::2] << 4)).astype(np.uint8)
    _link_shared(src, dst)
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
