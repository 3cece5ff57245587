"""The decoder-level surface of the reference for the alignment path: a
`Decoder` with the methods and result types of the reference's Python binding
(ref: py/_soundswallower.pyx:251-823, 1061-1148; py/soundswallower/__init__.py:43-88) and the
JSON of decoder_result_json (ref: src/decoder.c:1339-1593), on top of the batched GPU path
-- plus batch variants (`align_batch`, `dumps_batch`), because one utterance at a time
cannot fill a GPU.

Scope: forced alignment (`set_align_text`) and grammar decoding with a flattened grammar
(`set_fsg_graph`); JSGF parsing, add_word, VAD and the config parser stay with the caller.
Both passes run in the reference's default mode (compallsen = no: only the senones of the
active HMMs are scored, and the second pass starts from the flags the first one left), so
Hyp.score, Seg.ascore and the alignment are the default CLI's, bit for bit.  Models the
active-list search does not cover (semi-continuous, continuous) fall back to compallsen = yes.
"""
import collections
import wave

import numpy as np

from . import (AcousticModel, Frontend, Lexicon, SsbError, _left_active, align_batch, fsg_batch,
               propagate)

Seg = collections.namedtuple("Seg", ["text", "start", "duration", "ascore", "lscore"])
Hyp = collections.namedtuple("Hyp", ["text", "score", "prob"])


def get_audio_data(input_file):
    """(bytes, sample rate) of a single-channel WAV, or (bytes, None) for raw audio
    (ref: py/soundswallower/__init__.py:43-64)."""
    try:
        with wave.open(input_file) as wavfile:
            if wavfile.getnchannels() != 1:
                raise ValueError("Only supporting single-channel WAV")
            return wavfile.readframes(wavfile.getnframes()), wavfile.getframerate()
    except wave.Error:
        with open(input_file, "rb") as rawfile:
            return rawfile.read(), None


class AlignmentEntry:
    """A word, phone or state of an alignment; iterating yields its children
    (ref: py/_soundswallower.pyx:1061-1101)."""
    __slots__ = ("name", "start", "duration", "score", "_children")

    def __init__(self, name, start, duration, score, children=()):
        self.name, self.start, self.duration, self.score = name, int(start), int(duration), int(score)
        self._children = children

    def __iter__(self):
        return iter(self._children)

    def __repr__(self):
        return "AlignmentEntry(%r, %d, %d, %d)" % (self.name, self.start, self.duration, self.score)


class Alignment:
    """Sub-word alignment (ref: py/_soundswallower.pyx:1103-1148): words() / phones() /
    states(), and iteration over the words."""

    def __init__(self, words):
        self._words = words

    def __iter__(self):
        return self.words()

    def words(self):
        return iter(self._words)

    def phones(self):
        return (p for w in self._words for p in w)

    def states(self):
        return (s for w in self._words for p in w for s in p)


def _json_item(b, d, p, t):
    # HYP_FORMAT (ref: src/decoder.c:1339)
    return '{"b":%.3f,"d":%.3f,"p":%.3f,"t":"%s"' % (b, d, p, t)


class _Result:
    """What the reference keeps per utterance after decoder_end_utt (+ decoder_alignment)."""

    def __init__(self, dec, n_feat_frames, graph, p1):
        self.dec, self.graph, self.p1 = dec, graph, p1
        # decoder_n_frames = acmod->output_frame + 1 (ref: src/decoder.c:1247-1250)
        self.n_frames = n_feat_frames + 1
        self.alignment = None

    @property
    def has_hyp(self):
        return self.p1 is not None and self.p1["rv"] == 0 and self.p1["exit"] > 0

    def seg_rows(self):
        """[(word, sf, ef, ascr, lscr, fsg wid)] of the first pass (fsg_search_seg_iter)."""
        if not self.has_hyp:
            return []
        g = self.graph
        out = []
        for link, sf, ef, ascr, lscr in self.p1["segs"]:
            wid = int(g["link"][link, 3])
            out.append((g["words"][wid] if wid >= 0 else "(NULL)", int(sf), int(ef), int(ascr),
                        int(lscr), wid))
        return out

    def hyp_text(self):
        """fsg_search_hyp (ref: src/fsg_search.c:945-1026): no null transitions, no fillers,
        base strings."""
        if not self.has_hyp:
            return None
        lx = self.dec.lexicon
        words = []
        for word, _sf, _ef, _a, _l, wid in self.seg_rows():
            if wid < 0 or self.graph["filler"][wid]:
                continue
            words.append(lx.wordstr(lx.basewid(int(self.graph["dict_wid"][wid]))))
        return " ".join(words) if words else None


class Decoder:
    """The reference's `Decoder` for the alignment path, one utterance at a time, plus
    `align_batch` for many.  Keyword arguments mirror the reference's config keys: hmm, dict,
    fdict, samprate, beam/pbeam/wbeam/lw/wip/pip/silprob/fillprob/maxhmmpf/fsgusefiller/
    fsgusealtpron, and any frontend key of feat_params.json."""

    _SEARCH_KEYS = ("beam", "pbeam", "wbeam", "lw", "wip", "pip", "silprob", "fillprob", "maxhmmpf",
                    "fsgusefiller", "fsgusealtpron")

    def __init__(self, hmm, dict=None, fdict=None, device=0, logbase=1.0001, **config):
        self.hmm = hmm
        self.logbase = float(logbase)
        self.search_cfg = {k: config.pop(k) for k in list(config) if k in self._SEARCH_KEYS}
        self.fe_cfg = config
        self.model = AcousticModel(hmm, device=device, logbase=logbase)
        self.lexicon = Lexicon(self.model, dictfile=dict, fdictfile=fdict, hmmdir=hmm)
        self.frontend = Frontend(hmm, device=device, **self.fe_cfg)
        self.frate = int(self.frontend.cfg.frate)
        self._graph = None
        self._pcm = None
        self._res = None

    # ---- configuration
    @property
    def samprate(self):
        return int(self.frontend.cfg.samprate)

    def reinit_feat(self, **config):
        """New frontend parameters (ref: decoder_reinit_feat), e.g. samprate=8000."""
        self.fe_cfg.update(config)
        self.frontend.close()
        self.frontend = Frontend(self.hmm, device=self.model.device, **self.fe_cfg)
        self.frate = int(self.frontend.cfg.frate)

    def lookup_word(self, word):
        """Space-separated phones of a dictionary word, or None (decoder_lookup_word)."""
        wid = self.lexicon.wordid(word)
        if wid < 0:
            return None
        return " ".join(self.model.ciname(int(ci)) for ci in self.lexicon.pron(wid))

    def set_align_text(self, text):
        """decoder_set_align_text (ref: src/decoder.c:685-735): the grammar of one transcript."""
        try:
            self._graph = self._with_fillers(self.lexicon.align_graph(text, **self.search_cfg))
        except SsbError as e:
            raise RuntimeError("Failed to set up alignment of %s (%s)" % (text, e))

    def set_fsg_graph(self, graph):
        """decoder_set_fsg for a grammar the caller has compiled and flattened
        (see ssb_fsg_graph_t; needs `words` and `dict_wid` like Lexicon.align_graph's result)."""
        self._graph = self._with_fillers(dict(graph))

    def _with_fillers(self, g):
        if "filler" not in g:
            g["filler"] = np.array([self.lexicon.is_filler(int(w)) for w in g["dict_wid"]], bool)
        return g

    # ---- one utterance
    def start_utt(self):
        if self._pcm is not None:
            raise RuntimeError("Failed to start utterance processing")
        self._pcm = []
        self._res = None

    def process_raw(self, data, no_search=False, full_utt=False):
        """16-bit signed PCM bytes (or an int16 / float32 array).  Audio is collected here and
        decoded as one utterance at end_utt(): normalisation is per utterance, as with
        full_utt=True in the reference (what decode_file and the CLI use)."""
        if self._pcm is None:
            raise RuntimeError("Failed to process audio data: no utterance started")
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, np.int16, len(data) // 2)
        self._pcm.append(np.asarray(data))

    def end_utt(self):
        if self._pcm is None:
            raise RuntimeError("Failed to stop utterance processing")
        pcm = np.concatenate(self._pcm) if self._pcm else np.zeros(0, np.int16)
        self._pcm = None
        if self._graph is None:
            raise RuntimeError("No search module is selected, did you forget to specify a grammar?")
        self._res = self._decode([pcm], [self._graph])[0]

    def _decode(self, pcms, graphs):
        feats = self.frontend.run(pcms)
        n = [int(feats.frame_off[i + 1] - feats.frame_off[i]) for i in range(len(pcms))]
        self._feats = feats
        # the reference's default mode (compallsen = no) wherever the model allows it
        p1 = fsg_batch(self.model, feats, graphs, utt_graph=np.arange(len(graphs), dtype=np.int32),
                       compallsen=not self.model.fsg_active_ok)
        return [_Result(self, n[i], graphs[i], p1[i]) for i in range(len(pcms))]

    @property
    def n_frames(self):
        return self._res.n_frames if self._res else 0

    def _exp(self, score):
        return float(self.logbase ** score)   # logmath_exp

    @property
    def hyp(self):
        r = self._res
        text = r.hyp_text() if r else None
        if text is None:
            return Hyp(text=None, score=0., prob=0.)
        return Hyp(text=text, score=self._exp(r.p1["hyp_score"]), prob=self._exp(0))

    @property
    def seg(self):
        r = self._res
        for word, sf, ef, ascr, lscr, _wid in (r.seg_rows() if r else []):
            yield Seg(text=word, start=sf / self.frate, duration=(ef + 1 - sf) / self.frate,
                      ascore=self._exp(ascr), lscore=self._exp(lscr))

    # ---- second pass
    def _second_pass(self, results, feats):
        """decoder_alignment (ref: src/decoder.c:737-798) for every result that has a
        hypothesis: the words of pass 1 (null transitions dropped) with their frame windows ->
        alignment_populate -> state_align_search -> alignment_propagate."""
        lx, m = self.lexicon, self.model
        E = m.n_emit
        empty = dict(ssid=np.zeros(0, np.int32), tmat=np.zeros(0, np.int32), sf=np.zeros(0, np.int32),
                     ef=np.zeros(0, np.int32))
        chains, metas = [], []
        for r in results:
            rows = [s for s in r.seg_rows() if s[5] >= 0]
            if not r.has_hyp or not rows:
                chains.append(empty)
                metas.append(None)
                continue
            wids = np.array([r.graph["dict_wid"][s[5]] for s in rows], np.int32)
            start = np.array([s[1] for s in rows], np.int32)
            dur = np.array([s[2] - s[1] + 1 for s in rows], np.int32)
            c = lx.populate(wids, start, dur)
            chains.append(c)
            metas.append((wids, c))
        # the aligner starts from the flags the first pass left in acmod
        # (ref: src/state_align_search.c:186-188 never clears them)
        carried = [r.p1.get("carried") for r in results]
        p2 = align_batch(m, feats, chains, init_active=_left_active([r.p1 for r in results]),
                         init_topn=carried if all(c is not None for c in carried) else None)
        sseq = m.arrays()["sseq"]
        for r, meta, a in zip(results, metas, p2):
            if meta is None or a["rv"] != 0:
                r.alignment = None
                continue
            wids, c = meta
            ps, pd, pc = propagate(a["start"], a["dur"], a["score"], E)
            sen = sseq[c["ssid"]].reshape(-1)
            words = []
            for i, w in enumerate(wids):
                idx = np.nonzero(c["parent"] == i)[0]
                phones = []
                for q in idx:
                    states = [AlignmentEntry(str(int(sen[q * E + j])), a["start"][q * E + j],
                                             a["dur"][q * E + j], a["score"][q * E + j]) for j in range(E)]
                    phones.append(AlignmentEntry(m.ciname(int(c["ci"][q])), ps[q], pd[q], pc[q], states))
                words.append(AlignmentEntry(lx.wordstr(int(w)), ps[idx[0]], int(pd[idx].sum()),
                                            int(pc[idx].sum()), phones))
            r.alignment = Alignment(words)

    @property
    def alignment(self):
        """The sub-word alignment of the current hypothesis (runs the second pass)."""
        r = self._res
        if r is None or not r.has_hyp:
            return None
        if r.alignment is None:
            self._second_pass([r], self._feats)
        return r.alignment

    # ---- results as JSON (decoder_result_json, ref: src/decoder.c:1494-1593)
    def _dumps(self, r, start_time, align_level):
        frate = self.frate
        text = r.hyp_text() or ""
        out = [_json_item(start_time, r.n_frames / frate, self._exp(0), text), ',"w":[']
        items = []
        if align_level:
            if r.alignment is None:
                return None
            for w in r.alignment.words():
                s = [_json_item(start_time + w.start / frate, w.duration / frate, self._exp(w.score), w.name),
                     ',"w":[']
                ph = []
                for p in w:
                    q = _json_item(start_time + p.start / frate, p.duration / frate, self._exp(p.score), p.name)
                    if align_level > 1:
                        q += ',"w":[' + ",".join(
                            _json_item(start_time + st.start / frate, st.duration / frate,
                                       self._exp(st.score), st.name) + "}" for st in p) + "]"
                    ph.append(q + "}")
                items.append("".join(s) + ",".join(ph) + "]}")
        else:
            for word, sf, ef, ascr, lscr, _wid in r.seg_rows():
                items.append(_json_item(start_time + sf / frate, (ef + 1 - sf) / frate,
                                        self._exp(ascr + lscr), word) + "}")
        return "".join(out) + ",".join(items) + "]}\n"

    def dumps(self, start_time=0., align_level=0):
        """The decoding result as the reference's JSON line: align_level 0 = the first pass'
        word segmentation, 1 = words > phones from the second pass, 2 = + states."""
        r = self._res
        if r is None:
            raise RuntimeError("no utterance has been decoded")
        if align_level and r.has_hyp and r.alignment is None:
            self._second_pass([r], self._feats)
        return self._dumps(r, start_time, align_level)

    def decode_file(self, input_file):
        """(text, segmentation) of a single-channel WAV or raw file
        (ref: py/_soundswallower.pyx:734-772)."""
        data, sample_rate = get_audio_data(input_file)
        if sample_rate is not None and sample_rate != self.samprate:
            self.reinit_feat(samprate=sample_rate)
        self.start_utt()
        self.process_raw(data, no_search=False, full_utt=True)
        self.end_utt()
        if self.hyp.text is None:
            raise RuntimeError("Decoding produced no segments, "
                               "please examine dictionary/grammar and input audio.")
        return self.hyp.text, self.seg

    # ---- many utterances (what the GPU is for)
    def align_batch(self, pcms, texts, align_level=1):
        """Forced alignment of a batch of (audio, transcript) pairs: frontend, first pass,
        second pass, each one batched call.  Returns per utterance None (the transcript does
        not match the audio) or dict(text, n_frames, seg=[Seg], alignment=Alignment|None)."""
        graphs = []
        for t in texts:
            try:
                graphs.append(self._with_fillers(self.lexicon.align_graph(t, **self.search_cfg)))
            except SsbError as e:
                raise RuntimeError("Failed to set up alignment of %s (%s)" % (t, e))
        pcms = [np.frombuffer(p, np.int16, len(p) // 2) if isinstance(p, (bytes, bytearray)) else
                np.asarray(p) for p in pcms]
        results = self._decode(pcms, graphs)
        if align_level:
            self._second_pass(results, self._feats)
        self._batch = results
        out = []
        for r in results:
            if not r.has_hyp:
                out.append(None)
                continue
            seg = [Seg(text=w, start=sf / self.frate, duration=(ef + 1 - sf) / self.frate,
                       ascore=self._exp(a), lscore=self._exp(l)) for w, sf, ef, a, l, _ in r.seg_rows()]
            out.append(dict(text=r.hyp_text(), n_frames=r.n_frames, seg=seg, alignment=r.alignment))
        return out

    def dumps_batch(self, start_time=0., align_level=1):
        """decoder_result_json of every utterance of the last align_batch (None where the
        transcript did not match)."""
        return [self._dumps(r, start_time, align_level) if r.has_hyp else None for r in self._batch]

    def close(self):
        for o in (self.frontend, self.lexicon, self.model):
            try:
                o.close()
            except Exception:
                pass
