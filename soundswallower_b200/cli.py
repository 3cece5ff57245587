#!/usr/bin/env python3
"""Command-line interface, the reference's `soundswallower` command on the B200 path
(ref: py/soundswallower/cli.py:33-171): audio files in, one JSON line of time alignments per file
out (standard output, or --output).

  python -m soundswallower_b200 --align input.txt input.wav [more.wav ...]
  python -m soundswallower_b200 --align-text "go forward ten meters" --phone-align input.wav
  python -m soundswallower_b200 --fsg grammar.fsg input.wav
  python -m soundswallower_b200 --model fr-fr ...          (or --model /path/to/model/)
  python -m soundswallower_b200 --dict /path/to/dictionary.dict ...

New here -- what the GPU is for -- a whole list in one batched call (frontend, first pass, second
pass and the JSON of every utterance each run once for the list):

  python -m soundswallower_b200 --align-batch list.tsv [--phone-align]

with one `audio-file <TAB> transcript` (or `audio-file <TAB> @transcript-file`) per line.  Every
file gets its own alignment grammar; a line whose transcript does not match its audio yields
`null`.  The output lines are exactly those the reference CLI prints for the same files, one at a
time (tests/test_cli.py).

JSGF grammars (--grammar) are not compiled here (a compiled grammar can be
given as an FSG file)."""
import argparse
import os
import sys

from . import MODEL_DIR, read_fsg_file
from .decoder import Decoder, get_audio_data


def make_argparse():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter,
                                     prog="python -m soundswallower_b200")
    parser.add_argument("inputs", nargs="*", help="Input files.")
    parser.add_argument("--dict", help="Custom dictionary file.")
    parser.add_argument("--model", help="Specific model, built-in or from directory.", default="en-us")
    parser.add_argument("-s", "--set", action="append", help="Set search parameter (KEY=VALUE).")
    parser.add_argument("-o", "--output", help="Filename for output (default is standard output)")
    parser.add_argument("--device", type=int, default=0, help="CUDA device.")
    parser.add_argument("--phone-align", help="Produce phone-level alignments", action="store_true")
    grammars = parser.add_mutually_exclusive_group()
    grammars.add_argument("-a", "--align", help="Input text file for force alignment.")
    grammars.add_argument("-t", "--align-text", help="Input text for force alignment.")
    grammars.add_argument("-f", "--fsg", help="FSG file for recognition.")
    grammars.add_argument("-g", "--grammar", help="JSGF grammar file (not supported: compile it to an FSG file).")
    grammars.add_argument("-b", "--align-batch", help="TSV list: audio file <TAB> transcript (or @file) per line.")
    return parser


def model_path(name):
    """Built-in model name or a directory (ref: py/soundswallower/cli.py:96-100)."""
    if os.path.isdir(MODEL_DIR) and name in os.listdir(MODEL_DIR):
        return os.path.join(MODEL_DIR, name)
    return name


def read_batch_list(path):
    files, texts = [], []
    base = os.path.dirname(os.path.abspath(path))
    for n, ln in enumerate(open(path, encoding="utf-8"), 1):
        ln = ln.rstrip("\n")
        if not ln.strip() or ln.lstrip().startswith("#"):
            continue
        if "\t" not in ln:
            raise SystemExit("%s:%d: expected `audio-file<TAB>transcript`" % (path, n))
        f, t = ln.split("\t", 1)
        f, t = f.strip(), t.strip()
        if not os.path.isabs(f):
            f = os.path.join(base, f)
        if t.startswith("@"):
            tf = t[1:] if os.path.isabs(t[1:]) else os.path.join(base, t[1:])
            t = open(tf, encoding="utf-8").read().strip()
        files.append(f)
        texts.append(t)
    return files, texts


def main(argv=None):
    args = make_argparse().parse_args(argv)
    if args.grammar:
        raise SystemExit("--grammar: JSGF is not compiled here; give the compiled grammar with --fsg")
    if args.align:
        with open(args.align, encoding="utf-8") as fh:
            args.align_text = fh.read().strip()
    if not (args.align_text or args.fsg or args.align_batch):
        return  # nothing to do (as the reference)
    cfg = {}
    for kv in args.set or []:
        key, value = kv.split("=", 1)
        cfg[key] = float(value) if any(c in value for c in ".e") else int(value)
    decoder = Decoder(model_path(args.model), dict=args.dict, device=args.device, **cfg)
    results = []
    if args.align_batch:
        files, texts = read_batch_list(args.align_batch)
        pcms = []
        for f in files:
            data, rate = get_audio_data(f)
            if rate is not None and rate != decoder.samprate:
                raise SystemExit("%s: %d Hz; a batch shares one front end (%d Hz)" % (f, rate, decoder.samprate))
            pcms.append(data)
        decoder.align_batch(pcms, texts, align_level=1 if args.phone_align else 0)
        for js in decoder.dumps_batch(align_level=1 if args.phone_align else 0):
            results.append(js if js is not None else "null\n")
    else:
        if args.align_text is not None:
            decoder.set_align_text(args.align_text)
        else:
            n_state, start, final, trans = read_fsg_file(args.fsg)
            decoder.set_fsg_graph(decoder.lexicon.fsg_graph(n_state, start, final, trans, **decoder.search_cfg))
        for input_file in args.inputs:
            decoder.decode_file(input_file)
            results.append(decoder.dumps(align_level=1 if args.phone_align else 0))
    if args.output is not None:
        with open(args.output, "w", encoding="utf-8") as outfh:
            for json_line in results:
                outfh.write(json_line)
    else:
        for json_line in results:
            sys.stdout.write(json_line)
    decoder.close()


if __name__ == "__main__":
    main()
