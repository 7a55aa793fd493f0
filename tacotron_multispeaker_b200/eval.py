"""Evaluation loop with the command line of the reference's ``eval.py``
(``eval.py:31-65,68-88``): restore a checkpoint, synthesise a list of
sentences for one speaker identity, write ``<i>-identity-<id>-<text>.wav`` and
``<i>-identity-<id>.png`` under ``<checkpoint dir>/eval/eval-<step>/``.

Differences: the sentence list comes from ``--sentences_file`` (one per line;
``--use_eval_txt 1`` reads ``eval.txt`` like the reference) and defaults to two
short built-in lines rather than the reference's hard-coded news sentences;
``--ckpt_path`` may be omitted together with ``--id_num`` to use random-init
weights (there is no network to fetch a checkpoint).
"""
from __future__ import annotations

import argparse
import os
import re

from .hparams import hparams, hparams_debug_string
from .synthesizer import Synthesizer
from .tf_checkpoint import latest_checkpoint

DEFAULT_SENTENCES = ["安全是上海合作组织发展的前提", "经济与人文并重"]
_strip_re = re.compile("[A-Za-z0-9\\!\\%\\[\\]\\,\\，\\。\\…\\：\\“\\”]")   # eval.py:59


def get_output_base_path(checkpoint_path):
    base_dir = os.path.dirname(checkpoint_path)
    m = re.compile(r".*?\.ckpt\-([0-9]+)").match(checkpoint_path)
    name = "eval-%d" % int(m.group(1)) if m else "eval"
    return os.path.join(base_dir, "eval", name)


def run_eval(args, sentences):
    ckpt_path = args.ckpt_path
    if not ckpt_path and args.id_num is None:
        run_name = args.name or args.model
        log_dir = os.path.join(args.base_dir, "logs-%s-%s" % (run_name, args.description))
        print("Trying to restore saved checkpoints from {} ...".format(log_dir))
        ckpt_path = latest_checkpoint(log_dir)
        if not ckpt_path:
            raise RuntimeError("no model found")
        print("Checkpoint found: {}".format(ckpt_path))
    print(hparams_debug_string())
    synth = Synthesizer()
    synth.load(ckpt_path, id_num=args.id_num)
    base_path = get_output_base_path(ckpt_path or os.path.join(args.base_dir, "random-init"))
    os.makedirs(base_path, exist_ok=True)
    written = []
    for i, text in enumerate(sentences):
        text = _strip_re.sub("", text).strip()
        path = os.path.join(base_path, "%d-identity-%d-%s.wav" % (i, args.identity, text))
        path_alignment = os.path.join(base_path, "%d-identity-%d.png" % (i, args.identity))
        print("Synthesizing: %s" % path)
        synth.synthesize(text, args.identity, path, path_alignment)
        written.append(path)
    return written


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--base_dir", default="./logs/")
    parser.add_argument("--model", default="tacotron")
    parser.add_argument("--name", help="Name of the run. Used for logging. Defaults to model name.")
    parser.add_argument("--hparams", default="", help="Comma-separated list of name=value overrides")
    parser.add_argument("--ckpt_path", default=None, help="the model to be restored")
    parser.add_argument("--description", default=None)
    parser.add_argument("--identity", default=0, type=int, help="the person's speech to be synthesized")
    parser.add_argument("--use_eval_txt", default=0, type=int, help="append the lines of ./eval.txt")
    parser.add_argument("--sentences_file", default=None, help="one sentence per line")
    parser.add_argument("--id_num", default=None, type=int, help="random-init weights with this many speakers")
    args = parser.parse_args(argv)
    hparams.parse(args.hparams)
    sentences = list(DEFAULT_SENTENCES)
    if args.sentences_file:
        with open(args.sentences_file, "r", encoding="utf-8") as f:
            sentences = [line for line in f if line.strip()]
    if args.use_eval_txt:
        with open("eval.txt", "r", encoding="utf-8") as f:
            sentences.extend(line for line in f)
    return run_eval(args, sentences)


if __name__ == "__main__":
    main()
