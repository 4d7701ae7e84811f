"""What comes out of the decoder: the n-best result file, language-model rescoring, WER.

* `decode_to_file` -- the loop of L/decode.py:147-161: beam-search every batch of a loader (`translate_batch`, on the
  GPU) and write one line per hypothesis, `key \\t score \\t words` (BOS/EOS stripped, unknown indices -> <unk>).
* `rescore` -- L/rescore.py:12-64: per utterance pick `argmax(model_score + lm_score / w)` for every inverse weight w
  and write `rescore_<w>` files of `key words` lines.  Pinned by `tests/golden/rescore_small.json`, produced by running
  the reference script itself.
* `compute_wer` / `best_wer` -- the reference shells out to Kaldi's `compute-wer --mode=present` and `best_wer.sh`
  (P/run.sh:192-203); Kaldi is not part of the reference tree, so this is a restatement of its word-level Levenshtein
  scoring with the same report lines.  PARITY UNPINNED for the ins/del/sub split (no Kaldi binary here); the total
  error count of a minimum edit distance is unique.
"""
from __future__ import annotations

import os
import re
from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from .utils import constants, instances_handler

__all__ = ["write_nbest", "decode_to_file", "read_nbest", "rescore", "edit_counts", "compute_wer", "format_wer", "best_wer"]


# ------------------------------------------------------------------------------------------------ n-best file
def write_nbest(f, keys, all_hyp, all_scores, idx2word) -> int:
    """One line per (utterance, hypothesis): key, tab, score, tab, words (L/decode.py:154-161).  Hypotheses carry BOS
    and EOS, which are dropped; `zip` pairs the <= nbest hypotheses with the first scores of the beam.  -> lines written."""
    n = 0
    for key, hyps, scores in zip(keys, all_hyp, all_scores):
        for tokens, score in zip(hyps, scores):
            words = [idx2word.get(int(i), constants.UNK_WORD) for i in tokens[1:-1]]
            f.write(key + '\t' + str(score) + '\t' + ' '.join(words) + '\n')
            n += 1
    return n


def decode_to_file(model, decode_data, opt, model_options, read_vocab_file, save_result_file) -> int:
    """Decode every batch of `decode_data` (a BatchLoader in mode 'all') and write the n-best file.  `opt` carries
    beam_size, max_token_seq_len, nbest like L/decode.py:110-122.  -> number of utterances decoded.
    Data-parallel decoding: give each rank a loader built with `shard=(rank, world)` and its own result file."""
    from .decode import translate_batch
    if opt.nbest > opt.beam_size:
        raise ValueError('[ERROR] nbest should not larger than beam_size')
    word2idx = read_vocab_file if isinstance(read_vocab_file, dict) else instances_handler.read_vocab(read_vocab_file)
    idx2word = {index: word for word, index in word2idx.items()}
    decode_data.mode = 'all'
    n_utts = 0
    with open(save_result_file, 'w', encoding='utf-8') as f:
        for batch in decode_data:
            all_hyp, all_scores = translate_batch(model, batch, opt, model_options)
            write_nbest(f, batch[0], all_hyp, all_scores, idx2word)
            n_utts += len(batch[0])
    return n_utts


def read_nbest(decode_file) -> "OrderedDict[str, Tuple[List[float], List[str]]]":
    """n-best file -> {key: (scores, results)} in first-appearance order."""
    table: "OrderedDict[str, Tuple[List[float], List[str]]]" = OrderedDict()
    with open(decode_file, encoding='utf-8') as f:
        for n, line in enumerate(f, 1):
            parts = line.rstrip('\n').split('\t')
            if len(parts) != 3:
                raise ValueError('[ERROR] {} line {}: expected key<TAB>score<TAB>words'.format(decode_file, n))
            scores, results = table.setdefault(parts[0], ([], []))
            scores.append(float(parts[1].strip()))
            results.append(parts[2].strip())
    return table


# ------------------------------------------------------------------------------------------------ rescoring
def rescore(decode_file, lm_score, save_dir, inv_weight_list) -> List[str]:
    """For every inverse LM weight w write `<save_dir>/rescore_<w>` with, per utterance, the hypothesis maximising
    model_score + lm_score / w (first one on ties, like numpy argmax).  `lm_score` holds one score per line of
    `decode_file`, in the same order.  `inv_weight_list`: '10,11,13.5' or a sequence of numbers.  -> files written."""
    if isinstance(inv_weight_list, str):
        weights = [float(w) for w in inv_weight_list.split(',')]
    else:
        weights = [float(w) for w in inv_weight_list]
    table = read_nbest(decode_file)
    with open(lm_score, encoding='utf-8') as f:
        lm_all = [float(line.strip()) for line in f if line.strip()]
    n_lines = sum(len(scores) for scores, _ in table.values())
    if len(lm_all) < n_lines:
        raise ValueError('[ERROR] {} has {} scores for {} hypotheses'.format(lm_score, len(lm_all), n_lines))
    # hypotheses of one key need not be contiguous in the file: walk it again to pair line numbers with keys
    lm_of: Dict[str, List[float]] = {key: [] for key in table}
    with open(decode_file, encoding='utf-8') as f:
        for line, lm in zip(f, lm_all):
            lm_of[line.split('\t', 1)[0]].append(lm)
    files = []
    for weight in weights:
        name = save_dir + '/rescore_' + str(weight)
        with open(name, 'w', encoding='utf-8') as out:
            for key, (am, results) in table.items():
                total = np.array(am) + np.array(lm_of[key]) / weight
                out.write(key + ' ' + results[int(total.argmax())] + '\n')
        files.append(name)
    return files


# ------------------------------------------------------------------------------------------------ scoring
def edit_counts(ref: Sequence[str], hyp: Sequence[str]) -> Tuple[int, int, int]:
    """(insertions, deletions, substitutions) of a minimum word edit script turning `ref` into `hyp`; among equal-cost
    scripts a match/substitution is preferred, then a deletion, then an insertion."""
    n_hyp = len(hyp)
    # row over hyp positions; each cell = (total, ins, dele, sub)
    prev = [(j, j, 0, 0) for j in range(n_hyp + 1)]
    for r in ref:
        cur = [(prev[0][0] + 1, prev[0][1], prev[0][2] + 1, prev[0][3])]
        for j in range(1, n_hyp + 1):
            diag, up, left = prev[j - 1], prev[j], cur[j - 1]
            sub_cost = diag[0] + (0 if r == hyp[j - 1] else 1)
            del_cost = up[0] + 1
            ins_cost = left[0] + 1
            if sub_cost <= del_cost and sub_cost <= ins_cost:
                cur.append((sub_cost, diag[1], diag[2], diag[3] + (0 if r == hyp[j - 1] else 1)))
            elif del_cost <= ins_cost:
                cur.append((del_cost, up[1], up[2] + 1, up[3]))
            else:
                cur.append((ins_cost, left[1] + 1, left[2], left[3]))
        prev = cur
    _, ins, dele, sub = prev[n_hyp]
    return ins, dele, sub


def _read_text(path) -> "OrderedDict[str, List[str]]":
    table: "OrderedDict[str, List[str]]" = OrderedDict()
    with open(path, encoding='utf-8') as f:
        for line in f:
            fields = line.split()
            if fields:
                table[fields[0]] = fields[1:]
    return table


def compute_wer(ref_text, hyp_text, mode='present') -> dict:
    """Word error rate of `hyp_text` ('key words' lines) against `ref_text`, like `compute-wer --mode=<mode>`:
    'present' scores only utterances that have a hypothesis, 'all' counts a missing hypothesis as empty, 'strict'
    raises on a missing one.  -> dict(wer, ser, errors, words, ins, del, sub, sentences, sentence_errors, absent)."""
    if mode not in ('present', 'all', 'strict'):
        raise ValueError("[ERROR] compute_wer mode must be 'present', 'all' or 'strict'")
    ref = _read_text(ref_text) if not isinstance(ref_text, dict) else ref_text
    hyp = _read_text(hyp_text) if not isinstance(hyp_text, dict) else hyp_text
    words = ins = dele = sub = sents = sent_err = absent = 0
    for key, ref_words in ref.items():
        if key not in hyp:
            if mode == 'strict':
                raise ValueError('[ERROR] no hypothesis for utterance {}'.format(key))
            absent += 1
            if mode == 'present':
                continue
            hyp_words: List[str] = []
        else:
            hyp_words = hyp[key]
        i, d, s = edit_counts(ref_words, hyp_words)
        words += len(ref_words)
        ins, dele, sub = ins + i, dele + d, sub + s
        sents += 1
        sent_err += 1 if (i + d + s) else 0
    errors = ins + dele + sub
    return dict(wer=100.0 * errors / words if words else 0.0, ser=100.0 * sent_err / sents if sents else 0.0,
                errors=errors, words=words, ins=ins, sub=sub, sentences=sents, sentence_errors=sent_err, absent=absent,
                **{'del': dele})


def format_wer(stats: dict) -> str:
    """The three report lines of compute-wer."""
    return ('%WER {:.2f} [ {} / {}, {} ins, {} del, {} sub ]\n%SER {:.2f} [ {} / {} ]\n'
            'Scored {} sentences, {} not present in hyp.\n').format(
                stats['wer'], stats['errors'], stats['words'], stats['ins'], stats['del'], stats['sub'], stats['ser'],
                stats['sentence_errors'], stats['sentences'], stats['sentences'], stats['absent'])


def score_rescored(ref_text, scoring_dir, mode='present') -> "OrderedDict[str, dict]":
    """The WER loop of P/run.sh:196-199: score every `rescore_*` file of `scoring_dir`, write `<file>_wer` next to it."""
    out: "OrderedDict[str, dict]" = OrderedDict()
    ref = _read_text(ref_text)
    for name in sorted(os.listdir(scoring_dir)):
        if not name.startswith('rescore') or name.endswith('_wer'):
            continue
        stats = compute_wer(ref, os.path.join(scoring_dir, name), mode)
        with open(os.path.join(scoring_dir, name + '_wer'), 'w', encoding='utf-8') as f:
            f.write(format_wer(stats))
        out[name] = stats
    return out


_WER_LINE = re.compile(r'%WER\s+(\S+)')


def best_wer(wer_files: Iterable[str]) -> Tuple[str, float, str]:
    """`grep WER files | best_wer.sh` (K/best_wer.sh): the file with the lowest %WER (the first one on ties).
    -> (file, wer, '%WER ... file' line)."""
    best = None
    for name in wer_files:
        with open(name, encoding='utf-8') as f:
            for line in f:
                m = _WER_LINE.search(line)
                if m and (best is None or float(m.group(1)) < best[1]):
                    best = (name, float(m.group(1)), line.strip() + ' ' + name)
    if best is None:
        raise ValueError('[ERROR] no %WER line in the given files')
    return best
