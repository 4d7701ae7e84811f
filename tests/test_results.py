"""CPU tests of the decoder's output side: n-best file, LM rescoring (against files written by
the reference's own L/rescore.py, tests/golden/rescore_small.json), word error rate."""
import io
import json
import os

import pytest

from pytorch_kaldi_asr_b200 import results

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "rescore_small.json")


def test_nbest_file_format():
    idx2word = {0: "<blank>", 1: "<unk>", 2: "<s>", 3: "</s>", 4: "aa", 5: "b"}
    buf = io.StringIO()
    hyps = [[[2, 4, 5, 3], [2, 5, 3], [2, 3]], [[2, 4, 99, 3]]]
    scores = [[-1.5, -2.25, -3.0, -9.0], [-0.125]]                    # the beam may hold more scores than hypotheses
    assert results.write_nbest(buf, ("u1", "u2"), hyps, scores, idx2word) == 4
    assert buf.getvalue() == "u1\t-1.5\taa b\nu1\t-2.25\tb\nu1\t-3.0\t\nu2\t-0.125\taa <unk>\n"


def test_rescoring_equals_the_reference_script(tmp_path):
    g = json.load(open(GOLDEN))
    (tmp_path / "decode.txt").write_text(g["decode"])
    (tmp_path / "lm.txt").write_text(g["lm"])
    (tmp_path / "scoring").mkdir()
    files = results.rescore(str(tmp_path / "decode.txt"), str(tmp_path / "lm.txt"), str(tmp_path / "scoring"), g["weights"])
    assert sorted(os.path.basename(f) for f in files) == sorted(g["outputs"])
    for f in files:
        assert open(f).read() == g["outputs"][os.path.basename(f)], f
    table = results.read_nbest(str(tmp_path / "decode.txt"))
    assert list(table)[:2] == ["utt00", "utt01"] and table["utt03"][1][-1] == "sil aa sil"
    (tmp_path / "short.txt").write_text("-1.0\n")
    with pytest.raises(ValueError):
        results.rescore(str(tmp_path / "decode.txt"), str(tmp_path / "short.txt"), str(tmp_path / "scoring"), [10])
    (tmp_path / "bad.txt").write_text("utt00 -1.0 aa\n")
    with pytest.raises(ValueError):
        results.read_nbest(str(tmp_path / "bad.txt"))


def brute_force_distance(ref, hyp):
    import functools

    @functools.lru_cache(None)
    def d(i, j):
        if i == 0 or j == 0:
            return i + j
        return min(d(i - 1, j) + 1, d(i, j - 1) + 1, d(i - 1, j - 1) + (ref[i - 1] != hyp[j - 1]))
    return d(len(ref), len(hyp))


def test_edit_counts():
    assert results.edit_counts("a b c".split(), "a b c".split()) == (0, 0, 0)
    assert results.edit_counts("a b c".split(), "a x c".split()) == (0, 0, 1)
    assert results.edit_counts("a b c".split(), "a c".split()) == (0, 1, 0)
    assert results.edit_counts("a c".split(), "a b c".split()) == (1, 0, 0)
    assert results.edit_counts([], "a b".split()) == (2, 0, 0)
    assert results.edit_counts("a b".split(), []) == (0, 2, 0)
    import random
    rnd = random.Random(0)
    for _ in range(200):
        ref = [rnd.choice("abcd") for _ in range(rnd.randint(0, 9))]
        hyp = [rnd.choice("abcd") for _ in range(rnd.randint(0, 9))]
        i, d, s = results.edit_counts(ref, hyp)
        assert i + d + s == brute_force_distance(tuple(ref), tuple(hyp))
        assert len(ref) - d + i == len(hyp)                            # the script really maps ref onto hyp


def test_wer_report_modes_and_best_wer(tmp_path):
    (tmp_path / "text").write_text("u1 a b c d\nu2 a b\nu3 c c c\n")
    (tmp_path / "scoring").mkdir()
    (tmp_path / "scoring" / "rescore_10.0").write_text("u1 a x c d\nu2 a b\n")
    (tmp_path / "scoring" / "rescore_20.0").write_text("u1 a b c d\nu2 a b b\nu3 c\n")
    present = results.compute_wer(str(tmp_path / "text"), str(tmp_path / "scoring" / "rescore_10.0"))
    assert (present["errors"], present["words"], present["sentences"], present["absent"]) == (1, 6, 2, 1)
    assert results.format_wer(present) == ("%WER 16.67 [ 1 / 6, 0 ins, 0 del, 1 sub ]\n%SER 50.00 [ 1 / 2 ]\n"
                                           "Scored 2 sentences, 1 not present in hyp.\n")
    everything = results.compute_wer(str(tmp_path / "text"), str(tmp_path / "scoring" / "rescore_10.0"), mode="all")
    assert (everything["errors"], everything["words"], everything["del"]) == (4, 9, 3)
    with pytest.raises(ValueError):
        results.compute_wer(str(tmp_path / "text"), str(tmp_path / "scoring" / "rescore_10.0"), mode="strict")
    scored = results.score_rescored(str(tmp_path / "text"), str(tmp_path / "scoring"))
    assert list(scored) == ["rescore_10.0", "rescore_20.0"] and scored["rescore_20.0"]["errors"] == 3
    wer_files = sorted(str(p) for p in (tmp_path / "scoring").iterdir() if p.name.endswith("_wer"))
    assert len(wer_files) == 2 and len(list((tmp_path / "scoring").iterdir())) == 4
    name, wer, line = results.best_wer(wer_files)
    assert name.endswith("rescore_10.0_wer") and wer == pytest.approx(16.67) and line.startswith("%WER 16.67 [ 1 / 6")
    scored_again = results.score_rescored(str(tmp_path / "text"), str(tmp_path / "scoring"))     # *_wer files are skipped
    assert list(scored_again) == ["rescore_10.0", "rescore_20.0"]
