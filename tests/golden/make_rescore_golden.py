#!/usr/bin/env python
"""Generate tests/golden/rescore_small.json by running the REAL reference script L/rescore.py (build container only):

    python tests/golden/make_rescore_golden.py

Inputs (an n-best decode file in the format L/decode.py:154-161 writes, one LM score per line) are seeded and stored
next to the files the script produced, so the test needs nothing but the JSON."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
R = os.environ.get("PKA_REFERENCE", "/root/reference")
WEIGHTS = "10,11,13.5,20,1000"
PHONES = ["aa", "b", "ch", "d", "eh", "sil"]


def main():
    rng = np.random.RandomState(77)
    decode_lines, lm_lines = [], []
    for u in range(9):
        n_hyp = int(rng.randint(1, 5))
        am = np.sort(rng.uniform(-30, -1, size=n_hyp))[::-1]              # best first, like the lattice output
        for h in range(n_hyp):
            words = [PHONES[i] for i in rng.randint(0, len(PHONES), size=int(rng.randint(0, 7)))]
            decode_lines.append("utt%02d\t%s\t%s\n" % (u, str(float(am[h])), " ".join(words)))
            lm_lines.append("%s\n" % str(float(rng.uniform(-50, -20))))
    decode_lines.append("utt03\t-31.5\tsil aa sil\n")                       # a key that comes back later in the file
    lm_lines.append("-2.25\n")
    decode_txt, lm_txt = "".join(decode_lines), "".join(lm_lines)
    with tempfile.TemporaryDirectory() as tmp:
        open(tmp + "/decode.txt", "w").write(decode_txt)
        open(tmp + "/lm.txt", "w").write(lm_txt)
        os.mkdir(tmp + "/scoring")
        subprocess.run([sys.executable, R + "/project/attention-transformer-timit/local/rescore.py", "-decode_file",
                        tmp + "/decode.txt", "-lm_score", tmp + "/lm.txt", "-save_dir", tmp + "/scoring",
                        "-inv_weight_list", WEIGHTS], check=True, stdout=subprocess.DEVNULL)
        outputs = {name: open(tmp + "/scoring/" + name).read() for name in sorted(os.listdir(tmp + "/scoring"))}
    json.dump(dict(decode=decode_txt, lm=lm_txt, weights=WEIGHTS, outputs=outputs),
              open(os.path.join(HERE, "rescore_small.json"), "w"), indent=1)
    print("wrote rescore_small.json:", sorted(outputs))


if __name__ == "__main__":
    main()
