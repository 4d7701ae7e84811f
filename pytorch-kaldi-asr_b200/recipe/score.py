"""Scoring entry point: what P/run.sh:194-203 does with Kaldi's `compute-wer --mode=present` and `best_wer.sh` --
score every `rescore_*` file of a scoring directory against the reference transcripts, write `<file>_wer`, print and
return the best line."""
import argparse
import os


def main(argv=None):
    from .. import results
    parser = argparse.ArgumentParser()
    parser.add_argument('-text', required=True, help='reference transcripts: key word word ...')
    parser.add_argument('-scoring_dir', required=True)
    parser.add_argument('-mode', default='present', choices=('present', 'all', 'strict'))
    parser.add_argument('-result_file', default=None)
    opt = parser.parse_args(argv)
    print('[INFO] computing WER...')
    scored = results.score_rescored(opt.text, opt.scoring_dir, opt.mode)
    best = results.best_wer(os.path.join(opt.scoring_dir, name + '_wer') for name in scored)
    report = '[INFO] best wer presented in file:\n' + best[2] + '\n'
    if opt.result_file:
        with open(opt.result_file, 'w', encoding='utf-8') as f:
            f.write(report)
    print(report, end='')
    return best


if __name__ == '__main__':
    main()
