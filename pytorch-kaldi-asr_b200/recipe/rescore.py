"""LM rescoring entry point: the command line of L/rescore.py:12-21 in front of `results.rescore`.

`-inv_weight_list 5,10,15` -- each weight DIVIDES the language-model score; one `rescore_<w>` file per weight is written
to `-save_dir`, holding the best hypothesis of every utterance of `-decode_file` under `model + lm / w`."""
import argparse

FLAGS = ('decode_file', 'lm_score', 'save_dir', 'inv_weight_list')


def main(argv=None):
    from .. import results
    parser = argparse.ArgumentParser(description=__doc__)
    for name in FLAGS:
        parser.add_argument('-' + name, required=True)
    opt = parser.parse_args(argv)
    print('[PROCEDURE] start rescoring...')
    written = results.rescore(opt.decode_file, opt.lm_score, opt.save_dir, opt.inv_weight_list)
    print('[INFO] rescoring finished: {} files'.format(len(written)))
    return written


if __name__ == '__main__':
    main()
