"""LM rescoring entry point -- flags of L/rescore.py:12-21."""
import argparse


def main(argv=None):
    from .. import results
    parser = argparse.ArgumentParser()
    parser.add_argument('-decode_file', required=True)
    parser.add_argument('-lm_score', required=True)
    parser.add_argument('-save_dir', required=True)
    parser.add_argument('-inv_weight_list', required=True)        # '5,10,15': the weight divides the LM score
    opt = parser.parse_args(argv)
    print('[PROCEDURE] start rescoring...')
    files = results.rescore(opt.decode_file, opt.lm_score, opt.save_dir, opt.inv_weight_list)
    print('[INFO] rescoring finished')
    return files


if __name__ == '__main__':
    main()
