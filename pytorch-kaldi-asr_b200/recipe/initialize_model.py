"""Build the initial model file from Kaldi data -- flags and behaviour of L/initialize_model.py:23-99: feature dimension
from the first scp entry, vocabulary size from the vocab file, frozen LDA from `lda.mat`, the recipe's fixed TDNN
contexts; written as an epoch-0 state-dict checkpoint (checkpoint.py)."""
import argparse

TDNN_CONTEXTS = [[-1, 0, 1], [-1, 0, 1], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3], [-3, 0, 3]]      # L/initialize_model.py:48-53


def str2tuple(string):
    """'(-100,0)' -> (-100, 0)   (L/initialize_model.py:13-21)"""
    body = string.strip()
    if not (body.startswith('(') and body.endswith(')')):
        raise ValueError('[ERROR] invalid sub-sequence string!')
    parts = body[1:-1].split(',')
    if len(parts) != 2:
        raise ValueError('[ERROR] invalid sub-sequence string!')
    return int(parts[0]), int(parts[1])


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-read_feats_scp_file', required=True)
    parser.add_argument('-lda_mat_file', required=True)
    parser.add_argument('-read_vocab_file', required=True)
    parser.add_argument('-encoder_max_len', type=int, required=True)
    parser.add_argument('-decoder_max_len', type=int, required=True)
    parser.add_argument('-src_fold', type=int, default=1)
    parser.add_argument('-encoder_sub_sequence', default='(-100,0)')
    parser.add_argument('-decoder_sub_sequence', default='(-20,0)')
    parser.add_argument('-en_layers', type=int, default=2)
    parser.add_argument('-de_layers', type=int, default=2)
    parser.add_argument('-n_head', type=int, default=3)
    parser.add_argument('-en_d_model', type=int, default=256)
    parser.add_argument('-de_d_model', type=int, default=128)
    parser.add_argument('-d_k', type=int, default=64)
    parser.add_argument('-d_v', type=int, default=64)
    parser.add_argument('-en_dropout', type=float, default=0.2)
    parser.add_argument('-de_dropout', type=float, default=0.2)
    parser.add_argument('-save_model_file', required=True)
    parser.add_argument('-init_seed', type=int, default=None, help='torch seed for the weight initialisation')
    return parser


def main(argv=None):
    import torch
    from .. import checkpoint
    from ..utils import instances_handler, kaldi_ark
    opt = build_parser().parse_args(argv)
    opt.tdnn_contexts = [list(c) for c in TDNN_CONTEXTS]
    opt.encoder_sub_sequence = str2tuple(opt.encoder_sub_sequence)
    opt.decoder_sub_sequence = str2tuple(opt.decoder_sub_sequence)
    for _, matrix in kaldi_ark.read_mat_scp(opt.read_feats_scp_file):
        opt.src_dim = int(matrix.shape[1])
        break
    else:
        raise ValueError('[ERROR] {} lists no utterance'.format(opt.read_feats_scp_file))
    print('[INFO] get feature of dimension {} from {}.'.format(opt.src_dim, opt.read_feats_scp_file))
    opt.tgt_vocab_dim = len(instances_handler.read_vocab(opt.read_vocab_file))
    print('[INFO] get label of dimension {} from {}.'.format(opt.tgt_vocab_dim, opt.read_vocab_file))
    print('[INFO] model will initialized with add_argument:\n\t{}.'.format(opt))
    lda_mat = kaldi_ark.read_mat(opt.lda_mat_file)
    if opt.init_seed is not None:
        torch.manual_seed(opt.init_seed)
    from ..transformer.Models import Transformer
    model = Transformer(lda_mat=lda_mat, **checkpoint.model_kwargs(opt))
    checkpoint.save_checkpoint(opt.save_model_file, model, opt, 0)
    print('[INFO] initialized model is saved to {}.'.format(opt.save_model_file))
    return opt


if __name__ == '__main__':
    main()
