"""Build the initial model file from Kaldi data.  Same command line as L/initialize_model.py:23-46; same behaviour:
feature dimension from the first scp entry, vocabulary size from the vocab file, frozen LDA from `lda.mat`, the
recipe's fixed TDNN contexts (:48-53).  The result is an epoch-0 state-dict checkpoint (checkpoint.py)."""
import argparse

TDNN_CONTEXTS = ((-1, 0, 1),) * 2 + ((-3, 0, 3),) * 4
REQUIRED = object()

# flag, type, default -- the reference's names and defaults; `init_seed` is new (seed of the weight initialisation)
FLAGS = (
    ('read_feats_scp_file', str, REQUIRED), ('lda_mat_file', str, REQUIRED), ('read_vocab_file', str, REQUIRED),
    ('encoder_max_len', int, REQUIRED), ('decoder_max_len', int, REQUIRED), ('src_fold', int, 1),
    ('encoder_sub_sequence', str, '(-100,0)'), ('decoder_sub_sequence', str, '(-20,0)'),
    ('en_layers', int, 2), ('de_layers', int, 2), ('n_head', int, 3), ('en_d_model', int, 256), ('de_d_model', int, 128),
    ('d_k', int, 64), ('d_v', int, 64), ('en_dropout', float, 0.2), ('de_dropout', float, 0.2),
    ('save_model_file', str, REQUIRED), ('init_seed', int, None),
)


def str2tuple(text):
    """'(-100,0)' -> (-100, 0)"""
    try:
        lo, hi = text.strip().lstrip('(').rstrip(')').split(',')
        return int(lo), int(hi)
    except ValueError:
        raise ValueError('[ERROR] invalid sub-sequence string!') from None


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__)
    for name, kind, default in FLAGS:
        if default is REQUIRED:
            parser.add_argument('-' + name, type=kind, required=True)
        else:
            parser.add_argument('-' + name, type=kind, default=default)
    return parser


def main(argv=None):
    import torch
    from .. import checkpoint
    from ..transformer.Models import Transformer
    from ..utils import instances_handler, kaldi_ark
    opt = build_parser().parse_args(argv)
    opt.tdnn_contexts = [list(ctx) for ctx in TDNN_CONTEXTS]
    opt.encoder_sub_sequence, opt.decoder_sub_sequence = (str2tuple(s) for s in (opt.encoder_sub_sequence,
                                                                                 opt.decoder_sub_sequence))
    first = next(iter(kaldi_ark.read_mat_scp(opt.read_feats_scp_file)), None)
    if first is None:
        raise ValueError('[ERROR] {} lists no utterance'.format(opt.read_feats_scp_file))
    opt.src_dim = int(first[1].shape[1])
    opt.tgt_vocab_dim = len(instances_handler.read_vocab(opt.read_vocab_file))
    print('[INFO] feature dimension {} ({}), label dimension {} ({}).'.format(
        opt.src_dim, opt.read_feats_scp_file, opt.tgt_vocab_dim, opt.read_vocab_file))
    print('[INFO] model options:\n\t{}.'.format(opt))
    if opt.init_seed is not None:
        torch.manual_seed(opt.init_seed)
    model = Transformer(lda_mat=kaldi_ark.read_mat(opt.lda_mat_file), **checkpoint.model_kwargs(opt))
    checkpoint.save_checkpoint(opt.save_model_file, model, opt, 0)
    print('[INFO] initialized model is saved to {}.'.format(opt.save_model_file))
    return opt


if __name__ == '__main__':
    main()
