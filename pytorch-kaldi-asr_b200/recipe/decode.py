"""Decoding entry point -- flags and flow of L/decode.py:110-161: load a model file, beam-search a data directory, write
the n-best result file.  Under torchrun every rank decodes its share of the batches (no collective) into
`<save_result_file>.<rank>`; rank 0 concatenates them (rank by rank) at the end -- consumers group lines by key."""
import argparse
import os


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-read_data_dir', required=True)
    parser.add_argument('-read_vocab_file', required=True)
    parser.add_argument('-load_model_file', required=True)
    parser.add_argument('-save_result_file', required=True)
    parser.add_argument('-max_token_seq_len', type=int, required=True)
    parser.add_argument('-batch_size', type=int, default=64)
    parser.add_argument('-beam_size', type=int, default=20)
    parser.add_argument('-nbest', type=int, default=10)
    parser.add_argument('-use_gpu', action='store_true')
    # beyond the reference.  The TDNN encoder does not mask padding (its output on the last 16 real frames depends on
    # how many pad frames follow), so the default keeps the reference loader's whole-set padding and its exact results.
    parser.add_argument('-pad_to', choices=('dataset', 'batch'), default='dataset')
    return parser


def main(argv=None):
    from . import pick_device
    from .. import checkpoint, results
    from .. import train as T
    opt = build_parser().parse_args(argv)
    device, rank, world = pick_device()
    if opt.nbest > opt.beam_size:
        raise ValueError('[ERROR] nbest should not larger than beam_size')
    loaded = checkpoint.load_checkpoint(opt.load_model_file, device=device)
    model, model_options = loaded['model'], loaded['model_options']
    print('[INFO] loading model with parameter: {}'.format(model_options))
    decode_data = T.initialize_batch_loader(opt.read_data_dir + '/feats.scp', opt.read_data_dir + '/text',
                                            opt.read_vocab_file, opt.batch_size, mode='all', pad_to=opt.pad_to,
                                            seed=0 if world > 1 else None, shard=(rank, world))
    print('[INFO] batch loader is initialized')
    target = opt.save_result_file if world == 1 else '{}.{}'.format(opt.save_result_file, rank)
    n = results.decode_to_file(model, decode_data, opt, model_options, opt.read_vocab_file, target)
    print('[INFO] {} utterances decoded to {}'.format(n, target))
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=device)
        dist.barrier()
        if rank == 0:
            with open(opt.save_result_file, 'w', encoding='utf-8') as out:
                for r in range(world):
                    part = '{}.{}'.format(opt.save_result_file, r)
                    with open(part, encoding='utf-8') as f:
                        out.write(f.read())
                    os.remove(part)
        dist.barrier()
        dist.destroy_process_group()
    return n


if __name__ == '__main__':
    main()
