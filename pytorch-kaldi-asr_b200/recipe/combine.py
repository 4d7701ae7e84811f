"""Stand-alone model averaging -- flags and flow of L/combine.py:33-108: average the listed model files one by one
(running mean), evaluate each average on a data directory, save the best as `combined.accuXX.XX.torch`.

The reference stops after at most 10 models (`range(10)`, :85) and crashes with fewer; here the whole list is used."""
import argparse


def main(argv=None):
    from .. import train as T
    from ..utils import instances_handler
    parser = argparse.ArgumentParser()
    parser.add_argument('-read_test_dir', required=True)
    parser.add_argument('-read_vocab_file', required=True)
    parser.add_argument('-load_model_dir', required=True)
    parser.add_argument('-load_model_file_list', required=True, nargs='+')
    parser.add_argument('-save_model_dir', required=True)
    parser.add_argument('-use_gpu', action='store_true')
    parser.add_argument('-batch_size', type=int, default=96)          # L/combine.py:62 hard-wires 96
    opt = parser.parse_args(argv)
    from . import pick_device
    device, _, _ = pick_device()
    files = [opt.load_model_dir + '/' + name for name in opt.load_model_file_list]
    print('[INFO] reading test data...')
    test_data = T.initialize_batch_loader(opt.read_test_dir + '/feats.scp', opt.read_test_dir + '/text',
                                          opt.read_vocab_file, opt.batch_size)
    print('[INFO] batch loader is initialized')
    crit = T.get_criterion(len(instances_handler.read_vocab(opt.read_vocab_file)))
    return T.combine_files(files, crit, test_data, opt.save_model_dir, train_options=opt, device=device)


if __name__ == '__main__':
    main()
