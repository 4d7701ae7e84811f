"""The recipe's stage driver: stages 3-5 of P/run.sh:64-203 as one launcher with the reference's directory layout.

    python -m pytorch_kaldi_asr_b200.recipe.run -data_dir data -lang_dir data/lang -exp_dir exp -gpus 8

    stage 3  initialize_model  -> <model_dir>/model.init                                   (P/run.sh:67-92)
    stage 4  train             -> <model_dir>/epoch.N.torch, best.*, combined.accuXX.torch (P/run.sh:94-132); log in train.log
    stage 5  for dev, test:    decode -> <model_dir>/decode_<set>/decode.txt               (P/run.sh:137-175)
                               LM scores -> decode_<set>/lm.3k.score.txt                   (:177-183, external `ngram`)
                               rescore   -> decode_<set>/scoring/rescore_<w>               (:185-190)
                               WER       -> decode_<set>/scoring/rescore_<w>_wer, decode_<set>/result.txt (:191-203)

`<model_dir>` = `<exp_dir>/model_<time><suffix>` like the reference (or `-model_dir` to continue an existing one); data
sets are `<data_dir>/{train,dev,test}<data_suffix>` with `feats.scp` + `text`, the vocabulary is `<lang_dir>/vocab.txt`,
the LDA transform `<data_dir>/lda.mat`.  Stages 0-2 of the reference (Kaldi feature extraction, CMVN, SRILM) are external
tools and out of scope.

Devices: the reference picks one free GPU by polling nvidia-smi (U/get_gpu.py) and submits through queue.pl; here
`-gpus N` runs training and decoding as N processes under torch.distributed.run (one per GPU: gradients summed over
NCCL, utterances of a decode sharded with no collective), N = 1 runs in this process.  The language-model scores come
from SRILM's `ngram` in the reference; `-lm_score_cmd` is a shell command that reads the hypothesis text on stdin and
writes one log-probability per line on stdout.  Without it every hypothesis gets LM score 0 (rescoring then keeps the
model's order) and the result says so.
"""
import argparse
import glob
import os
import subprocess
import sys
import time

MODEL_FLAGS = (('encoder_max_len', 500), ('decoder_max_len', 100), ('src_fold', 1), ('encoder_sub_sequence', '(-100,0)'),
               ('decoder_sub_sequence', '(-10,0)'), ('en_layers', 3), ('de_layers', 3), ('n_head', 2), ('en_d_model', 256),
               ('de_d_model', 128), ('d_k', 64), ('d_v', 64), ('en_dropout', 0.35), ('de_dropout', 0.35))   # P/run.sh:77-91


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('-stage', type=int, default=3, help='first stage to run (3 init, 4 train, 5 decode + score)')
    p.add_argument('-stop_stage', type=int, default=5)
    p.add_argument('-data_dir', default='data')
    p.add_argument('-lang_dir', default='data/lang')
    p.add_argument('-exp_dir', default='exp')
    p.add_argument('-data_suffix', default='_filtered', help='data sets are <data_dir>/{train,dev,test}<data_suffix>')
    p.add_argument('-model_suffix', default='')
    p.add_argument('-model_dir', default=None, help='use this model directory instead of a new exp/model_<time><suffix>')
    p.add_argument('-gpus', type=int, default=1)
    p.add_argument('-master_port', type=int, default=29517)
    for name, default in MODEL_FLAGS:
        p.add_argument('-' + name, type=type(default), default=default)
    p.add_argument('-init_seed', type=int, default=None)
    # stage 4 (P/run.sh:99-113)
    p.add_argument('-epoch', type=int, default=500)
    p.add_argument('-batch_size', type=int, default=100)
    p.add_argument('-optim_start_lr', type=float, default=0.001)
    p.add_argument('-optim_soft_coefficient', type=float, default=25000)
    p.add_argument('-save_interval', type=int, default=1)
    p.add_argument('-compute_mode', choices=('fp32', 'bf16'), default='bf16')
    p.add_argument('-graphed', action='store_true')
    p.add_argument('-clean_dir', action='store_true', help='remove epoch.* after training (P/run.sh:128-131)')
    # stage 5 (P/run.sh:150-160, 187)
    p.add_argument('-decode_sets', default='dev,test')
    p.add_argument('-max_token_seq_len', type=int, default=100)
    p.add_argument('-decode_batch_size', type=int, default=8)
    p.add_argument('-beam_size', type=int, default=25)
    p.add_argument('-nbest', type=int, default=10)
    p.add_argument('-inv_weight_list', default='10,11,12,13,13.5,14,14.5,15,15.5,16,16.5,17,18,19,20,1000')
    p.add_argument('-lm_score_cmd', default=None)
    return p


def _run_module(module, argv, gpus, port, log=None):
    """One recipe stage: in this process for one GPU, under torch.distributed.run (one process per GPU) otherwise."""
    if gpus <= 1:
        import importlib
        return importlib.import_module('pytorch_kaldi_asr_b200.recipe.' + module).main(argv)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(gpus), '--master-addr',
           '127.0.0.1', '--master-port', str(port), '-m', 'pytorch_kaldi_asr_b200.recipe.' + module] + list(argv)
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env['PYTHONPATH'] = root + os.pathsep + env.get('PYTHONPATH', '')
    if log is None:
        rc = subprocess.run(cmd, env=env).returncode
    else:
        with open(log, 'w') as f:
            rc = subprocess.run(cmd, env=env, stdout=f, stderr=subprocess.STDOUT).returncode
    if rc != 0:
        raise RuntimeError('[ERROR] stage {} failed with exit code {}{}'.format(module, rc, ' (see %s)' % log if log else ''))
    return None


def lm_scores(decode_file, out_file, cmd):
    """One LM log-probability per hypothesis line of `decode_file` (P/run.sh:179-182 pipes the text through SRILM)."""
    texts = []
    with open(decode_file, encoding='utf-8') as f:            # key <TAB> score <TAB> words, one hypothesis per line
        for line in f:
            parts = line.rstrip('\n').split('\t')
            texts.append(parts[2].strip() if len(parts) > 2 else '')
    if cmd is None:
        with open(out_file, 'w') as f:
            f.write('0.0\n' * len(texts))
        return False
    done = subprocess.run(cmd, shell=True, input='\n'.join(texts) + '\n', capture_output=True, text=True)
    if done.returncode != 0:
        raise RuntimeError('[ERROR] -lm_score_cmd failed: {}'.format(done.stderr.strip()[:300]))
    scores = [s for s in done.stdout.split('\n') if s.strip()]
    if len(scores) != len(texts):
        raise RuntimeError('[ERROR] -lm_score_cmd wrote {} scores for {} hypotheses'.format(len(scores), len(texts)))
    with open(out_file, 'w') as f:
        f.write('\n'.join(scores) + '\n')
    return True


def main(argv=None):
    opt = build_parser().parse_args(argv)
    vocab = os.path.join(opt.lang_dir, 'vocab.txt')
    sets = {name: os.path.join(opt.data_dir, name + opt.data_suffix) for name in ('train', 'dev', 'test')}
    model_dir = opt.model_dir or os.path.join(opt.exp_dir, 'model_' + time.strftime('%Y%m%d-%H%M%S') + opt.model_suffix)
    summary = {'model_dir': model_dir}
    if opt.stage <= 3 <= opt.stop_stage:
        print('[PROCEDURE] reading dimension from data file and initialize the model')
        os.makedirs(model_dir, exist_ok=True)
        argv3 = ['-read_feats_scp_file', sets['train'] + '/feats.scp', '-read_vocab_file', vocab, '-save_model_file',
                 model_dir + '/model.init', '-lda_mat_file', os.path.join(opt.data_dir, 'lda.mat')]
        for name, _ in MODEL_FLAGS:
            argv3 += ['-' + name, str(getattr(opt, name))]
        if opt.init_seed is not None:
            argv3 += ['-init_seed', str(opt.init_seed)]
        _run_module('initialize_model', argv3, 1, opt.master_port)
    if opt.stage <= 4 <= opt.stop_stage:
        print('[PROCEDURE] trainning start... log is in train.log')
        argv4 = ['-read_train_dir', sets['train'], '-read_dev_dir', sets['dev'], '-read_test_dir', sets['test'],
                 '-read_vocab_file', vocab, '-load_model_file', model_dir + '/model.init', '-seq_error_prob', '0',
                 '-optim_start_lr', str(opt.optim_start_lr), '-optim_soft_coefficient', str(opt.optim_soft_coefficient),
                 '-epoch', str(opt.epoch), '-batch_size', str(opt.batch_size), '-save_model_dir', model_dir,
                 '-save_interval', str(opt.save_interval), '-compute_mode', opt.compute_mode, '-use_gpu']
        if opt.graphed:
            argv4.append('-graphed')
        summary['train_accuracy'] = _run_module('train', argv4, opt.gpus, opt.master_port, model_dir + '/train.log'
                                                if opt.gpus > 1 else None)
        print('[INFO] trainning finish.')
        if opt.clean_dir:
            for f in glob.glob(model_dir + '/epoch.*'):
                os.remove(f)
            print('[INFO] trainning dir cleaned')
    if opt.stage <= 5 <= opt.stop_stage:
        combined = sorted(glob.glob(model_dir + '/combine*'))
        if len(combined) != 1:
            raise RuntimeError('[ERROR] {} is not a file.'.format(model_dir + '/combine*'))   # P/run.sh:141-145
        model_file = combined[0]
        from . import rescore, score
        for name in [s for s in opt.decode_sets.split(',') if s]:
            data_dir = os.path.join(opt.data_dir, name + opt.data_suffix)
            decode_dir = os.path.join(model_dir, 'decode_' + name)
            os.makedirs(decode_dir + '/scoring', exist_ok=True)
            print('[PROCEDURE] decoding {} set... model file is {}'.format(name, model_file))
            argv5 = ['-read_data_dir', data_dir, '-read_vocab_file', vocab, '-load_model_file', model_file,
                     '-max_token_seq_len', str(opt.max_token_seq_len), '-batch_size', str(opt.decode_batch_size),
                     '-beam_size', str(opt.beam_size), '-nbest', str(opt.nbest), '-save_result_file',
                     decode_dir + '/decode.txt', '-use_gpu']
            _run_module('decode', argv5, opt.gpus, opt.master_port + 1, decode_dir + '/decode.log' if opt.gpus > 1 else None)
            print('[PROCEDURE] rescoring...')
            print('[INFO] caculating language model score...')
            with_lm = lm_scores(decode_dir + '/decode.txt', decode_dir + '/lm.3k.score.txt', opt.lm_score_cmd)
            print('[INFO] language model score computed.' if with_lm else
                  '[INFO] no -lm_score_cmd: language model scores are 0, rescoring keeps the acoustic order')
            rescore.main(['-decode_file', decode_dir + '/decode.txt', '-lm_score', decode_dir + '/lm.3k.score.txt',
                          '-inv_weight_list', opt.inv_weight_list, '-save_dir', decode_dir + '/scoring'])
            best = score.main(['-text', data_dir + '/text', '-scoring_dir', decode_dir + '/scoring', '-result_file',
                               decode_dir + '/result.txt'])
            summary['wer_' + name] = best[1]
    return summary


if __name__ == '__main__':
    main()
