"""Vocabulary control symbols -- same ids as the reference (U/constants.py:2-10); they define the tensor contract:
PAD doubles as the padding value of label tensors, the `ignore_index` of the loss and the "masked" value of pad masks."""
PAD, UNK, BOS, EOS = 0, 1, 2, 3
PAD_WORD, UNK_WORD, BOS_WORD, EOS_WORD = "<blank>", "<unk>", "<s>", "</s>"
