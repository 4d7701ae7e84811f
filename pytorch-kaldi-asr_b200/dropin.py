"""Make the reference's own import lines resolve to this package.

The reference scripts (L/train.py, L/decode.py, L/initialize_model.py) import their model code through PYTHONPATH
(P/path.sh:7-13) as top-level modules:

    from transformer.Models import Transformer          from transformer.Optim import ScheduledOptim
    from transformer.Lattice import Lattice             from TDNN import TDNNLayer, ConcatLayer, LDALayer
    from utils import constants, instances_handler

`import pytorch_kaldi_asr_b200.dropin` (before those lines) registers the B200 implementations under exactly these
names, so the scripts run unchanged apart from that one import.  `utils.BatchLoader` (U/BatchLoader.py) resolves to
the loader of this package, and -- only when the external `kaldi_io` package the reference depends on is not
installed -- `kaldi_io` resolves to `utils/kaldi_ark.py` (`read_mat`, `read_mat_scp`, `read_mat_ark`).  `utils`
submodules this package does not provide (`utils.get_gpu`, imported by L/train.py:18) keep resolving to the reference's
`pytorch/utils` directory when it is on `sys.path`: that directory is appended to the `__path__` of our `utils`.
"""
import os
import sys

from . import TDNN as _tdnn
from . import transformer as _transformer
from . import utils as _utils
from .transformer import Lattice as _lattice, Layers as _layers, Models as _models, Modules as _modules
from .transformer import Optim as _optim, SubLayers as _sublayers
from .utils import BatchLoader as _batch_loader, constants as _constants, instances_handler as _ih, kaldi_ark as _ark


def install(override_utils: bool = False):
    sys.modules["transformer"] = _transformer
    for name, mod in (("Models", _models), ("Modules", _modules), ("SubLayers", _sublayers), ("Layers", _layers),
                      ("Optim", _optim), ("Lattice", _lattice)):
        sys.modules["transformer." + name] = mod
    sys.modules["TDNN"] = _tdnn
    if override_utils or "utils" not in sys.modules:
        sys.modules.setdefault("utils", _utils)
        sys.modules.setdefault("utils.constants", _constants)
        sys.modules.setdefault("utils.instances_handler", _ih)
        sys.modules.setdefault("utils.BatchLoader", _batch_loader)
        if sys.modules["utils"] is _utils:
            own = [os.path.realpath(p) for p in _utils.__path__]
            for entry in sys.path:
                cand = os.path.join(entry or ".", "utils")
                if os.path.isdir(cand) and os.path.realpath(cand) not in own and cand not in _utils.__path__:
                    _utils.__path__.append(cand)
    if "kaldi_io" not in sys.modules:
        import importlib.util
        if importlib.util.find_spec("kaldi_io") is None:
            sys.modules["kaldi_io"] = _ark


install()
