"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (read from /root/reference).

The reference (FedeMont/collision_handling_in_instantNGP) is a pure-PyTorch project whose modules
cannot be imported on a CPU-only box as they are:

* functions.py:8-10 imports matplotlib (absent in this image)            -> stub modules are injected;
* functions.py:49-52 forces ``torch.set_default_device('cuda')``          -> made a no-op during import;
* models.py / utils.py star-import the global ``device`` from functions   -> rebound to CPU afterwards.

Nothing here is shipped or used by the product path; it only exists so that ``oracle/make_goldens.py``
can run the reference in THIS container and write golden vectors to ``tests/golden``.  The reference tree
does not exist on the GPU box, so nothing under tests/, bench.py or smoke() imports this file at run time.
"""
import importlib
import importlib.machinery
import os
import sys
import types
from unittest import mock

REFERENCE_DIR = os.environ.get("GNGF_REFERENCE_DIR", "/root/reference")


def _stub_module(name):
    m = mock.MagicMock(name=name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__name__ = name
    return m


def load_reference(device="cpu"):
    """Returns the reference's (functions, models, utils, params) modules, bound to `device`."""
    import torch

    if not os.path.isdir(REFERENCE_DIR):
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    os.environ.setdefault("WANDB_MODE", "disabled")

    if "matplotlib" not in sys.modules:
        mpl = _stub_module("matplotlib")
        ticker = _stub_module("matplotlib.ticker")
        pyplot = _stub_module("matplotlib.pyplot")
        figure = _stub_module("matplotlib.figure")

        class Figure:  # used in a return annotation (functions.py:363)
            pass

        figure.Figure = Figure
        mpl.figure = figure
        mpl.ticker = ticker
        mpl.pyplot = pyplot
        pyplot.subplots = lambda *a, **k: (mock.MagicMock(), mock.MagicMock())
        sys.modules.update({"matplotlib": mpl, "matplotlib.ticker": ticker,
                            "matplotlib.pyplot": pyplot, "matplotlib.figure": figure})

    # our own drop-in `models` shim lives at the repo root; make sure the reference's wins here
    for name in ("functions", "models", "utils", "params"):
        sys.modules.pop(name, None)
    sys.path.insert(0, REFERENCE_DIR)
    real_set_default_device = torch.set_default_device
    try:
        if device == "cpu":
            torch.set_default_device = lambda *_a, **_k: None
        functions = importlib.import_module("functions")
        params = importlib.import_module("params")
        utils = importlib.import_module("utils")
        models = importlib.import_module("models")
    finally:
        torch.set_default_device = real_set_default_device
        sys.path.remove(REFERENCE_DIR)
    assert os.path.realpath(models.__file__).startswith(os.path.realpath(REFERENCE_DIR)), models.__file__
    dev = torch.device(device)
    for m in (functions, models, utils):
        m.device = dev
    return types.SimpleNamespace(functions=functions, models=models, utils=utils, params=params)


def set_flag(ref, name, value):
    """params.py flags are star-imported (copied) into every module; rebind all copies."""
    for m in (ref.functions, ref.models, ref.utils, ref.params):
        setattr(m, name, value)
