"""Shims that let the UNMODIFIED reference import and run on a CPU-only box.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py in the build container, where
/root/reference exists; nothing on the GPU box reads /root/reference).

SURVEY.md section 8(c): (1) stub matplotlib / hdf5storage (imported at
tools/hyper_tools.py:2,5, absent here); (2) ``.cuda()`` -> identity
(train.py:127-128,157-171; hyper_tools.py:422-423); (3) run from a cwd that
contains ./dataset/.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def install():
    import torch

    for name in ("matplotlib", "matplotlib.pyplot", "hdf5storage"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    plt = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].pyplot = plt
    for fn in ("axis", "imshow", "imsave", "xticks", "savefig"):
        setattr(plt, fn, lambda *a, **k: None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def capture_locals(func, *args, names=(), **kwargs):
    """Run ``func`` and return (result, {name: value}) for the named locals of its frame
    at return time -- lets us read ``loss_hist`` etc. out of train.main unmodified."""
    grabbed = {}
    code = func.__code__

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is code:
            for n in names:
                if n in frame.f_locals:
                    grabbed[n] = frame.f_locals[n]

    sys.setprofile(prof)
    try:
        res = func(*args, **kwargs)
    finally:
        sys.setprofile(None)
    return res, grabbed
