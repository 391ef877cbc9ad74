"""Load the UNMODIFIED reference functions from a checkout of
PlasmaControl/spectrogram-enhancement (default /root/reference).

TEST INFRASTRUCTURE ONLY (see oracle/spec_oracle.py header).  The reference tree exists only
in the build container, never on the GPU box, so this module is used solely by
tests/golden/make_golden.py (to generate committed fixtures) and by CPU tests that skip when
the tree is absent.  No reference source is copied: the files are imported / exec'd where they
lie.

* spec_denoising/pipeline_data.py is imported as a module; its plot/IO-only imports that are not
  installed here (matplotlib, patchify, skimage, h5py) are satisfied with empty stub modules --
  none of them is touched by specgr/norm/rescale/quantfilt/gaussblr/meansub/morph.
* omega / computeSignal / denoiseSignal live in a notebook: the code cell holding them
  (spec_denoising/denoising_by_svd.ipynb, cell 1) is exec'd in a namespace.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

REF_ROOT = os.environ.get("SPECGPU_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "patchify", "skimage",
    "skimage.exposure", "skimage.color", "skimage.data", "skimage.restoration", "h5py",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "spec_denoising", "pipeline_data.py"))


def _install_stubs():
    added = []
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError):
            pass
        m = types.ModuleType(name)
        m.__dict__.update(rescale_intensity=None, color=None, data=None, restoration=None,
                          patchify=None, unpatchify=None)
        sys.modules[name] = m
        added.append(name)
    return added


def load_pipeline_data():
    """The module object of spec_denoising/pipeline_data.py (its __main__ block does not run)."""
    added = _install_stubs()
    try:
        path = os.path.join(REF_ROOT, "spec_denoising", "pipeline_data.py")
        spec = importlib.util.spec_from_file_location("_ref_pipeline_data", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for name in added:
            sys.modules.pop(name, None)


def load_svd_notebook():
    """Namespace with omega, computeSignal, denoiseSignal, specgr (BES), quantfilt, ... from
    spec_denoising/denoising_by_svd.ipynb cell 1 (raw JSON lines 37-230)."""
    import numpy as np
    import pickle
    import scipy.signal
    import cv2

    nb = json.load(open(os.path.join(REF_ROOT, "spec_denoising", "denoising_by_svd.ipynb")))
    src = "".join(nb["cells"][1]["source"])
    ns = {"np": np, "pickle": pickle, "scipy": scipy, "cv2": cv2}
    exec(compile(src, "denoising_by_svd.ipynb#cell1", "exec"), ns)
    return ns
