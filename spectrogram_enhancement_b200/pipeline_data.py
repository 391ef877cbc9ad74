"""Drop-in namespace for `spec_denoising/pipeline_data.py` of the reference: the same module-level
names (specgr, norm, rescale, quantfilt), computed by libspecgpu.

    # before:  from pipeline_data import specgr, quantfilt
    # after:   from spectrogram_enhancement_b200.pipeline_data import specgr, quantfilt

The cv2 stages (gaussblr, meansub, morph; pipeline_data.py:52-72) run on the GPU too: OpenCV's fixed-point uint8
Gaussian and rectangle morphology are reproduced bit for bit, so cv2 is not needed.
"""
from .api import specgr, norm, rescale, quantfilt, gaussblr, meansub, morph, filter_chain  # noqa: F401

# pipeline_data.py:77-84
spec_params = {"nperseg": 512, "noverlap": 256, "fs": 500000, "window": "hamm", "scaling": "density",
               "detrend": "linear", "eps": 1e-11}
