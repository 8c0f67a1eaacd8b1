"""lerf_pytorch_b200 -- B200-native (sm_100a) LUT inference hot path of LeRF.

Drop-in for the reference's (ddlee-cn/LeRF-PyTorch) LUT evaluation path only:
``FourSimplexInterpFaster`` (resample/eval_lut_sr.py:24-470), the ``*Resize2d*`` / ``*Warp2d*`` operators
(resize_right/resize_right2d_numpy.py) and the per-image bodies of eval_lut_sr.py / eval_lut_warp.py.
Everything runs in hand-written CUDA behind the C ABI in include/lerf_b200.h; there is no CPU fallback.
"""
from ._lib import LerfError, lib  # noqa: F401
from .lut_interp import FourSimplexInterpFaster, lut_stage1, lut_stage2, lut_stages, mode_pad_dict  # noqa: F401
from .luts import LutSet, load_lut_dict  # noqa: F401
from .pipeline import LerfSR, LerfWarp  # noqa: F401
from .resize_right2d import (  # noqa: F401
    AmplifiedLinearResize2d, AmplifiedLinearResize2dNumpy, AmplifiedLinearResize2dTorch,
    AmplifiedLinearWarp2d, AmplifiedLinearWarp2dNumpy, AmplifiedLinearWarp2dTorch,
    BicubicWarp2d, BicubicWarp2dNumpy, BilinearWarp2d, BilinearWarp2dNumpy, Lanczos2Warp2d, Lanczos2Warp2dNumpy,
    Lanczos3Warp2d, Lanczos3Warp2dNumpy, NearestWarp2d, NearestWarp2dNumpy, NearestWarp2dTorch,
    SteeringGaussianResize2d, SteeringGaussianResize2dNumpy, SteeringGaussianResize2dTorch,
    SteeringGaussianWarp2d, SteeringGaussianWarp2dNumpy, SteeringGaussianWarp2dTorch, sr_axis_tables)

from . import metrics  # noqa: F401
from .lut_finetune import LutFineTune, interp_torch_batch, steering_gaussian_resize  # noqa: F401
from .png_gpu import encode_png, png_bytes, save_png  # noqa: F401
from .sharding import band_halo_rows, band_input_rows, image_shard, row_bands  # noqa: F401

__version__ = "0.1.0"
