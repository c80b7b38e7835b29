"""yololite_b200: B200 (sm_100a) engine for YoloLite's detection forward pass + postprocess.

Host side in Python (the reference is Python); the product is the CUDA library behind include/yololite_b200.h.
Importing this package never touches oracle/ and never falls back to a CPU path.
"""
from ._lib import EXPORTS, LIB_PATH, lib            # noqa: F401
from .engine import YoloLiteB200                    # noqa: F401
from .post import (Detections, PostProcessor, backmap, decode_batch_to_coco_dets,   # noqa: F401
                   decode_preds_anchorfree, detect, unpack)
from .infer import (YoloLite, letterbox_geometry, load_model_names_imgsize_from_ckpt, preprocess,   # noqa: F401
                    preprocess_batch)

__version__ = "0.1.0"
