"""yad_b200 - B200-native (sm_100a) audio-activity-detection hot path.

Drop-in for the reference's ``modules.AudioDetectionNetwork`` and ``inference.process_model_outputs``;
every kernel is hand-written CUDA behind the C ABI in ``include/yad_b200.h``.  No CPU fallback."""
from ._lib import YadError, LIB_PATH  # noqa: F401
from .config import DEFAULT_CONFIG_PATH, default_config, load_config  # noqa: F401
from .network import AudioDetectionNetwork  # noqa: F401
from .postprocess import nms_raw, process_model_outputs  # noqa: F401
from .hostpipe import run_host_batch  # noqa: F401
from .evaluate import evaluate_waveform, rle_rows  # noqa: F401
from .anchors import compute_anchors, kmeans_lloyd, set_config_anchors, durations_from_annotations  # noqa: F401
from .train_ops import (AudioDetectionLoss, EMAParamsSmoothener, FusedAdamEMA, build_target_by_scale,  # noqa: F401
                        clip_targets, collate_batch, load_checkpoint, save_checkpoint)

__all__ = ["AudioDetectionNetwork", "process_model_outputs", "nms_raw", "run_host_batch", "evaluate_waveform", "rle_rows", "load_config", "default_config",
           "build_target_by_scale", "clip_targets", "collate_batch", "save_checkpoint", "load_checkpoint", "compute_anchors", "kmeans_lloyd", "set_config_anchors", "durations_from_annotations", "AudioDetectionLoss", "FusedAdamEMA", "EMAParamsSmoothener", "YadError"]
