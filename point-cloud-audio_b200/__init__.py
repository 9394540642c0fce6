"""pcaudio_b200 -- B200-native drop-in for the audio -> point-cloud -> set-encoder hot path of
SubramaniKrishna/point-cloud-audio.  Host mirrors keep the reference's names and signatures; the
arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI (include/pcaudio_b200.h)."""
from . import _lib
from .data_processing import load_esc, tt_split
from .dataset import (ESC_pc, ESC_pc_ss, ESC_pc_temp, ESC_pc_temp_importancerandKSS, ESC_pc_temp_maxKSS,
                      ESC_pc_temp_randKSS)
from .frontend import (build_clouds, coord_tables, gather_points, importance_heat, importance_points, random_points, resample,
                       select_points, spectral_point_cloud, stft_logmag, topk_points)
from .models import ST, DeepSet, SetTransformer, SetTransformerSAB, strip_module_prefix
from .modules import ISAB, MAB, PMA, SAB, invalidate_packed
from .pipeline import AudioConfig, AudioSetPipeline
from .training import SetTrainer, STTrainFunction
from .utils import pc_maxK, pc_randK

__all__ = ["load_esc", "tt_split", "ESC_pc", "ESC_pc_ss", "ESC_pc_temp", "ESC_pc_temp_maxKSS", "build_clouds",
           "coord_tables", "select_points", "spectral_point_cloud", "stft_logmag", "topk_points", "ST", "DeepSet",
           "SetTransformer", "SetTransformerSAB", "strip_module_prefix", "ISAB", "MAB", "PMA", "SAB", "AudioConfig",
           "AudioSetPipeline", "pc_maxK", "pc_randK", "ESC_pc_temp_randKSS", "ESC_pc_temp_importancerandKSS", "gather_points",
           "importance_heat", "importance_points", "random_points", "resample", "SetTrainer", "STTrainFunction", "invalidate_packed"]
