"""B200-native TensoRF-VM ray renderer: host mirror of tensorf-myc's TensorVMSplit /
OctreeRender_trilinear_fast over libtvmrender.so (hand-written sm_100a CUDA, include/tvmrender.h)."""
from . import _lib
from ._lib import TvmError, LIB_PATH
from .tensorf import TensorVMSplit, REFTensoRF, NerfPlusPlus, AlphaGridMask, MLPRender_Fea, derive_march_scalars, unpack_bits, model_from_params
from .renderer import OctreeRender_trilinear_fast
from .train_ops import TVLoss, Adam, TrainStepGraph
from .maintain import get_rays_frame
from .checkpoint import save_checkpoint, load_checkpoint
from . import dist

__all__ = ["TensorVMSplit", "REFTensoRF", "NerfPlusPlus", "AlphaGridMask", "MLPRender_Fea", "OctreeRender_trilinear_fast",
           "derive_march_scalars", "unpack_bits", "model_from_params", "TvmError", "LIB_PATH", "TVLoss", "Adam", "TrainStepGraph", "get_rays_frame",
           "save_checkpoint", "load_checkpoint"]
