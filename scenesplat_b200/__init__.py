"""scenesplat_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of SceneSplat's
Pointcept / PTv3 3DGS encoder, behind the reference's own operator surface.

Importing the package does not touch the GPU; the first kernel call loads the in-tree C-ABI library
(`_C/libscenesplat_b200.so`, built by `python -m scenesplat_b200.build`).  There is no CPU fallback.
"""
from .registry import LOSSES, MODELS, TRANSFORMS, build_model, register_into_pointcept  # noqa: F401
from .structure import Point  # noqa: F401
from .ptv3 import (Block, Embedding, MLP, PointTransformerV3, PointTransformerV3SimDINO, SerializedAttention,  # noqa: F401
                   SerializedPooling, SerializedUnpooling)
from .lang import (AggregatedContrastiveLoss, ChunkPipeline, CosineSimilarity, Criteria, L2Loss,  # noqa: F401
                   LangPretrainer, zero_shot_accumulate, zero_shot_labels)
from .transform import GridSample, SphereCrop  # noqa: F401
from .voting import clustering_voting, confusion_update, neighbor_voting  # noqa: F401
from .spconv_compat import SparseConvTensor, SubMConv3d  # noqa: F401
from . import scene_io  # noqa: F401  (packed scene files: pack_scene / load_scene)
from . import compat  # noqa: F401  (torch_scatter / flash_attn / spconv stand-ins: compat.install())

__version__ = "0.1.0"
