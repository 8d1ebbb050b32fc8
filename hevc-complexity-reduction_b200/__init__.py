"""hevc_complexity_reduction_b200 -- B200-native ETH-CNN CU-partition predictor (host-side mirror).

The product is the C-ABI library ``libethcnn_b200.so`` (include/ethcnn.h; sources under csrc/).  This
package is the thin Python layer above it, mirroring the reference's Python interface for the path
(HM-16.5_Test_AI/bin/video_to_cu_depth.py, net_CNN.py): same names, same argument meaning, same
failure behaviour -- but every number is produced by the sm_100a kernels.  There is NO CPU fallback:
if the shared library is missing or no CUDA device is present, calls raise.

The directory is named ``hevc-complexity-reduction_b200`` (as the repo layout asks); import it through
the ``ethcnn_b200`` loader at the repo root, which registers it as ``hevc_complexity_reduction_b200``.
"""
from .binding import (  # noqa: F401
    EthCnn,
    EthCnnError,
    MODE_AI,
    MODE_LDP,
    PROBS_PER_CTU,
    FC1_WIDTH,
    STAGE_CONV,
    STAGE_FC1,
    STAGE_HEADS,
    STAGE_GATE,
    STAGE_NAMES,
    OPT_FC1_PATH,
    OPT_CHUNK_CTUS,
    OPT_STAGED_OUTPUT,
    OPT_CONV_PATH,
    Q_CONV_PATH,
    library_path,
    load_library,
    ctu_grid,
    unpack_decisions,
    request,
    request_quit,
)
from . import video_to_cu_depth, net_CNN, sharding, evaluation  # noqa: F401

__all__ = ["EthCnn", "EthCnnError", "MODE_AI", "MODE_LDP", "PROBS_PER_CTU", "FC1_WIDTH", "library_path",
           "load_library", "ctu_grid", "video_to_cu_depth", "net_CNN", "sharding", "evaluation", "unpack_decisions"]
