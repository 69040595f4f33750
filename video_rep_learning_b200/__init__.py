"""video_rep_learning_b200 -- B200-native MV-Former head + SCL training hot path.

Layout mirrors the reference's interface for this path (CARL_MVF/models, CARL_MVF/algos):
  models/   TransformerModel, MultiEntityTransformerEmbModel, MLPHead, encoder parameter containers, build_model
  algos/    SCL, get_algo
  datasets/ sample_frames (integer-exact two-view temporal sampling)
  engine.py autograd Functions over the C ABI;  _lib.py ctypes binding;  csrc/ CUDA kernels (sm_100a)
"""
__version__ = "0.1.0"
