"""`algos/` registry of the reference (CARL_MVF/algos/__init__.py:7-20); only SCL is on the hot path."""
from .scl import SCL

ALGO_NAME_TO_ALGO_CLASS = {"scl": SCL}


def get_algo(cfg):
    """Returns training algo."""
    name = cfg.TRAINING_ALGO
    if name not in ALGO_NAME_TO_ALGO_CLASS:
        raise ValueError("%s not supported yet." % name)
    return ALGO_NAME_TO_ALGO_CLASS[name](cfg)
