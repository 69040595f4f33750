from .sampling import sample_frames  # noqa: F401
