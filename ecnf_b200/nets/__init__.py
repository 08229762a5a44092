from .egnn import EGNN, init_flat_params  # noqa: F401
