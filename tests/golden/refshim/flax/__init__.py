"""Stand-in for flax: only `flax.linen` (see ../README.md)."""
from . import linen  # noqa: F401
