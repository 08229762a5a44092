"""ecnf_b200: B200-native (sm_100a) drop-in for the hot path of Kalyan0821/ecnf-baseline-neurips-2023."""
from .engine import CnfConfig, Engine, PackedParams, ess_from_stats, key_to_seed, split_key  # noqa: F401
from . import lib  # noqa: F401
