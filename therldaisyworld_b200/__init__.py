"""B200-native RLDaisyWorld simulation step (drop-in for daisy.daisy_world_rl.RLDaisyWorld).

The compute path is hand-written sm_100a CUDA behind the C ABI in include/daisyworld_b200.h; there is
no CPU fallback -- constructing an environment without the built library or without a CUDA device raises."""
from ._lib import DaisyWorldError, LIB_PATH  # noqa: F401
from .env import RLDaisyWorld, make_neighborhood, query_kwargs  # noqa: F401

__all__ = ["RLDaisyWorld", "DaisyWorldError", "make_neighborhood", "query_kwargs", "LIB_PATH"]
