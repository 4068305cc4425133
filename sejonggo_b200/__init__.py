"""sejonggo_b200 — B200-native batched self-play engine with the call surface of
drsagitn/sejonggo's self-play hot path (see DESIGN.md / INTEGRATION.md).

Importing the package does not touch CUDA; constructing an Engine does and raises
if the in-tree CUDA library or a CUDA device is missing (no CPU fallback)."""
from .conf import conf  # noqa: F401

__all__ = ["conf"]
