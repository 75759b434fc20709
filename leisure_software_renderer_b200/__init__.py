"""B200-native rasterization hot path of leisure-software-renderer."""
from . import capi  # noqa: F401
