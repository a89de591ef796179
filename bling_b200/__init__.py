"""bling_b200: B200-native path-tracing core for waldheinz/bling's path integrator."""
from . import ir  # noqa: F401
