"""B200-native FP64 SpGEMM engine behind the IA-SpGEMM front end."""
from . import workloads  # noqa: F401
