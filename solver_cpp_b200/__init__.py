"""solver_cpp_b200 -- B200-native hot path of the DAMASK grid solver (dmentock/solver_cpp), host-side Python mirror.

The compute path is libdamask_grid_b200.so (hand-written sm_100a CUDA behind the C ABI in include/dmk_grid.h).
PyTorch is used for device memory, streams and torch.distributed only."""
from . import lib  # noqa: F401
