"""fplplus_b200 -- B200-native (sm_100a) implementation of the FPL+ hot path:
the DSBN 3-D U-Net train step and the filtered-pseudo-label pass, behind PyMIC's
plugin surface (net_dict / loss_dict / Inferer / net_run_dsbn agent).

Host code is Python/PyTorch (device memory, streams, torch.distributed); all
arithmetic on the path is hand-written CUDA reached through the C ABI in
include/fplplus_b200.h.  There is no CPU or library fallback: importing the
compute modules builds/loads libfplplus_b200.so and fails loudly otherwise.
"""
__version__ = "0.1.0"

from .registry import net_dict, loss_dict  # noqa: E402,F401
