"""sigmarl_b200 — SigmaRL's vectorised road-traffic environment step as hand-written CUDA for sm_100a.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all environment arithmetic
runs in ``libsigmarl_b200.so`` (``csrc/``) behind the C-ABI of ``include/sigmarl_b200.h``.
There is no CPU fallback: creating an environment without the library or without a GPU raises.
"""
from .config import EnvConfig  # noqa: F401
from .maps import MapLibrary  # noqa: F401
from .lib import SgbError, load_library, library_path  # noqa: F401
from .env import RoadTrafficEnv  # noqa: F401
from .scenario import ScenarioRoadTrafficB200, make_env  # noqa: F401

__version__ = "0.1.0"
