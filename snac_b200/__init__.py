"""snac_b200 -- B200-native batched simulator for SNAC's mobile-construction environments.

Importing the package loads libdmp.so (hand-written sm_100a kernels behind the C ABI of
include/dmp.h) and fails loudly if it is missing: there is no CPU fallback."""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is absent)
from .vecenv import BatchedDMPEnv, generate_plans, load_plan_dataset  # noqa: F401
from .policy_loop import DeviceRollout, EpsilonGreedy, QSAAdapter, RandomPolicy  # noqa: F401
from .compat import (FlatObsEnv, HostStepper, VectorizedEnvWrapper, deep_mobile_printing_1d1r,  # noqa: F401
                     deep_mobile_printing_1d1r_dynamic, deep_mobile_printing_2d1r,
                     deep_mobile_printing_2d1r_dynamic, deep_mobile_printing_3d1r,
                     deep_mobile_printing_3d1r_dynamic, deep_mobile_printing_1d1r_Lnet,
                     deep_mobile_printing_2d1r_Lnet, deep_mobile_printing_3d1r_Lnet,
                     deep_mobile_printing_1d1r_hindsight, deep_mobile_printing_2d1r_hindsight,
                     deep_mobile_printing_3d1r_hindsight, deep_mobile_printing_1d1r_hindsight_static,
                     deep_mobile_printing_2d1r_hindsight_static, deep_mobile_printing_3d1r_hindsight_static,
                     deep_mobile_printing_1d1r_MCTS, deep_mobile_printing_1d1r_MCTS_obs,
                     deep_mobile_printing_2d1r_MCTS, deep_mobile_printing_2d1r_MCTS_dynamic,
                     deep_mobile_printing_3d1r_MCTS, deep_mobile_printing_3d1r_MCTS_dynamic)

__all__ = ["FlatObsEnv", "BatchedDMPEnv", "load_plan_dataset", "generate_plans", "DeviceRollout", "EpsilonGreedy", "QSAAdapter",
           "RandomPolicy", "HostStepper", "VectorizedEnvWrapper",
           "deep_mobile_printing_1d1r", "deep_mobile_printing_1d1r_dynamic",
           "deep_mobile_printing_2d1r", "deep_mobile_printing_2d1r_dynamic",
           "deep_mobile_printing_3d1r", "deep_mobile_printing_3d1r_dynamic",
           "deep_mobile_printing_1d1r_Lnet", "deep_mobile_printing_2d1r_Lnet", "deep_mobile_printing_3d1r_Lnet",
           "deep_mobile_printing_1d1r_hindsight", "deep_mobile_printing_2d1r_hindsight",
           "deep_mobile_printing_3d1r_hindsight", "deep_mobile_printing_1d1r_hindsight_static",
           "deep_mobile_printing_2d1r_hindsight_static", "deep_mobile_printing_3d1r_hindsight_static",
           "deep_mobile_printing_1d1r_MCTS", "deep_mobile_printing_1d1r_MCTS_obs", "deep_mobile_printing_2d1r_MCTS",
           "deep_mobile_printing_2d1r_MCTS_dynamic", "deep_mobile_printing_3d1r_MCTS",
           "deep_mobile_printing_3d1r_MCTS_dynamic"]
