"""Drop-in alias: ``import dpilqr`` resolves to the B200-native implementation."""

from dpilqr_b200 import *  # noqa: F401,F403
from dpilqr_b200 import _reset_ids  # noqa: F401
from dpilqr_b200 import control, cost, distributed, dynamics, problem, util  # noqa: F401
