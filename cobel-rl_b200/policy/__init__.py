"""Action-selection policies of the hot path (reference: cobel/policy/__init__.py)."""
from .policy import Policy  # noqa: F401
from .greedy import EpsilonGreedy, ExclusiveEpsilonGreedy  # noqa: F401
from .softmax import Softmax  # noqa: F401
