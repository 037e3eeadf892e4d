"""cobel_rl_b200 -- B200-native batched simulator for CoBeL-RL's tabular closed loop.

Runs N independent agents (seeds or parameter-sweep points) per launch behind
the reference's class API: ``interface.Gridworld`` / ``Topology``, agents
``DynaQ`` / ``QAgent`` / ``SR`` / ``PMA`` / ``SFMA``, memories ``DynaQMemory`` /
``PMAMemory`` / ``SFMAMemory``, policies ``EpsilonGreedy`` /
``ExclusiveEpsilonGreedy`` / ``Softmax``.  All work is done by hand-written
sm_100a CUDA kernels in ``libcobel_b200.so`` (C ABI: include/cobel_b200.h);
PyTorch only owns device memory, streams and ``torch.distributed``.  There is no
CPU fallback.
"""
__version__ = '0.1.0'

from .stream import BatchStream  # noqa: F401
from . import spaces  # noqa: F401
from . import interface, policy, memory, agent, misc, monitor, optimizer  # noqa: F401
