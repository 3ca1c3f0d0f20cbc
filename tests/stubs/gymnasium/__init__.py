"""Stand-in for `gymnasium` on boxes where it is not installed (this image has no wheel for it): just enough for the reference's
entry points and unittests - `gym.make(id)` and the `gym.Env` base class (/root/reference/AsyncTools/AsyncPPO.py:3,35,
train.py:8, unittests/test_AsyncPPO.py:43,56).  `make` hands out the B200 build's env descriptor (prl_b200.make): the physics lives
in the CUDA kernels, not on the host.  Test infrastructure only; put this directory on PYTHONPATH to use it."""
from prl_b200.envs import make  # noqa: F401


class Env:
    pass
