"""dm_env-compatible TimeStep containers.

The reference's `BaseMDP.reset/step` return `dm_env.TimeStep(step_type, reward, discount, observation)`
(colosseum/mdp/base.py:1277,1316-1317).  If `dm_env` is installed its own classes are used, so results are
indistinguishable from the reference's; otherwise structurally identical stand-ins are defined here.
"""
import enum
from typing import Any, NamedTuple

try:  # pragma: no cover - depends on the environment
    from dm_env import StepType, TimeStep  # type: ignore
except Exception:  # dm_env is not part of this image

    class StepType(enum.IntEnum):
        FIRST = 0
        MID = 1
        LAST = 2

        def first(self):
            return self is StepType.FIRST

        def mid(self):
            return self is StepType.MID

        def last(self):
            return self is StepType.LAST

    class TimeStep(NamedTuple):
        step_type: Any
        reward: Any
        discount: Any
        observation: Any

        def first(self):
            return self.step_type == StepType.FIRST

        def mid(self):
            return self.step_type == StepType.MID

        def last(self):
            return self.step_type == StepType.LAST


try:  # pragma: no cover - depends on the environment
    from dm_env.specs import Array, BoundedArray, DiscreteArray  # type: ignore
except Exception:  # structurally identical stand-ins for the three spec classes the reference returns
    import numpy as _np

    class Array:
        """dm_env.specs.Array: shape, dtype, name"""

        def __init__(self, shape, dtype, name=None):
            self.shape, self.dtype, self.name = tuple(int(d) for d in shape), _np.dtype(dtype), name

        def __repr__(self):
            return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype!r}, name={self.name!r})"

    class BoundedArray(Array):
        def __init__(self, shape, dtype, minimum, maximum, name=None):
            super().__init__(shape, dtype, name)
            self.minimum, self.maximum = _np.asarray(minimum, self.dtype), _np.asarray(maximum, self.dtype)

    class DiscreteArray(BoundedArray):
        """dm_env.specs.DiscreteArray(num_values, dtype=np.int32, name): a scalar in {0, ..., num_values-1}"""

        def __init__(self, num_values, dtype=_np.int32, name=None):
            assert int(num_values) > 0
            super().__init__((), dtype, 0, int(num_values) - 1, name)
            self.num_values = int(num_values)


class BatchedTimeStep(NamedTuple):
    """N parallel TimeSteps as arrays (CUDA tensors): step_type u8[N], reward f32[N] (NaN where the reference has
    None, i.e. FIRST), discount f32[N] (1.0 MID, 0.0 LAST, NaN FIRST), observation i32[N] (-1 on LAST)."""

    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self):
        return self.step_type == int(StepType.FIRST)

    def mid(self):
        return self.step_type == int(StepType.MID)

    def last(self):
        return self.step_type == int(StepType.LAST)

    def __len__(self):  # number of envs, not number of fields
        return int(self.step_type.shape[0])

    def scalar(self, i=0) -> TimeStep:
        """the i-th env's TimeStep in exactly the reference's scalar form"""
        st = StepType(int(self.step_type[i]))
        if st == StepType.FIRST:
            return TimeStep(st, None, None, int(self.observation[i]))
        return TimeStep(st, float(self.reward[i]), 0.0 if st == StepType.LAST else 1.0, int(self.observation[i]))
