"""B200-native batched simulator for the SSD grid worlds (Cleanup / Harvest) of drdh/Homophily-MARL.

Public surface:
  * ``SSDBatchEnv``      -- batched tensor API over B HBM-resident env instances
  * ``REGISTRY``         -- PyMARL ``MultiAgentEnv`` constructors ('cleanup', 'harvest')
  * ``mapspec``          -- host-side map compiler (ASCII map -> device tables)
The compute path is the CUDA library ``libssd_b200.so`` (C ABI: include/ssd_b200.h).
"""
from . import mapspec  # noqa: F401

__all__ = ["mapspec", "SSDBatchEnv", "BatchedEpisodeRunner", "incentive_rewards", "REGISTRY", "CleanupEnv", "HarvestEnv",
           "MultiAgentEnv"]


def __getattr__(name):
    if name == "SSDBatchEnv":
        from .batch_env import SSDBatchEnv
        return SSDBatchEnv
    if name == "BatchedEpisodeRunner":
        from .batched_runner import BatchedEpisodeRunner
        return BatchedEpisodeRunner
    if name == "incentive_rewards":
        from .incentive import incentive_rewards
        return incentive_rewards
    if name in ("REGISTRY", "CleanupEnv", "HarvestEnv", "MultiAgentEnv"):
        from . import pymarl_env
        return getattr(pymarl_env, name)
    raise AttributeError(name)
