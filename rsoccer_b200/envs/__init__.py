"""Batched envs.  `make(id, num_envs=...)` mirrors gym.make for the ids the reference
registers (rsoccer_gym/__init__.py:3-30) that are in scope (BASELINE.json configs)."""
from .base import BoxSpec, SSLBaseVecEnv, VSSBaseVecEnv
from .fused import SSLContestedPossessionVecEnv, SSLStaticDefendersVecEnv, VSSVecEnv

REGISTRY = {
    "VSS-v0": VSSVecEnv,
    "SSLStaticDefenders-v0": SSLStaticDefendersVecEnv,
    "SSLContestedPossession-v0": SSLContestedPossessionVecEnv,
}


def make(id, num_envs=1, **kwargs):
    if id not in REGISTRY:
        raise KeyError("%r is not a batched env of rsoccer_b200 (have: %s)" % (id, ", ".join(sorted(REGISTRY))))
    return REGISTRY[id](num_envs=num_envs, **kwargs)


__all__ = ["make", "REGISTRY", "BoxSpec", "VSSBaseVecEnv", "SSLBaseVecEnv", "VSSVecEnv",
           "SSLStaticDefendersVecEnv", "SSLContestedPossessionVecEnv"]
