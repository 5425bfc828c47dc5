"""Batched envs.  `make(id, num_envs=...)` mirrors gym.make for the five ids the reference
registers (rsoccer_gym/__init__.py:3-30)."""
from .base import BoxSpec, SSLBaseVecEnv, VSSBaseVecEnv
from .fused import (SSLContestedPossessionVecEnv, SSLDribblingVecEnv, SSLPassEnduranceVecEnv,
                    SSLStaticDefendersVecEnv, VSSVecEnv)

REGISTRY = {
    "VSS-v0": VSSVecEnv,
    "SSLStaticDefenders-v0": SSLStaticDefendersVecEnv,
    "SSLContestedPossession-v0": SSLContestedPossessionVecEnv,
    "SSLDribbling-v0": SSLDribblingVecEnv,
    "SSLPassEndurance-v0": SSLPassEnduranceVecEnv,
}


def make(id, num_envs=1, **kwargs):
    if id not in REGISTRY:
        raise KeyError("%r is not a batched env of rsoccer_b200 (have: %s)" % (id, ", ".join(sorted(REGISTRY))))
    return REGISTRY[id](num_envs=num_envs, **kwargs)


__all__ = ["make", "REGISTRY", "BoxSpec", "VSSBaseVecEnv", "SSLBaseVecEnv", "VSSVecEnv",
           "SSLStaticDefendersVecEnv", "SSLContestedPossessionVecEnv", "SSLDribblingVecEnv",
           "SSLPassEnduranceVecEnv"]
