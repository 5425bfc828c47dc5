from dataclasses import dataclass, field
from typing import Any, Dict, Optional

registry: Dict[str, "EnvSpec"] = {}


@dataclass
class EnvSpec:
    id: str
    entry_point: Any
    max_episode_steps: Optional[int] = None
    kwargs: Dict[str, Any] = field(default_factory=dict)


def register(id, entry_point, max_episode_steps=None, kwargs=None, **_):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, dict(kwargs or {}))
