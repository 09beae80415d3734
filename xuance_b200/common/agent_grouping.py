"""Agent -> parameter-group assignment (reference: xuance/common/agent_grouping.py:4-92), the subset the QMIX path
uses: every agent in one shared group, or an explicit mapping."""
from dataclasses import dataclass
from typing import Dict, Tuple


@dataclass(frozen=True)
class AgentGrouping:
    agent_keys: Tuple[str, ...]
    assignments: Tuple[Tuple[str, str], ...]

    @property
    def agent_to_group(self) -> Dict[str, str]:
        return dict(self.assignments)

    @property
    def group_keys(self):
        m = self.agent_to_group
        return tuple(dict.fromkeys(m[a] for a in self.agent_keys))

    @property
    def groups(self):
        m = self.agent_to_group
        return {g: tuple(a for a in self.agent_keys if m[a] == g) for g in self.group_keys}

    def agent_indices(self, group_key):
        members = set(self.groups[group_key])
        return tuple(i for i, a in enumerate(self.agent_keys) if a in members)

    @property
    def full_shared(self):
        return len(self.group_keys) == 1

    @classmethod
    def shared(cls, agent_keys):
        agent_keys = tuple(agent_keys)
        return cls(agent_keys=agent_keys, assignments=tuple((a, "shared") for a in agent_keys))
