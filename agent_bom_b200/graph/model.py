"""Host-side node / edge / path records of the exposure graph.

These are the light records the host keeps per node and per edge; the device
only ever sees dense integer codes derived from them (``csr.py``).  Field names
follow the reference records so either object family can be fed to the engine
(duck-typed): ``UnifiedNode`` (``/root/reference/src/agent_bom/graph/node.py:74-123``),
``UnifiedEdge`` (``graph/edge.py:13-60``), ``AttackPath``
(``graph/container.py:18-57``).
"""

from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Any

from .schema import EntityType, NodeStatus, RelationshipType, enum_value


@dataclass(slots=True)
class UnifiedNode:
    id: str
    entity_type: EntityType
    label: str
    status: NodeStatus = NodeStatus.ACTIVE
    risk_score: float = 0.0
    severity: str = ""
    first_seen: str = ""
    last_seen: str = ""
    attributes: dict[str, Any] = field(default_factory=dict)
    compliance_tags: list[str] = field(default_factory=list)
    data_sources: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, Any]:
        d = asdict(self)
        d["entity_type"] = enum_value(self.entity_type)
        d["status"] = enum_value(self.status)
        return d


@dataclass(slots=True)
class UnifiedEdge:
    source: str
    target: str
    relationship: RelationshipType
    direction: str = "directed"  # "directed" | "bidirectional"
    weight: float = 1.0
    traversable: bool = True
    evidence: dict[str, Any] = field(default_factory=dict)

    @property
    def is_bidirectional(self) -> bool:
        return self.direction == "bidirectional"

    @property
    def id(self) -> str:
        return f"{enum_value(self.relationship)}:{self.source}:{self.target}"

    def reversed_copy(self) -> "UnifiedEdge":
        """The twin a bidirectional edge contributes to the other endpoint's adjacency."""
        return UnifiedEdge(
            source=self.target,
            target=self.source,
            relationship=self.relationship,
            direction=self.direction,
            weight=self.weight,
            traversable=self.traversable,
            evidence=self.evidence,
        )

    def to_dict(self) -> dict[str, Any]:
        return {
            "source": self.source,
            "target": self.target,
            "relationship": enum_value(self.relationship),
            "direction": self.direction,
            "weight": self.weight,
            "traversable": self.traversable,
            "evidence": self.evidence,
        }


@dataclass(slots=True)
class AttackPath:
    """One exposure path: agent → server → (package) → finding, with the server's cred/tool fan-out."""

    source: str
    target: str
    hops: list[str] = field(default_factory=list)
    edges: list[str] = field(default_factory=list)
    composite_risk: float = 0.0
    summary: str = ""
    credential_exposure: list[str] = field(default_factory=list)
    tool_exposure: list[str] = field(default_factory=list)
    vuln_ids: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, Any]:
        return asdict(self)

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "AttackPath":
        return cls(
            source=data["source"],
            target=data["target"],
            **{k: data[k] for k in ("hops", "edges", "composite_risk", "summary", "credential_exposure", "tool_exposure", "vuln_ids") if k in data},
        )
