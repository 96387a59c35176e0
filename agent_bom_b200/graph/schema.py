"""Typed-graph vocabulary: entity kinds, relationship kinds and their dense device codes.

The exposure graph is typed; the device CSR stores one ``uint8`` entity code per
node and a 5-bit relationship code per adjacency entry, so the *ordinal* of each
enum member below is part of the on-device data format.

Mirrors the value tables of the reference schema
(``/root/reference/src/agent_bom/graph/types.py:8-43`` entity kinds, ``:63-111``
relationship kinds, ``graph/ocsf.py:49-54`` finding kinds,
``graph/container.py:777`` dynamic relationships,
``graph/severity.py:49-58`` severity ranks).  Members compare equal to their
string values (``str`` mix-in) so reference objects and ours interoperate.
"""

from __future__ import annotations

from enum import Enum

# Ordinal == device code.  Order is the reference's declaration order.
_ENTITY_VALUES = (
    "agent server package tool model dataset container cloud_resource "
    "vulnerability misconfiguration credential "
    "org account user group role policy service_account service_principal federated_identity "
    "provider environment fleet cluster"
).split()

_RELATIONSHIP_VALUES = (
    "hosts uses depends_on provides_tool exposes_cred reaches_tool serves_model contains "
    "affects vulnerable_to exploitable_via remediates triggers "
    "shares_server shares_cred lateral_path "
    "manages owns part_of member_of assumes trusts attached inherits can_access cross_account_trust "
    "invoked accessed delegated_to "
    "correlates_with possibly_correlates_with"
).split()

EntityType = Enum("EntityType", [(v.upper(), v) for v in _ENTITY_VALUES], type=str, module=__name__)
RelationshipType = Enum("RelationshipType", [(v.upper(), v) for v in _RELATIONSHIP_VALUES], type=str, module=__name__)
NodeStatus = Enum("NodeStatus", [(v.upper(), v) for v in ("active", "inactive", "vulnerable", "remediated")], type=str, module=__name__)

N_ENTITY_TYPES = len(_ENTITY_VALUES)  # 24
N_RELATIONSHIP_TYPES = len(_RELATIONSHIP_VALUES)  # 31
assert N_ENTITY_TYPES == 24 and N_RELATIONSHIP_TYPES == 31

#: device code for an edge endpoint that has no node record ("ghost")
ENTITY_CODE_GHOST = 255
#: device code for a relationship string outside the enum (never matches a typed mask)
REL_CODE_OTHER = 31

ENTITY_CODE = {v: i for i, v in enumerate(_ENTITY_VALUES)}
REL_CODE = {v: i for i, v in enumerate(_RELATIONSHIP_VALUES)}
ENTITY_VALUES = tuple(_ENTITY_VALUES)
RELATIONSHIP_VALUES = tuple(_RELATIONSHIP_VALUES)

FINDING_ENTITY_TYPES = frozenset({EntityType.VULNERABILITY, EntityType.MISCONFIGURATION})
FINDING_CODES = (ENTITY_CODE["vulnerability"], ENTITY_CODE["misconfiguration"])

DYNAMIC_RELS = frozenset({RelationshipType.INVOKED, RelationshipType.ACCESSED, RelationshipType.DELEGATED_TO})

SEVERITY_RANK = {
    "critical": 5,
    "high": 4,
    "medium": 3,
    "low": 2,
    "info": 1,
    "informational": 1,
    "none": 0,
    "unknown": 0,
}

# Adjacency-entry meta byte layout (device format)
META_REL_MASK = 0x1F
META_TRAVERSABLE = 0x20
META_FIRST_PAIR = 0x40
META_REVERSED_COPY = 0x80

ALL_RELS_MASK = 0xFFFFFFFF


def enum_value(x) -> str:
    """``x.value`` for enum members, ``str(x)`` otherwise (the reference accepts raw strings)."""
    return x.value if isinstance(x, Enum) else str(x)


def entity_code(entity_type) -> int:
    return ENTITY_CODE.get(enum_value(entity_type), ENTITY_CODE_GHOST - 1)


def rel_code(relationship) -> int:
    return REL_CODE.get(enum_value(relationship), REL_CODE_OTHER)


def rel_mask(relationships) -> int:
    """32-bit mask for a set of relationships; empty/None means "all" (reference: falsy set = no filter)."""
    if not relationships:
        return ALL_RELS_MASK
    m = 0
    for r in relationships:
        code = REL_CODE.get(enum_value(r))
        if code is not None:
            m |= 1 << code
    return m


DYNAMIC_MASK = rel_mask(DYNAMIC_RELS)
#: dependency-reach walk mask (reference: graph/dependency_reach.py:44-51)
REACH_MASK = rel_mask({RelationshipType.USES, RelationshipType.DEPENDS_ON, RelationshipType.CONTAINS, RelationshipType.PROVIDES_TOOL})
#: vulnerability→package attachment mask (reference: graph/dependency_reach.py:54-59)
VULN_PKG_MASK = rel_mask({RelationshipType.AFFECTS, RelationshipType.VULNERABLE_TO})
