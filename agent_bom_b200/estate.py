"""Synthetic benchmark estates: skewed agent → MCP-server → package/tool/credential inventories with findings.

Follows the *shape rules* of the reference scaffold
(``/root/reference/scripts/generate_graph_benchmark_estate.py:56-97``: 1/97 of
agents are "platform" agents with 18–32 servers, 1/23 have 7–14, 1/7 have 3–6,
the rest 1–2; 42 % of packages come from 12 popular names; versions
``{1..3}.{0..9}.{0..16}``; five ecosystems) and adds the density knobs
SURVEY.md §8(d) calls for, using only report fields the reference graph
builder reads (``graph/builder.py:41-507``):

* ``creds_per_server`` credential env vars drawn from a name bucket shared by
  ``cred_bucket`` consecutive agents  → EXPOSES_CRED, REACHES_TOOL (creds×tools),
  SHARES_CRED cliques (builder.py:331-370, 493-507);
* ``vulns_per_server`` package-level ``vulnerabilities[]`` on the first packages
  of every server → pkg→vuln VULNERABLE_TO (builder.py:1059-1067);
* declared tool ``capabilities`` → EXPLOITABLE_VIA vuln×tools (builder.py:971-1019);
* ``ecosystem`` on blast-radius rows so pkg→vuln resolves (builder.py:419-430).

Two outputs from ONE deterministic description:

* ``generate(...)`` → ``Estate``: the graph *as the reference builder would build
  it* — node order = ``graph.nodes`` insertion order, edge order = ``graph.edges``
  order, first-wins de-duplication — as flat integer arrays, vectorised (numpy)
  so a 10 M-node / 100 M-edge estate is generated in tens of seconds;
* ``Estate.report_json()`` → the scanner-style report the reference ingests.
  ``tests/test_estate_identity.py`` pins (ids, types, edges) of the direct path
  against the reference builder's output for that JSON (golden fixture made by
  ``oracle/make_golden.py``).

Randomness is counter-based (splitmix64 of (seed, stream, index)), so every
quantity is a pure function of its indices — no sequential RNG stream.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .graph.schema import ENTITY_CODE, REL_CODE

SOURCES = ("local", "github-action", "k8s-fleet", "operator-push", "cloud-inventory")
AGENT_TYPES = ("claude-desktop", "cursor", "windsurf", "vscode", "cortex-code")
ECOSYSTEMS = ("npm", "pypi", "go", "maven", "oci")
SEVERITIES = ("critical", "high", "medium", "low")
POPULAR_PACKAGES = ("langchain", "openai", "anthropic", "mcp-sdk", "fastapi", "requests", "zod", "react", "next", "protobuf", "grpc", "boto3")

ET = ENTITY_CODE
R = REL_CODE
FLAG_T, FLAG_B = 1, 2

# node-key kinds (high bits of the int64 identity key)
K_PROV, K_AGENT, K_SERVER, K_PKG, K_TOOL, K_CRED, K_VULN27, K_VULN26 = range(8)
_KSHIFT = 58


@dataclass(frozen=True)
class Knobs:
    """Density knobs; the defaults are the shipped scaffold's shape (|E|/|V| ≈ 1.2)."""

    creds_per_server: int = 0      # 0 = shipped rule: every 9th agent has one private token
    cred_bucket: int = 1           # agents sharing one credential-name bucket
    vulns_per_server: int = 0      # package-level vulnerabilities on the first packages of each server
    tool_capabilities: bool = False
    blast_ecosystem: bool = False  # fill `ecosystem` on blast-radius rows
    vulnerable_package_rate: float = 0.08

    @classmethod
    def dense(cls, c: int = 16, b: int = 64, v: int = 8) -> "Knobs":
        return cls(creds_per_server=c, cred_bucket=b, vulns_per_server=v, tool_capabilities=True, blast_ecosystem=True)


#: the benchmark estate of BASELINE.json's configs (≈58 nodes and ≈10 adjacency entries per node)
BENCH_KNOBS = Knobs.dense(20, 80, 8)


def _mix(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _u01(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = _mix(_mix(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream)) ^ idx.astype(np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _randint(seed: int, stream: int, idx: np.ndarray, lo, hi) -> np.ndarray:
    """Inclusive integer range, elementwise lo/hi allowed."""
    u = _u01(seed, stream, idx)
    return (np.floor(u * (np.asarray(hi) - np.asarray(lo) + 1)) + lo).astype(np.int64)


@dataclass
class Estate:
    agents: int
    seed: int
    knobs: Knobs
    node_key: np.ndarray      # int64 identity key per node (kind << 58 | fields)
    node_type: np.ndarray     # uint8 entity code
    node_rank: np.ndarray     # int32 order of the id strings (exact when strings were materialised, else kind/field order)
    node_sev: np.ndarray      # int8 severity index into SEVERITIES for findings, -1 otherwise
    src: np.ndarray
    dst: np.ndarray
    rel: np.ndarray
    flags: np.ndarray
    findings: np.ndarray      # int32 node indices of vulnerability nodes (node order)
    agent_nodes: np.ndarray   # int32 node indices of agent nodes
    layout: dict = field(default_factory=dict, repr=False)

    @property
    def n_nodes(self) -> int:
        return int(self.node_type.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.src.shape[0])

    def summary(self) -> dict:
        nb = int((self.flags & FLAG_B).astype(bool).sum())
        return {
            "agents": self.agents, "seed": self.seed, "knobs": self.knobs.__dict__, "nodes": self.n_nodes, "edges": self.n_edges,
            "adjacency_entries_per_direction": self.n_edges + nb, "findings": int(self.findings.shape[0]),
            "edges_by_relationship": {name: int(c) for name, c in zip(REL_CODE, np.bincount(self.rel, minlength=31)) if c},
        }

    # ── string materialisation (small estates / identity tests) ─────────
    def node_ids(self) -> list[str]:
        L = self.layout
        out = []
        for key in self.node_key.tolist():
            out.append(_key_to_id(key, L))
        return out

    def with_exact_rank(self) -> "Estate":
        ids = self.node_ids()
        order = sorted(range(len(ids)), key=ids.__getitem__)
        rank = np.empty(len(ids), dtype=np.int32)
        rank[np.asarray(order, dtype=np.int64)] = np.arange(len(ids), dtype=np.int32)
        self.node_rank = rank
        return self

    def report_json(self) -> dict:
        """Scanner-style report for the reference builder (fields it reads only)."""
        return _report_json(self)


# ── naming (must stay in sync with _report_json) ────────────────────────────

def _agent_name(ai: int) -> str:
    return f"agent-{ai:07d}"


def _server_name(ai: int, si: int) -> str:
    return f"{_agent_name(ai)}-mcp-{si:02d}"


def _pkg_fields(key: int):
    body = key & ((1 << _KSHIFT) - 1)
    nk = body & 3; body >>= 2
    eco = body & 7; body >>= 3
    v1 = body & 3; body >>= 2
    v2 = body & 15; body >>= 4
    v3 = body & 31; body >>= 5
    if nk == 0:
        name = POPULAR_PACKAGES[body]
    elif nk == 1:
        name = f"team-shared-{body >> 4}-{body & 15}"
    else:
        pi = body & 31; si = (body >> 5) & 63; ai = body >> 11
        name = f"svc-{ai:07d}-{si:02d}-{pi:02d}"
    return ECOSYSTEMS[eco], name, f"{v1}.{v2}.{v3}"


def _key_to_id(key: int, L: dict) -> str:
    kind = key >> _KSHIFT
    body = key & ((1 << _KSHIFT) - 1)
    if kind == K_PROV:
        return f"provider:{SOURCES[body]}"
    if kind == K_AGENT:
        return f"agent:{_agent_name(body)}"
    if kind == K_SERVER:
        ai, si = body >> 6, body & 63
        return f"server:{_agent_name(ai)}:{_server_name(ai, si)}"
    if kind == K_PKG:
        eco, name, ver = _pkg_fields(key)
        return f"pkg:{eco}:{name}@{ver}"
    if kind == K_TOOL:
        ti = body & 63; si = (body >> 6) & 63; ai = body >> 12
        return f"tool:server:{_agent_name(ai)}:{_server_name(ai, si)}:{_server_name(ai, si)}-tool-{ti:02d}"
    if kind == K_CRED:
        if body & 1:
            return f"cred:{_agent_name(body >> 1).upper().replace('-', '_')}_TOKEN"
        k = (body >> 1) & 63; bucket = body >> 7
        return f"cred:TEAM{bucket:07d}_TOKEN_{k:02d}"
    if kind == K_VULN27:
        pi = body & 31; si = (body >> 5) & 63; ai = body >> 11
        return f"vuln:CVE-2027-{ai:07d}{si:02d}{pi:02d}"
    if kind == K_VULN26:
        return f"vuln:CVE-2026-{body + 1:07d}"
    raise ValueError(kind)


def _mk(kind: int, body) -> np.ndarray:
    return (np.int64(kind) << np.int64(_KSHIFT)) | np.asarray(body, dtype=np.int64)


def generate(agents: int, seed: int = 2145, knobs: Knobs = Knobs(), exact_rank: bool | None = None) -> Estate:
    """Build the estate's graph directly in the reference builder's node / edge order."""
    A = int(agents)
    kn = knobs
    ai = np.arange(A, dtype=np.int64)

    # ---- per-agent shape (scaffold :56-63)
    u_srv = _u01(seed, 1, ai)
    n_srv = np.where(ai % 97 == 0, _randint(seed, 2, ai, 18, 32),
             np.where(ai % 23 == 0, _randint(seed, 3, ai, 7, 14),
              np.where(ai % 7 == 0, _randint(seed, 4, ai, 3, 6), np.where(u_srv < 0.68, 1, 2)))).astype(np.int64)
    S = int(n_srv.sum())
    srv_agent = np.repeat(ai, n_srv)
    srv_first = np.zeros(A + 1, dtype=np.int64); srv_first[1:] = np.cumsum(n_srv)
    srv_si = np.arange(S, dtype=np.int64) - srv_first[srv_agent]
    sidx = np.arange(S, dtype=np.int64)

    # ---- per-server shape (scaffold :66-79)
    n_tool = np.where((srv_agent % 97 == 0) & (srv_si < 4), _randint(seed, 5, sidx, 20, 45),
              np.where(srv_agent % 23 == 0, _randint(seed, 6, sidx, 8, 18), _randint(seed, 7, sidx, 1, 5))).astype(np.int64)
    n_pkg = np.where(srv_agent % 97 == 0, _randint(seed, 8, sidx, 16, 28),
             np.where(srv_si % 5 == 0, _randint(seed, 9, sidx, 8, 14), _randint(seed, 10, sidx, 3, 8))).astype(np.int64)
    n_vul = np.minimum(n_pkg, kn.vulns_per_server)
    if kn.creds_per_server > 0:
        n_cred = np.full(S, kn.creds_per_server, dtype=np.int64)
    else:
        n_cred = (srv_agent % 9 == 0).astype(np.int64)

    # ---- package instances
    P = int(n_pkg.sum())
    pk_srv = np.repeat(sidx, n_pkg)
    pk_first = np.zeros(S + 1, dtype=np.int64); pk_first[1:] = np.cumsum(n_pkg)
    pk_pi = np.arange(P, dtype=np.int64) - pk_first[pk_srv]
    pk_ai, pk_si = srv_agent[pk_srv], srv_si[pk_srv]
    pidx = np.arange(P, dtype=np.int64)
    u1, u2 = _u01(seed, 11, pidx), _u01(seed, 12, pidx)
    nk = np.where(u1 < 0.42, 0, np.where(u2 < 0.18, 1, 2)).astype(np.int64)
    name_body = np.where(nk == 0, (pk_ai + pk_si + pk_pi) % 12,
                 np.where(nk == 1, (((pk_ai // 25) % 25) << 4) | (pk_pi % 9), (pk_ai << 11) | (pk_si << 5) | pk_pi))
    eco = (pk_ai + pk_si + pk_pi) % 5
    v1, v2, v3 = 1 + pk_pi % 3, pk_si % 10, pk_ai % 17
    pkg_body = ((((((name_body << 5) | v3) << 4 | v2) << 2 | v1) << 3 | eco) << 2) | nk
    pkg_key = _mk(K_PKG, pkg_body)
    name_key = (name_body << 2) | nk                      # identity of the package NAME (blast-radius rule)

    # ---- tools
    T = int(n_tool.sum())
    tl_srv = np.repeat(sidx, n_tool)
    tl_first = np.zeros(S + 1, dtype=np.int64); tl_first[1:] = np.cumsum(n_tool)
    tl_ti = np.arange(T, dtype=np.int64) - tl_first[tl_srv]
    tool_key = _mk(K_TOOL, (srv_agent[tl_srv] << 12) | (srv_si[tl_srv] << 6) | tl_ti)

    # ---- credentials (server, k)
    Cn = int(n_cred.sum())
    cr_srv = np.repeat(sidx, n_cred)
    cr_first = np.zeros(S + 1, dtype=np.int64); cr_first[1:] = np.cumsum(n_cred)
    cr_k = np.arange(Cn, dtype=np.int64) - cr_first[cr_srv]
    if kn.creds_per_server > 0:
        cred_key = _mk(K_CRED, ((srv_agent[cr_srv] // kn.cred_bucket) << 7) | (cr_k << 1))
    else:
        cred_key = _mk(K_CRED, (srv_agent[cr_srv] << 1) | 1)

    agent_key = _mk(K_AGENT, ai)
    source_idx = _agent_source(ai)
    prov_key = _mk(K_PROV, source_idx)
    srv_key = _mk(K_SERVER, (srv_agent << 6) | srv_si)
    has_v = pk_pi < n_vul[pk_srv]
    vul27_key = _mk(K_VULN27, (pk_ai << 11) | (pk_si << 5) | pk_pi)

    # ---- node timeline: position of every node MENTION inside the per-agent builder loop
    srv_nodes = 1 + n_pkg + n_vul + n_tool + n_cred
    srv_edges = 1 + n_pkg + n_vul + n_tool + n_cred * (1 + n_tool)
    ag_nodes = 2 + np.bincount(srv_agent, weights=srv_nodes, minlength=A).astype(np.int64)
    ag_edges = 1 + np.bincount(srv_agent, weights=srv_edges, minlength=A).astype(np.int64)
    ag_nbase = np.zeros(A + 1, dtype=np.int64); ag_nbase[1:] = np.cumsum(ag_nodes)
    ag_ebase = np.zeros(A + 1, dtype=np.int64); ag_ebase[1:] = np.cumsum(ag_edges)
    cs_n = np.cumsum(srv_nodes) - srv_nodes
    cs_e = np.cumsum(srv_edges) - srv_edges
    srv_nbase = ag_nbase[srv_agent] + 2 + (cs_n - cs_n[srv_first[srv_agent]])
    srv_ebase = ag_ebase[srv_agent] + 1 + (cs_e - cs_e[srv_first[srv_agent]])
    NM, EM = int(ag_nbase[-1]), int(ag_ebase[-1])

    hv = np.flatnonzero(has_v)
    ppos = 1 + pk_pi + np.minimum(pk_pi, n_vul[pk_srv])
    tbase = 1 + n_pkg + n_vul
    pos_prov, pos_agent, pos_srv = ag_nbase[:-1], ag_nbase[:-1] + 1, srv_nbase
    pos_pkg = srv_nbase[pk_srv] + ppos
    pos_v27 = pos_pkg[hv] + 1
    pos_tool = srv_nbase[tl_srv] + tbase[tl_srv] + tl_ti
    pos_cred = srv_nbase[cr_srv] + (tbase + n_tool)[cr_srv] + cr_k

    # first-wins merge (add_node): only providers, packages and credentials can be mentioned twice
    is_first = np.ones(NM, dtype=bool)

    def first_positions(keys: np.ndarray, pos: np.ndarray) -> np.ndarray:
        """timeline position of the first mention of each key, per mention; later mentions are flagged as merged."""
        order = np.argsort(pos, kind="stable")
        _, first_idx, inv = np.unique(keys[order], return_index=True, return_inverse=True)
        fp_sorted = pos[order][first_idx][inv]
        fp = np.empty_like(pos)
        fp[order] = fp_sorted
        is_first[pos[fp != pos]] = False
        return fp

    fp_prov = first_positions(prov_key, pos_prov)
    fp_pkg = first_positions(pkg_key, pos_pkg)
    fp_cred = first_positions(cred_key, pos_cred) if Cn else pos_cred
    node_of_pos = np.cumsum(is_first, dtype=np.int64) - 1           # node index of the mention at a first-occurrence position
    N0 = int(node_of_pos[-1]) + 1 if NM else 0
    ix_prov = node_of_pos[fp_prov].astype(np.int32); ix_agent = node_of_pos[pos_agent].astype(np.int32); ix_srv = node_of_pos[pos_srv].astype(np.int32)
    ix_pkg = node_of_pos[fp_pkg].astype(np.int32); ix_v27 = node_of_pos[pos_v27].astype(np.int32)
    ix_tool = node_of_pos[pos_tool].astype(np.int32); ix_cred = node_of_pos[fp_cred].astype(np.int32)

    # ---- blast-radius rows (scaffold :136-155 -> builder.py:385-471): first instance of each package NAME that drew "vulnerable"
    cand = np.flatnonzero(_u01(seed, 13, pidx) < kn.vulnerable_package_rate)
    if cand.size:
        _, first_idx = np.unique(name_key[cand], return_index=True)
        rows = cand[np.sort(first_idx)]
    else:
        rows = cand
    K = int(rows.shape[0])
    ix_v26 = (N0 + np.arange(K, dtype=np.int64)).astype(np.int32)
    N = N0 + K

    node_key = np.empty(N, dtype=np.int64); node_type = np.empty(N, dtype=np.uint8); node_sev = np.full(N, -1, dtype=np.int8)

    def set_nodes(ix, keys, etype, sev=None):
        node_key[ix] = keys; node_type[ix] = etype
        if sev is not None:
            node_sev[ix] = sev

    sev27 = ((pk_ai + pk_pi) % 4).astype(np.int8)
    set_nodes(ix_prov, prov_key, ET["provider"]); set_nodes(ix_agent, agent_key, ET["agent"]); set_nodes(ix_srv, srv_key, ET["server"])
    set_nodes(ix_pkg, pkg_key, ET["package"]); set_nodes(ix_v27, vul27_key[hv], ET["vulnerability"], sev27[hv])
    set_nodes(ix_tool, tool_key, ET["tool"]); set_nodes(ix_cred, cred_key, ET["credential"])
    set_nodes(ix_v26, _mk(K_VULN26, np.arange(K, dtype=np.int64)), ET["vulnerability"], ((pk_ai[rows] + pk_pi[rows]) % 4).astype(np.int8))

    # ---- edge timeline of the per-agent loop, written straight in node-index space
    e_s = np.zeros(EM, dtype=np.int32); e_d = np.zeros(EM, dtype=np.int32); e_r = np.zeros(EM, dtype=np.uint8)
    keep = np.ones(EM, dtype=bool)

    def put_edges(pos, s, d, r):
        e_s[pos] = s; e_d[pos] = d; e_r[pos] = r

    put_edges(ag_ebase[:-1], ix_prov, ix_agent, R["hosts"])
    put_edges(srv_ebase, ix_agent[srv_agent], ix_srv, R["uses"])
    dep_pos = srv_ebase[pk_srv] + ppos
    put_edges(dep_pos, ix_srv[pk_srv], ix_pkg, R["depends_on"])
    # add_edge drops a repeated (server, package, depends_on): same package id listed twice by one server
    dkey = pk_srv * np.int64(N) + ix_pkg
    _, dfirst = np.unique(dkey, return_index=True)
    dup = np.ones(P, dtype=bool); dup[dfirst] = False
    keep[dep_pos[dup]] = False
    put_edges(dep_pos[hv] + 1, ix_pkg[hv], ix_v27, R["vulnerable_to"])
    put_edges(srv_ebase[tl_srv] + tbase[tl_srv] + tl_ti, ix_srv[tl_srv], ix_tool, R["provides_tool"])
    cpos = srv_ebase[cr_srv] + (tbase + n_tool)[cr_srv] + cr_k * (1 + n_tool[cr_srv])
    put_edges(cpos, ix_srv[cr_srv], ix_cred, R["exposes_cred"])
    # REACHES_TOOL: every credential of a server x every tool of that server
    nt_c = n_tool[cr_srv]
    rt_cred = np.repeat(np.arange(Cn, dtype=np.int64), nt_c)
    rt_first = np.zeros(Cn + 1, dtype=np.int64); rt_first[1:] = np.cumsum(nt_c)
    rt_ti = np.arange(int(rt_first[-1]), dtype=np.int64) - rt_first[rt_cred]
    put_edges(cpos[rt_cred] + 1 + rt_ti, ix_cred[rt_cred], ix_tool[tl_first[cr_srv[rt_cred]] + rt_ti], R["reaches_tool"])
    del rt_cred, rt_ti

    seg_s = [e_s[keep]]; seg_d = [e_d[keep]]; seg_r = [e_r[keep]]
    del e_s, e_d, e_r, keep

    def tools_of(servers: np.ndarray):
        """(repeat index, tool node indices) for the tools of each listed server, in tool order."""
        cnt = n_tool[servers]
        rep = np.repeat(np.arange(servers.shape[0], dtype=np.int64), cnt)
        first = np.zeros(servers.shape[0] + 1, dtype=np.int64); first[1:] = np.cumsum(cnt)
        ti = np.arange(int(first[-1]), dtype=np.int64) - first[rep]
        return rep, first, ix_tool[tl_first[servers[rep]] + ti]

    # ---- pending EXPLOITABLE_VIA edges of package-level vulnerabilities (builder.py:371-382, 971-1019)
    if kn.tool_capabilities and hv.size:
        rep, _, tk = tools_of(pk_srv[hv])
        seg_s.append(ix_v27[rep]); seg_d.append(tk); seg_r.append(np.full(tk.shape[0], R["exploitable_via"], dtype=np.uint8))

    if K:
        if kn.tool_capabilities and kn.blast_ecosystem:
            rep, first, tk = tools_of(pk_srv[rows])
            nt_rows = n_tool[pk_srv[rows]]
        else:
            rep = np.zeros(0, dtype=np.int64); tk = np.zeros(0, dtype=np.int32); nt_rows = np.zeros(K, dtype=np.int64); first = None
        per_row = (1 if kn.blast_ecosystem else 0) + 1 + nt_rows
        rbase = np.cumsum(per_row) - per_row
        tot = int(per_row.sum())
        bs = np.zeros(tot, dtype=np.int32); bd = np.zeros(tot, dtype=np.int32); br = np.zeros(tot, dtype=np.uint8)
        o = 0
        if kn.blast_ecosystem:                                  # pkg -> vuln resolves only with an ecosystem (builder.py:419-430)
            bs[rbase] = ix_pkg[rows]; bd[rbase] = ix_v26; br[rbase] = R["vulnerable_to"]; o = 1
        bs[rbase + o] = ix_srv[pk_srv[rows]]; bd[rbase + o] = ix_v26; br[rbase + o] = R["vulnerable_to"]
        if rep.size:
            pos = rbase[rep] + o + 1 + (np.arange(rep.shape[0], dtype=np.int64) - first[rep])
            bs[pos] = ix_v26[rep]; bd[pos] = tk; br[pos] = R["exploitable_via"]
        seg_s.append(bs); seg_d.append(bd); seg_r.append(br)

    # ---- SHARES_CRED cliques (builder.py:493-507): all agents of a bucket share its first credential name;
    #      the later names of the bucket repeat the same (a1,a2,shares_cred) keys and are dropped by add_edge
    bidir_from = sum(int(x.shape[0]) for x in seg_s)
    if kn.creds_per_server > 0 and kn.cred_bucket > 1 and A > 1:
        b = kn.cred_bucket
        nb = (A + b - 1) // b
        sizes = np.minimum(b, A - np.arange(nb, dtype=np.int64) * b)
        i_loc, j_loc = np.triu_indices(b, k=1)                  # row-major: (a1, a2) ascending == the builder's nested loops
        sc_s, sc_d, sc_b = [], [], []
        for size in np.unique(sizes):
            buckets = np.flatnonzero(sizes == size)
            sel = j_loc < size
            ii, jj = i_loc[sel], j_loc[sel]
            sc_s.append((buckets[:, None] * b + ii[None, :]).ravel())
            sc_d.append((buckets[:, None] * b + jj[None, :]).ravel())
        sa, sd = np.concatenate(sc_s), np.concatenate(sc_d)
        if len(sc_s) > 1:                                       # the (single) short last bucket comes last in bucket order
            order = np.argsort(sa // b, kind="stable")
            sa, sd = sa[order], sd[order]
        seg_s.append(ix_agent[sa]); seg_d.append(ix_agent[sd]); seg_r.append(np.full(sa.shape[0], R["shares_cred"], dtype=np.uint8))

    src = np.concatenate(seg_s); dst = np.concatenate(seg_d); rel = np.concatenate(seg_r)
    del seg_s, seg_d, seg_r
    flags = np.full(src.shape[0], FLAG_T, dtype=np.uint8)
    flags[bidir_from:] |= FLAG_B

    # ---- id-string order surrogate: kind-major, then fields (exact among agents, whose ids are fixed-width)
    kind = node_key >> _KSHIFT
    prefix_order = np.asarray([3, 0, 4, 2, 5, 1, 6, 6], dtype=np.int64)     # agent < cred < pkg < provider < server < tool < vuln
    sort_key = (prefix_order[kind] << 60) | (node_key & ((1 << _KSHIFT) - 1))
    sort_key = np.where(kind == K_VULN27, sort_key | (1 << 59), sort_key)  # "CVE-2026" < "CVE-2027"
    rank = np.empty(N, dtype=np.int32)
    rank[np.argsort(sort_key, kind="stable")] = np.arange(N, dtype=np.int32)

    est = Estate(
        agents=A, seed=seed, knobs=kn, node_key=node_key, node_type=node_type, node_rank=rank, node_sev=node_sev,
        src=src, dst=dst, rel=rel, flags=flags,
        findings=np.flatnonzero(node_type == ET["vulnerability"]).astype(np.int32),
        agent_nodes=np.flatnonzero(node_type == ET["agent"]).astype(np.int32),
        layout={"n_srv": n_srv, "n_tool": n_tool, "n_pkg": n_pkg, "srv_first": srv_first, "pk_first": pk_first, "nk": nk, "name_body": name_body,
                "rows": rows, "pk_srv": pk_srv, "pk_pi": pk_pi, "srv_agent": srv_agent, "srv_si": srv_si},
    )
    if exact_rank or (exact_rank is None and N <= 200_000):
        est.with_exact_rank()
    return est


def _agent_source(ai: np.ndarray) -> np.ndarray:
    """Index into SOURCES of sorted({primary, +operator-push if ai%5==0, +cloud-inventory if ai%11==0})[0] (scaffold :82-89)."""
    names = np.asarray(SOURCES)
    primary = ai % 5
    best = names[primary]
    op = np.where(ai % 5 == 0, "operator-push", best)
    best = np.where(op < best, op, best)
    ci = np.where(ai % 11 == 0, "cloud-inventory", best)
    best = np.where(ci < best, ci, best)
    lut = {n: i for i, n in enumerate(SOURCES)}
    return np.asarray([lut[x] for x in best.tolist()], dtype=np.int64) if ai.shape[0] < 1_000_000 else _agent_source_fast(ai)


def _agent_source_fast(ai: np.ndarray) -> np.ndarray:
    # alphabetical order: cloud-inventory(4) < github-action(1) < k8s-fleet(2) < local(0) < operator-push(3)
    alpha = np.asarray([3, 1, 2, 4, 0], dtype=np.int64)      # rank of SOURCES[i]
    primary = ai % 5
    r = alpha[primary]
    r = np.where(ai % 5 == 0, np.minimum(r, alpha[3]), r)
    r = np.where(ai % 11 == 0, np.minimum(r, alpha[4]), r)
    inv = np.argsort(alpha)
    return inv[r]


def _report_json(est: Estate) -> dict:
    L, kn = est.layout, est.knobs
    n_srv, n_tool, n_pkg = L["n_srv"], L["n_tool"], L["n_pkg"]
    srv_first, pk_first = L["srv_first"], L["pk_first"]
    agents_out = []
    src_idx = _agent_source(np.arange(est.agents, dtype=np.int64))
    pkg_dicts: dict[int, dict] = {}
    for ai in range(est.agents):
        name = _agent_name(ai)
        servers = []
        for si in range(int(n_srv[ai])):
            s = int(srv_first[ai]) + si
            sname = _server_name(ai, si)
            pkgs = []
            for pi in range(int(n_pkg[s])):
                p = int(pk_first[s]) + pi
                nk, body = int(L["nk"][p]), int(L["name_body"][p])
                if nk == 0:
                    pname = POPULAR_PACKAGES[body]
                elif nk == 1:
                    pname = f"team-shared-{body >> 4}-{body & 15}"
                else:
                    pname = f"svc-{ai:07d}-{si:02d}-{pi:02d}"
                d = {"name": pname, "version": f"{1 + pi % 3}.{si % 10}.{ai % 17}", "ecosystem": ECOSYSTEMS[(ai + si + pi) % 5], "is_direct": pi < 2}
                if pi < min(int(n_pkg[s]), kn.vulns_per_server):
                    d["vulnerabilities"] = [{"id": f"CVE-2027-{ai:07d}{si:02d}{pi:02d}", "severity": SEVERITIES[(ai + pi) % 4]}]
                pkgs.append(d)
                pkg_dicts[p] = d
            tools = []
            for ti in range(int(n_tool[s])):
                t = {"name": f"{sname}-tool-{ti:02d}"}
                if kn.tool_capabilities:
                    t["capabilities"] = ["execute" if ti % 5 == 0 else "read"]
                tools.append(t)
            if kn.creds_per_server > 0:
                creds = [f"TEAM{ai // kn.cred_bucket:07d}_TOKEN_{k:02d}" for k in range(kn.creds_per_server)]
            else:
                creds = [f"{name.upper().replace('-', '_')}_TOKEN"] if ai % 9 == 0 else []
            servers.append({"name": sname, "transport": "sse" if si % 3 else "stdio", "surface": "mcp-server", "credential_env_vars": creds,
                            "packages": pkgs, "tools": tools})
        agents_out.append({"name": name, "type": AGENT_TYPES[ai % 5], "status": "configured", "mcp_servers": servers, "source": SOURCES[int(src_idx[ai])]})
    blast = []
    for k, p in enumerate(L["rows"].tolist()):
        s = int(L["pk_srv"][p]); ai = int(L["srv_agent"][s]); si = int(L["srv_si"][s]); pi = int(L["pk_pi"][p])
        d = pkg_dicts[p]
        row = {"vulnerability_id": f"CVE-2026-{k + 1:07d}", "severity": SEVERITIES[(ai + pi) % 4], "package": d["name"], "package_name": d["name"],
               "package_version": d["version"], "affected_agents": [_agent_name(ai)], "affected_servers": [{"name": _server_name(ai, si)}]}
        if kn.blast_ecosystem:
            row["ecosystem"] = d["ecosystem"]
        blast.append(row)
    return {"scan_id": f"b200-estate-{est.agents}-{est.seed}", "scan_sources": ["synthetic-estate"], "agents": agents_out, "blast_radius": blast}
