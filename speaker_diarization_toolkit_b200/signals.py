"""Host-side mirror of speaker-assign's signal collection and fusion.

  * constants                      speaker-assign:49-70
  * `Signal`, `Assignment`         speaker-assign:249-258, :406-415
  * `signals_from_matches`         the body of collect_embedding_signals, speaker-assign:296-324
                                   (min-trust filter :304-311, default score 0.5 :316) -- plus the per-label
                                   filter the reference lacks (SURVEY 8b "per-label pitfall": rows carrying a
                                   `label` other than the one being assigned are skipped)
  * `collect_context_signals`      speaker-assign:331-353
  * `combine_signals`              speaker-assign:418-492 (float64, stable sort, bands, threshold)
Pinned by tests/golden/combine_signals_golden.json and embedding_signals_golden.json (reference outputs).
On the GPU the embedding-only case of combine_signals is `sdk_assign` (csrc/select.cu k_assign); this module is
the general case (context / LLM signals mixed in) and runs per label in Python like the reference.
"""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional

SIGNAL_WEIGHTS = {
    "embedding_match": 0.4,
    "llm_name_detection": 0.3,
    "context_expected": 0.2,
    "cross_backend_agreement": 0.1,
}
TRUST_MULTIPLIERS = {"high": 1.0, "medium": 0.7, "low": 0.4, "invalidated": 0.0, "unknown": 0.5}
CONFIDENCE_THRESHOLDS = {"high": 0.7, "medium": 0.4, "low": 0.2}
_TRUST_ORDER = ["low", "medium", "high"]


@dataclass
class Signal:
    type: str
    speaker_id: Optional[str]
    score: float
    evidence: dict = field(default_factory=dict)


@dataclass
class Assignment:
    speaker_label: str
    speaker_id: Optional[str]
    confidence: str
    score: float
    signals: List[dict]
    candidates: List[dict] = field(default_factory=list)


def signals_from_matches(matches: Any, min_trust: str = "low", label: Optional[str] = None) -> List[Signal]:
    """identify's JSON rows -> embedding_match signals.  `label` (new): keep only rows of that diarization label;
    rows without a `label` key are whole-recording rows and are kept (reference behaviour)."""
    out: List[Signal] = []
    if not isinstance(matches, list):
        return out
    for row in matches:
        if not row.get("speaker_id"):
            continue
        if label is not None and row.get("label") is not None and row["label"] != label:
            continue
        trust = row.get("trust_level", "unknown")
        if min_trust in _TRUST_ORDER and trust in _TRUST_ORDER:
            if _TRUST_ORDER.index(trust) < _TRUST_ORDER.index(min_trust):
                continue
        out.append(Signal("embedding_match", row["speaker_id"], row.get("score", 0.5),
                          {"embedding_id": row.get("embedding_id"), "trust_level": trust, "backend": row.get("backend")}))
    return out


def collect_context_signals(speaker_label: str, context_name: Optional[str], expected_speakers: Iterable[str]) -> List[Signal]:
    return [Signal("context_expected", sid, 0.5, {"context": context_name, "reason": "in expected_speakers list"})
            for sid in expected_speakers]


def combine_signals(speaker_label: str, signals: List[Signal], threshold: float = 0.5) -> Assignment:
    totals: Dict[str, float] = defaultdict(float)
    proof: Dict[str, list] = defaultdict(list)
    for sig in signals:
        if sig.speaker_id is None:
            continue
        weight = SIGNAL_WEIGHTS.get(sig.type, 0.1)
        if sig.type == "embedding_match":
            weight *= TRUST_MULTIPLIERS.get(sig.evidence.get("trust_level", "unknown"), 0.5)
        totals[sig.speaker_id] += weight * sig.score
        proof[sig.speaker_id].append({"type": sig.type, "score": sig.score, **sig.evidence})
    if not totals:
        return Assignment(speaker_label, None, "unassigned", 0.0, [], [])
    ranked = sorted(totals.items(), key=lambda kv: kv[1], reverse=True)   # stable: ties keep insertion order
    top_id, top = ranked[0]
    band = "unassigned"
    for name in ("high", "medium", "low"):
        if top >= CONFIDENCE_THRESHOLDS[name]:
            band = name
            break
    as_dicts = lambda pairs: [{"speaker_id": s, "score": v} for s, v in pairs]
    if top < threshold:
        return Assignment(speaker_label, None, "unassigned", top, proof.get(top_id, []), as_dicts(ranked[:3]))
    return Assignment(speaker_label, top_id, band, top, proof.get(top_id, []), as_dicts(ranked[1:4]))
