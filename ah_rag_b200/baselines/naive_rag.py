"""Drop-in for baselines/naive_rag.py:10-35: flat top-k retrieval through the graph's hybrid search, then
generation by whatever AnswerGenerator the caller supplies (LLM orchestration is out of scope)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional


class NaiveRAG:
    def __init__(self, hg: Any, answer_generator: Any) -> None:
        self.hg = hg
        self.answer_generator = answer_generator

    def retrieve(self, query: str, top_k: int = 5) -> List[str]:
        res = self.hg.search(query, top_k=top_k, return_cluster=False)  # naive_rag.py:19
        return [x.get("node_id") for x in (res or []) if x.get("node_id")]

    def _context(self, ids: List[str]) -> Dict[str, Any]:
        lines = []
        for nid in ids:
            d = self.hg.G.nodes.get(nid, {})
            title = d.get("title") or d.get("name") or ""
            body = d.get("summary_text") or d.get("summary") or d.get("description") or ""
            lines.append(f"- [{nid}] {title} :: {body[:200]}")
        return {"context_text": "\n".join(lines), "used_nodes": list(ids), "stats": {}}

    def run(self, query: str, top_k: int = 5, gen_cfg: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        return self.answer_generator.generate(query, self._context(self.retrieve(query, top_k)), gen_cfg or {})

    def run_batch(self, queries: List[str], query_embeddings, top_k: int = 5,
                  gen_cfg: Optional[Dict[str, Any]] = None) -> List[Dict[str, Any]]:
        """All questions in one fused top-k + one rerank launch (run_benchmark.py --system naive shape)."""
        batch = self.hg.search_batch(query_embeddings, top_k=top_k, return_cluster=False)
        out = []
        for q, res in zip(queries, batch):
            ids = [x.get("node_id") for x in res if x.get("node_id")]
            out.append(self.answer_generator.generate(q, self._context(ids), gen_cfg or {}))
        return out
