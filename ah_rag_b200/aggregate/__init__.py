"""Drop-in pieces of ``ah_rag.aggregate`` that sit next to the retrieval path."""
