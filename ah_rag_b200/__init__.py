"""ah_rag_b200 -- B200-native (sm_100a) retrieval engine behind AH-RAG's graph-search API.

Hot path only: batched query-by-corpus similarity with fused top-k (tcgen05 / TMEM / TMA), k-hop frontier
expansion with score aggregation over a device-resident CSR, node-range sharding with one candidate
all-gather.  Host code is Python/PyTorch over a C-ABI CUDA library (include/ahrag_b200.h); there is no CPU
fallback and no other backend.

Drop-in modules (same names and signatures as the reference's ``ah_rag`` package):
    ah_rag_b200.graph.hierarchical_graph.HierarchicalGraph
    ah_rag_b200.agent.environment.GraphEnvironment
    ah_rag_b200.agent.gym_env.AHRAGGymEnv
    ah_rag_b200.baselines.naive_rag.NaiveRAG
"""
__version__ = "0.1.0"
