"""Host-side profile of one HierarchicalGraph.search call on the c1 fixture (B = 1): where the latency goes."""
import copy, cProfile, os, pstats, sys, tempfile, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ah_rag_b200 import synth
from ah_rag_b200.graph.hierarchical_graph import HierarchicalGraph
from tests import golden_util as gu

fx = gu.load("g_c1_hotpotqa_mini")
tmp = tempfile.mkdtemp()
gdir, db = os.path.join(tmp, "graph"), os.path.join(tmp, "db")
nl = copy.deepcopy(fx["node_link"]); nl["graph"]["vector_index"]["db_path"] = db
synth.write_graph_dir(gdir, nl)
hg = HierarchicalGraph.load(gdir)
qmap = {q: fx["query_vecs"][i] for i, q in enumerate(fx["queries"])}
hg.set_embedder(lambda texts: np.stack([qmap[t] for t in texts]))
hg.build_vector_index(db, layers={0, 1, 2}, reset=True, embeddings={n: fx["index_rows"][r] for r, n in enumerate(fx["index_ids"])})
for _ in range(20):
    hg.search("Tim Burton", top_k=3)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    hg.search("Tim Burton", top_k=3)
print("search B=1: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
t0 = time.perf_counter()
for _ in range(50):
    hg.search_batch(fx["query_vecs"], top_k=5)
print("search_batch B=64: %.1f us" % ((time.perf_counter() - t0) / 50 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    hg.search("Tim Burton", top_k=3)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
