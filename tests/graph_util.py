"""Shared test helper: CSR arrays built directly (no NetworkX), with hub nodes and duplicates across segments."""
import numpy as np


def random_csr(rng, n_ent, n_sum, n_hyp, hub_deg):
    """CSR arrays built directly (dense ids: entities | summaries | hyperedges), with hub nodes whose lists are longer
    than many batches of the hash-set kernel and plenty of duplicates across segments."""
    n = n_ent + n_sum + n_hyp
    s0, h0 = n_ent, n_ent + n_sum

    def csr(lists):
        rp = np.zeros(n + 1, dtype=np.int64)
        rp[1:] = np.cumsum([len(l) for l in lists])
        return rp, (np.concatenate([np.asarray(l, dtype=np.int32) for l in lists]) if rp[-1] else np.zeros(0, np.int32))

    def transpose(lists):
        inv = [[] for _ in range(n)]
        for u, l in enumerate(lists):
            for v in l:
                inv[int(v)].append(u)
        return inv

    bel = [[] for _ in range(n)]
    for e in range(n_ent):
        bel[e] = list(s0 + rng.choice(n_sum, size=int(rng.integers(1, 4)), replace=False))
    for e in rng.choice(n_ent, size=min(n_ent, hub_deg), replace=False):  # summary s0 becomes a hub: hub_deg members
        if s0 not in bel[e]:
            bel[e].append(s0)
    part = [[] for _ in range(n)]
    for h in range(n_hyp):
        for e in rng.choice(n_ent, size=int(rng.integers(2, 9)), replace=False):
            part[int(e)].append(h0 + h)
    rel = [[] for _ in range(n)]
    for s_ in range(n_sum):
        rel[s0 + s_] = list(s0 + rng.choice(n_sum, size=int(rng.integers(0, min(n_sum, 40))), replace=False))
    rel[s0 + 1] = list(s0 + rng.permutation(n_sum))  # related hub
    slots = [part, transpose(part), bel, transpose(bel), rel, transpose(rel)]
    arrs = [csr(l) for l in slots]
    nt = np.zeros(n, np.uint8)
    nt[s0:h0] = 1
    nt[h0:] = 2
    return {"n_nodes": n, "row_ptr": [a[0] for a in arrs], "col_idx": [a[1] for a in arrs], "node_type": nt,
            "level": np.zeros(n, np.uint8), "judge": np.full(n, np.nan), "judge_term": np.zeros(n), "conf": np.full(n, np.nan)}
