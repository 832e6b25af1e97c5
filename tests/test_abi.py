"""CPU: the C-ABI library loads and exports every symbol include/ahrag_b200.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "ahrag_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ahr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ah_rag_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms, "ctypes binding and header disagree"
    assert lib.ahr_version() == 200
    assert lib.ahr_launch_count() >= 0


def test_rerank_params_struct_layout_matches_header():
    from ah_rag_b200 import _lib
    # 10 doubles + 6 int32 = 104 bytes, 8-byte aligned, no implicit padding
    assert ctypes.sizeof(_lib.RerankParams) == 10 * 8 + 6 * 4
    assert _lib.RerankParams.boost.offset == 32 and _lib.RerankParams.has_judge_min.offset == 80


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ah_rag_b200 import _lib, engine
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.ahr_index_create(ctypes.byref(h), 0, 10, 64, 1, ctypes.c_void_p(256), 0)
    assert rc != 0 and lib.ahr_last_error()  # no silent CPU path
    with pytest.raises(RuntimeError):
        engine.VectorIndex(torch.zeros((4, 64), dtype=torch.bfloat16))  # host tensor refused
    with pytest.raises(RuntimeError):
        engine.GraphCSR({"n_nodes": 0}, device="cpu")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no module of the package may import it."""
    pkg = os.path.join(ROOT, "ah_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, fn)


def test_env_batch_struct_mirrors_header():
    """ctypes mirror of ``struct ahr_env_batch``: same field names in the same order, 4 int32 + 17 pointers."""
    import ctypes
    import re
    from ah_rag_b200 import _lib
    with open(os.path.join(ROOT, "include", "ahrag_b200.h"), "r", encoding="utf-8") as f:
        text = f.read()
    body = re.search(r"typedef struct ahr_env_batch \{(.*?)\} ahr_env_batch;", text, re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = re.sub(r"/\*.*?\*/", "", decl, flags=re.S).strip()
        if not decl:
            continue
        _, rest = decl.split(None, 1)
        names += [n.strip().lstrip("*") for n in rest.split(",")]
    assert names == [n for n, _ in _lib.EnvBatch._fields_]
    assert ctypes.sizeof(_lib.EnvBatch) == 4 * 4 + 17 * ctypes.sizeof(ctypes.c_void_p)
