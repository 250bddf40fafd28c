"""Import the UNMODIFIED reference boundary modules (main/graph.py, main/context.py, main/message.py) from
/root/reference for oracle pinning.  TEST INFRASTRUCTURE ONLY; build container only (/root/reference does not
exist on the GPU box — nothing in the `-m gpu` tests, smoke() or bench.py calls this).

main/context.py needs exactly one thing from Django: ``django.conf.settings.BASE_DIR`` (context.py:4,99,156);
a stub module provides it, pointing at a scratch directory that has the ``main/nodes`` and ``static/{models,
graphs}`` sub-directories ``scan_nodes`` lists at import time (context.py:176).  The reference's own cos node
(main/nodes/cos.py) is linked into the scratch tree so that plugin discovery is exercised with a real plugin.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("VITB200_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main", "context.py"))


_loaded = None


def load():
    """Returns (graph_module, context_module, message_module, base_dir) of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    base = tempfile.mkdtemp(prefix="refhost_")
    for sub in ("main/nodes", "static/models", "static/graphs"):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    os.symlink(os.path.join(REFERENCE_ROOT, "main", "nodes", "cos.py"), os.path.join(base, "main/nodes/cos.py"))

    django = types.ModuleType("django")
    conf = types.ModuleType("django.conf")
    conf.settings = types.SimpleNamespace(BASE_DIR=base)
    django.conf = conf
    sys.modules.setdefault("django", django)
    sys.modules.setdefault("django.conf", conf)
    sys.modules["django.conf"].settings = conf.settings

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    graph = importlib.import_module("main.graph")
    context = importlib.import_module("main.context")
    message = importlib.import_module("main.message")
    _loaded = (graph, context, message, base)
    return _loaded
