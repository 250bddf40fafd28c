"""Scheduler + plugin API: ``NodeKind`` / ``Model`` / ``ModelNode`` / ``Context`` / ``scan_nodes``.

Mirror of the reference's main/context.py:16-176 with the same names, argument meaning and error behaviour.
The one deliberate difference: the reference reads ``django.conf.settings.BASE_DIR`` (context.py:4,99,156);
here the base directory is a plain module-level setting (``set_base_dir``) so the path works without Django —
inside the reference's Django process the integration shim simply passes ``settings.BASE_DIR``.

Hot-path contract reproduced:
* ``Context.compute(graph)``: for each node in ``graph.order()``: ``get_node(name)`` (KeyError for an unknown
  endpoint) -> ``node.compute(params, node.get_pinin())`` -> ``node.set_pinout(...)``    (context.py:143-147)
* ``Model``: wraps an ``nn.Module`` in eval mode; node names are ``"<model>:<dotted leaf path>"`` for every
  leaf module (context.py:39-47); ``compute`` runs ``get_submodule(path)(pinin["o"])`` under ``no_grad`` and
  returns channel ``"o"`` (context.py:79-88); ``io`` is ``{"ins": ["o"], "outs": ["o"]}`` (context.py:94-96);
  ``generate_graph_json`` lays the nodes on a floor(sqrt(n))-wide grid, 200 px apart, chained o->o
  (context.py:55-73); ``register`` writes ``static/graphs/<name>.json`` if absent and registers one
  ``ModelNode`` per name (context.py:98-112).
* ``scan_nodes``: import every ``*.py`` of the given sub-directories, call ``module.instances()``, register each
  instance; any exception is logged and swallowed (context.py:154-174).
"""
from __future__ import annotations

import importlib.util
import json
import logging
import math
import os
import sys
from typing import Dict, Iterable, List, Optional
from urllib.parse import urlencode

import torch

from .graph import Graph, Pinout

logger = logging.getLogger(__name__)

_base_dir: Optional[str] = None


def set_base_dir(path: Optional[str]) -> None:
    """Directory that holds ``static/graphs`` and the plugin directories (the reference's settings.BASE_DIR)."""
    global _base_dir
    _base_dir = None if path is None else str(path)


def get_base_dir() -> str:
    if _base_dir is not None:
        return _base_dir
    try:  # inside the reference's Django process
        from django.conf import settings  # type: ignore

        return str(settings.BASE_DIR)
    except Exception:
        return os.getcwd()


class NodeKind:
    """Abstract operator.  Subclasses override io() and compute()."""

    def __init__(self, name: str):
        self.name = name

    def get_name(self) -> str:
        return self.name

    def contents(self, params: Dict[str, str]) -> str:
        return f"{self.name}?{urlencode(params)}"

    def io(self, params: Dict[str, str]) -> Dict:
        raise Exception(f"TODO: implement Node.io() for {self.name}")

    def compute(self, params: Dict[str, str], inputs: Pinout) -> Pinout:
        raise Exception(f"TODO: implement Node.compute() for {self.name}")

    def register(self, ctx: "Context") -> None:
        ctx.register(self)


class Model:
    """A network whose sub-modules are exposed as graph nodes named ``<model>:<path>``."""

    def __init__(self, model: torch.nn.Module, name: str):
        self.model = model
        self.model.eval()
        self.name = name
        # leaves only: a module whose named_modules() yields nothing but itself
        self.node_names: List[str] = [
            self.prefix() + path for path, sub in self.model.named_modules() if len(list(sub.named_modules())) == 1
        ]

    def get_name(self) -> str:
        return self.name

    def prefix(self) -> str:
        return self.name + ":"

    def list_node_names(self) -> List[str]:
        return self.node_names

    def generate_graph_json(self) -> Dict:
        names = self.list_node_names()
        width = int(math.sqrt(len(names)))
        graph: Dict[str, list] = {"nodes": [], "edges": []}
        for i, endpoint in enumerate(names):
            graph["nodes"].append({
                "instance": {"kind": "net_node", "endpoint": f"{endpoint}", "params": {}},
                "pos": {"x": (i % width) * 200, "y": int(i / width) * 200},
            })
            if i > 0:
                graph["edges"].append({
                    "in_port": {"node": i - 1, "channel": "o"},
                    "out_port": {"node": i, "channel": "o"},
                })
        return graph

    def _submodule(self, node_name: str) -> torch.nn.Module:
        return self.model.get_submodule(node_name.removeprefix(self.prefix()))

    def compute(self, node_name: str, pinin: Pinout) -> Pinout:
        with torch.no_grad():
            sub = self._submodule(node_name)
            x = pinin.get("o")
            assert x is not None
            y = sub(x)
            assert isinstance(y, torch.Tensor)
            out = Pinout()
            out.set("o", y)
            return out

    def contents(self, node_name: str) -> str:
        return f"<p>{node_name}</p> <p>{self._submodule(node_name)._get_name()}</p>"

    def io(self, node_name: str) -> Dict:
        return {"ins": ["o"], "outs": ["o"]}

    def register(self, ctx: "Context") -> None:
        path = os.path.join(get_base_dir(), "static/graphs/" + self.name + ".json")
        if not os.path.exists(path):
            try:
                with open(path, "w") as f:
                    f.write(json.dumps(self.generate_graph_json()))
                logger.info("generated graph %s", path)
            except Exception as e:  # same policy as the reference: log, keep registering nodes
                logger.error("could not generate graph %s: %s", path, str(e))
        for node_name in self.list_node_names():
            ModelNode(self, node_name).register(ctx)


class ModelNode(NodeKind):
    """One node of a ``Model``; every call forwards to the parent with this node's name."""

    def __init__(self, parent: Model, name: str):
        super().__init__(name)
        self.parent = parent

    def compute(self, params: Dict[str, str], inputs: Pinout) -> Pinout:
        return self.parent.compute(self.get_name(), inputs)

    def contents(self, params: Dict[str, str]) -> str:
        return self.parent.contents(self.get_name())

    def io(self, params: Dict[str, str]) -> Dict:
        return self.parent.io(self.get_name())


class Context:
    """Registry of node kinds + the interpreter loop."""

    def __init__(self) -> None:
        self.nodes: Dict[str, NodeKind] = {}

    def register(self, node: NodeKind) -> None:
        logger.info("Registered node: '%s'", node.get_name())
        self.nodes[node.get_name()] = node  # last registration wins

    def get_node(self, name: str) -> NodeKind:
        return self.nodes[name]

    def compute(self, graph: Graph) -> None:
        for n in graph.order():
            kind = self.get_node(n.name)
            n.set_pinout(kind.compute(n.params, n.get_pinin()))


instance = Context()


def context() -> Context:
    return instance


def scan_nodes(dirs: Iterable[str], ctx: Optional[Context] = None) -> List[str]:
    """Load every plugin file below ``get_base_dir()/<dir>``; returns the files that registered."""
    ctx = ctx if ctx is not None else context()
    ok: List[str] = []
    for sub in dirs:
        full = os.path.join(get_base_dir(), sub)
        for fname in os.listdir(full):
            path = os.path.join(full, fname)
            if not (os.path.isfile(path) and path.endswith(".py")):
                continue
            mod_name = os.path.splitext(fname)[0]
            try:
                spec = importlib.util.spec_from_file_location(mod_name, path)
                module = importlib.util.module_from_spec(spec)
                sys.modules[mod_name] = module
                spec.loader.exec_module(module)
                for inst in module.instances():
                    inst.register(ctx)
                ok.append(path)
            except Exception as err:
                logger.info("Could not register '%s': %s", path, str(err))
    return ok
