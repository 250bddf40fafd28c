"""Graph IR at the node boundary: ``Node`` / ``Port`` / ``Edge`` / ``Graph`` / ``Pinout``.

Mirror of the reference's main/graph.py:6-132 — same class names, attributes, method names and observable
behaviour (including its quirks, which the parity tests pin):

* tensors travel by reference on ``Edge.tensor``; a ``Pinout`` is a ``{channel: tensor}`` bag (graph.py:123-132);
* ``Node.set_pinout`` creates a dangling out-edge for a channel nobody consumes, so the value still reaches the
  response (graph.py:22-29);
* ``Graph.connect`` keeps ONE edge per output channel — a second consumer of the same channel overwrites the
  first producer-side record (graph.py:64-70), so fan-out must be expressed inside a node;
* ``Graph.order`` is the reference's tail-pop / head-requeue worklist (graph.py:79-99); the visit order — not
  just any topological order — is reproduced because it fixes the order of Model.compute calls.  Unlike the
  reference it raises on a cycle instead of spinning forever.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple
from urllib.parse import urlencode

import torch


class Pinout:
    """Channel-name -> tensor bag handed to / returned from ``NodeKind.compute``."""

    def __init__(self) -> None:
        self.pinout: Dict[str, torch.Tensor] = {}

    def set(self, ch: str, t: torch.Tensor) -> None:
        self.pinout[ch] = t

    def get(self, ch: str) -> Optional[torch.Tensor]:
        return self.pinout.get(ch)

    def items(self) -> Iterator[Tuple[str, torch.Tensor]]:
        return iter(self.pinout.items())


class Port:
    def __init__(self, node: "Node", channel: str, direction: str) -> None:
        self.node = node
        self.channel = channel
        self.direction = direction


class Edge:
    """Producer port (``input``; None for a wire tensor) -> consumer port (``output``; None when dangling)."""

    def __init__(self, src: Optional[Port], tgt: Optional[Port]) -> None:
        assert src is None or src.direction == "out"
        assert tgt is None or tgt.direction == "in"
        self.input = src
        self.output = tgt
        self.tensor: Optional[torch.Tensor] = None
        self.taps: List["Edge"] = []   # further consumers of the same producer channel (Graph(fan_out=True) only)


class Node:
    def __init__(self, name: str, params: Dict[str, str], index: int):
        self.name = name
        self.params = params
        self.index = index
        self.inputs: Dict[str, Edge] = {}
        self.outputs: Dict[str, Edge] = {}

    @staticmethod
    def _collect(edges: Dict[str, Edge]) -> Pinout:
        bag = Pinout()
        for ch, edge in edges.items():
            assert edge.tensor is not None
            bag.set(ch, edge.tensor)
        return bag

    def get_pinin(self) -> Pinout:
        return self._collect(self.inputs)

    def get_pinout(self) -> Pinout:
        return self._collect(self.outputs)

    def set_pinout(self, pinout: Pinout) -> None:
        for ch, t in pinout.pinout.items():
            edge = self.outputs.get(ch)
            if edge is None:
                edge = self.outputs[ch] = Edge(Port(self, ch, "out"), None)
            edge.tensor = t
            for tap in edge.taps:
                tap.tensor = t

    def label(self) -> str:
        return self.name + "?" + urlencode(self.params)


class Graph:
    def __init__(self, fan_out: bool = False) -> None:
        self.nodes: List[Node] = []
        self.fan_out = fan_out

    def add_node(self, name: str, params: Dict[str, str]) -> Node:
        node = Node(name, params, len(self.nodes))
        self.nodes.append(node)
        return node

    def connect(self, a: Node, a_ch: str, b: Node, b_ch: str) -> Edge:
        edge = Edge(Port(a, a_ch, "out"), Port(b, b_ch, "in"))
        first = a.outputs.get(a_ch) if self.fan_out else None
        if first is not None:
            first.taps.append(edge)      # extension: a second consumer taps the first consumer's edge
        else:
            a.outputs[a_ch] = edge       # reference behaviour: the newest consumer replaces the record
        b.inputs[b_ch] = edge
        return edge

    def add_input(self, value: torch.Tensor, node: Node, channel: str) -> Edge:
        edge = Edge(None, Port(node, channel, "in"))
        edge.tensor = value
        node.inputs[channel] = edge
        return edge

    def order(self) -> List[Node]:
        done: set = set()
        ordered: List[Node] = []
        pending = list(self.nodes)
        stalled = 0  # consecutive re-queues; == len(pending) means no node can make progress (cycle)
        while pending:
            cand = pending.pop()
            ready = all(e.input is None or e.input.node in done for e in cand.inputs.values())
            if ready:
                done.add(cand)
                ordered.append(cand)
                stalled = 0
            else:
                pending.insert(0, cand)
                stalled += 1
                if stalled > len(pending):
                    raise ValueError("graph has a cycle: no schedulable node among " +
                                     ", ".join(n.name for n in pending))
        return ordered

    def __str__(self) -> str:
        lines = ["graph:"]
        for node in self.nodes:
            me = node.label()
            for ch, e in node.outputs.items():
                tgt = e.output.node.label() if e.output is not None else "*"
                shape = f" {e.tensor.shape}" if e.tensor is not None else ""
                lines.append(f"\t{me} --[{ch}]--> {tgt}{shape}")
            for ch, e in node.inputs.items():
                if e.input is None:
                    assert e.tensor is not None
                    lines.append(f"\t* --[{ch}]--> {me} {e.tensor.shape}")
        return "\n".join(lines)
