"""interactive-vit_b200 — B200-native ViT forward engine behind interactive-vit's node/operator API.

Layout (only what the hot path needs):
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/vitb200.h) -> libvitb200.so
  engine.py    ctypes binding (the FFI stub a reference maintainer would add)
  graph.py     Node / Port / Edge / Graph / Pinout      (mirror of main/graph.py)
  context.py   NodeKind / Model / ModelNode / Context    (mirror of main/context.py, no Django needed)
  message.py   binary wire codec Request / Response      (mirror of main/message.py)
  vit_plugin.py  the ViT `Model` plugin whose compute() runs on the engine
  dist.py      data-parallel sharding + gather to rank 0 (torch.distributed)

The directory name carries a hyphen, so Python imports go through the `interactive_vit_b200` shim package at
the repository root, which points its __path__ here.
"""
__all__ = ["engine", "graph", "context", "message", "vit_plugin", "dist"]
