"""Data-parallel training: one process per GPU (torchrun), identical replicas, gradient all-reduce over NCCL.

Replaces the reference's denet/multi: spawned worker processes exchanging the ENTIRE model state through host shared
memory and a numpy mean on one thread every iteration (multi/worker.py:85-122, multi/shared.py:105-119,
model/train_multi.py:100-139).  For the sgd and torch/nesterov solvers the update is linear in the gradient, so
averaging the gradients before one update equals the reference's average of the per-worker updated models from an
identical start state (batch_size_factor = 1); batch-norm running statistics are averaged once per step like every
other update target of the reference (shared.py:155-158).
"""
from .ddp import GradientAllReduce, init_process_group, shard_batch  # noqa: F401
