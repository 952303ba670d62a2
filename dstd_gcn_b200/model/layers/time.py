"""Static temporal adjacency (`/root/reference/model/layers/time.py:4-41`).

The reference's "neighboor" matrix is built by two overlapping block
assignments (time.py:19-22) whose second write clobbers part of the first, so
the result is NOT tridiagonal: it is the sub-diagonal shift matrix plus ones at
``[0,0] [0,1] [T-2,T-1] [T-1,T-1]``.  That matrix is a frozen parameter in every
reference checkpoint, so it is reproduced here by direct construction and
pinned against the reference in ``tests/test_layers.py``.
"""
import numpy as np


class Time:

    def __init__(self, seq_length):
        self.seq_length = seq_length
        self.input_length = 10
        self.output_length = seq_length - self.input_length

    def _neighbour(self):
        t = self.seq_length
        adj = np.zeros((t, t))
        if t == 1:
            adj[0, 0] = 1
            return adj
        adj[np.arange(1, t), np.arange(0, t - 1)] = 1   # what survives of both block writes
        adj[0, 0] = adj[0, 1] = 1                        # first row is never overwritten
        adj[t - 2, t - 1] = adj[t - 1, t - 1] = 1        # last column is never overwritten
        return adj

    def _inout(self):
        t, i = self.seq_length, self.input_length
        adj = np.zeros((t, t))
        adj[:i, i:] = 1
        adj[i:, :i] = 1
        return adj

    def get_adjacency(self):
        return self._neighbour()

    def get_adjacency_type(self, type="self"):
        if type == "self":
            return np.eye(self.seq_length)
        if type == "neighboor":
            return self._neighbour()
        if type == "inout":
            return self._inout()
        if type == "all":
            adj = self._neighbour()
            i = self.input_length
            adj[:i, i:] = 1
            adj[i:, :i] = 1
            return adj
        raise ValueError(f"Invalid graph type {type}")

    def get_all_adjacency(self):
        return np.stack([self.get_adjacency_type("neighboor")], axis=0)
