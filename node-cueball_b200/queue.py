"""Queue -- the pending-claim FIFO with the reference's lib/queue.js surface
(push -> node with remove(), shift, isEmpty, length, forEach).

The reference keeps a sentinel doubly-linked list of JS objects (lib/queue.js:13-75).
Here the FIFO is what the device consumes: handles live in a host list in arrival order and
`snapshot()` hands the drain loop flat int64/uint8 arrays (enqueue timestamp, waiting
flag) -- 9 bytes per claim instead of a heap node per claim.
"""
import numpy as np


class QueueNode(object):
    __slots__ = ("n_value", "n_queue", "_alive")

    def __init__(self, queue, value):
        self.n_value = value
        self.n_queue = queue
        self._alive = True

    def remove(self):  # lib/queue.js:28-37, O(1): tombstone, skipped at shift()
        assert self._alive, "node already removed"
        self._alive = False
        self.n_queue.q_len -= 1


class Queue(object):
    def __init__(self):
        self._nodes = []
        self._head = 0
        self.q_len = 0

    def isEmpty(self):  # lib/queue.js:46-48
        return self.q_len == 0

    def peek(self):  # lib/queue.js:49-51: the SENTINEL's value, i.e. always null
        return None

    def push(self, v):  # lib/queue.js:52-56
        n = QueueNode(self, v)
        self._nodes.append(n)
        self.q_len += 1
        return n

    def shift(self):  # lib/queue.js:57-62 (callers check length first)
        while True:
            n = self._nodes[self._head]
            self._head += 1
            if n._alive:
                n._alive = False
                self.q_len -= 1
                self._compact()
                return n.n_value

    def forEach(self, cb):  # lib/queue.js:63-70
        for n in list(self._nodes[self._head:]):
            if n._alive:
                cb(n.n_value, n)

    @property
    def length(self):  # lib/queue.js:71-75
        return self.q_len

    def _compact(self):
        if self._head > 4096 and self._head * 2 > len(self._nodes):
            self._nodes = self._nodes[self._head:]
            self._head = 0

    def take_all(self):
        """Detach every live value in FIFO order (the batch drain)."""
        out = [n.n_value for n in self._nodes[self._head:] if n._alive]
        for n in self._nodes[self._head:]:
            n._alive = False
        self._nodes, self._head, self.q_len = [], 0, 0
        return out


def snapshot(handles):
    """handles -> (start_ns int64[n], flags uint8[n]) for the device batch."""
    n = len(handles)
    start = np.fromiter((h.ch_started_ns for h in handles), dtype=np.int64, count=n)
    flags = np.fromiter((1 if h.isInState("waiting") else 0 for h in handles), dtype=np.uint8, count=n)
    return start, flags
