"""Frame sharding across GPUs (the reference's `--num_gpu N`: N workers popping one queue, outputs re-ordered by
frame index in buffer_and_order, examples/rtpose/rtpose.cpp:1463-1472, 1214-1273).

One process per GPU: frame i belongs to rank i % world; results are gathered to rank 0 and emitted in frame
order.  No data-path collective: the gather moves KB-sized joint lists (gloo on CPU tensors / objects, NCCL is
used only for the init weight broadcast in bench.py)."""
import heapq


def shard_indices(n_frames, rank, world):
    """Frame indices processed by `rank` (round-robin, like workers popping a shared queue in order)."""
    return list(range(rank, n_frames, world))


class Orderer:
    """Min-heap re-ordering of (index, payload) pairs arriving out of order, with dropped indices skipped
    (buffer_and_order, rtpose.cpp:1214-1273)."""

    def __init__(self, dropped=()):
        self.heap = []
        self.next = 0
        self.dropped = set(dropped)

    def push(self, index, payload):
        heapq.heappush(self.heap, (index, payload))
        out = []
        while True:
            while self.next in self.dropped:
                self.next += 1
            if self.heap and self.heap[0][0] == self.next:
                out.append(heapq.heappop(self.heap))
                self.next += 1
            else:
                break
        return out


def gather_ordered(local_results, rank, world, dist):
    """local_results: list of (frame_index, payload) on every rank -> ordered list on rank 0 (None elsewhere)."""
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local_results, gathered, dst=0)
    if rank != 0:
        return None
    o = Orderer()
    out = []
    for lst in gathered:
        for idx, payload in lst:
            out.extend(o.push(idx, payload))
    return out
