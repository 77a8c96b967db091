"""Device-resident per-step scalars (riqn_dyn_state, include/riqn_b200.h) so that a learner step can be captured in a
CUDA graph once and replayed: the host rewrites this 32-byte struct with one async copy before each replay."""
import struct

import torch

_FMT = "<Qffdd"          # rng_offset, adam_neg_step_size, adam_sqrt_bc2, is_capacity, is_beta
_SIZE = struct.calcsize(_FMT)
_RING = 16


class DynState:
    def __init__(self, device):
        assert _SIZE == 32
        self.dev = torch.zeros(_SIZE, dtype=torch.uint8, device=device)
        self._host = [torch.zeros(_SIZE, dtype=torch.uint8).pin_memory() for _ in range(_RING)]
        self._events = [None] * _RING
        self._i = 0
        self.epoch = 0       # learner steps issued; Philox streams advance by 64 per epoch

    def ptr(self):
        return self.dev.data_ptr()

    def write(self, neg_step_size, sqrt_bc2, capacity, beta):
        """Stage the values of the NEXT step and enqueue the copy on the current stream."""
        slot = self._i % _RING
        if self._events[slot] is not None:
            self._events[slot].synchronize()          # the copy that last used this pinned slot has completed
        buf = self._host[slot]
        struct.pack_into(_FMT, buf.numpy(), 0, 64 * self.epoch, float(neg_step_size), float(sqrt_bc2), float(capacity),
                         float(beta))
        self.dev.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[slot] = ev
        self._i += 1
        self.epoch += 1
